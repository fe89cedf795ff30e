// One-node sum all-reduce over NVLink peer memory (SURVEY.md §8b `apb_allreduce`; fit/lm.py:256-260 on sharded pixels).
//
// What a fit sharded over GPUs exchanges per LM evaluation is small: J^T W J | J^T W r | chi^2 (P^2 + P + 2 doubles for a
// joint fit, the block-sparse array of a crowded field: <= a few MB) and P + 3 doubles per lambda-trial.  At these sizes
// a collective is pure latency, so it is ONE kernel, stream-ordered behind the kernel that produced the buffer:
//   1. every rank copies its buffer into its own exchange slot (device memory exported with cudaIpc, mapped by every
//      peer) and, when the copy is complete, stores the call's sequence number into its flag in EVERY peer's flag array
//      (peer stores over NVLink, system-scope release);
//   2. every rank spins on its own (local) flag array until all peers have signalled, then reads all slots -- its own
//      and the peers', through the NVSwitch -- and adds them IN RANK ORDER: the sum is deterministic and bit-identical
//      on every rank (NCCL's ring / tree orders are neither across ranks' roles nor across sizes).
// Slots are double-buffered by the parity of the sequence number: a rank can run at most one call ahead of the slowest
// peer (it cannot pass the flag wait of call k+1 before every peer has finished reading call k), so the slot being
// overwritten is never one a peer still reads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define APB_COMM_MAX_RANKS 16

struct apb_comm {
  int rank = 0, world = 1;
  size_t slot_doubles = 0;                 // capacity of one slot
  void* local = nullptr;                   // this rank's exchange buffer: [slot 0 | slot 1 | flags (APB_COMM_MAX_RANKS x u64)]
  void* peer[APB_COMM_MAX_RANKS] = {};     // every rank's exchange buffer as mapped here (peer[rank] == local)
  unsigned long long seq = 0;              // calls made so far
  unsigned int* done = nullptr;            // CTA completion counter of phase 1
  int grid_max = 1;
};

struct CommArgs {
  double* slot[APB_COMM_MAX_RANKS];               // slot of this call's parity in every rank's buffer
  unsigned long long* flags[APB_COMM_MAX_RANKS];  // flag array of every rank
  int rank, world;
  unsigned long long seq;
  unsigned int* done;
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// peer data: volatile (never from a stale L1 line), no memory clobber -- the loads of one element from all ranks are
// independent and must be in flight together (a clobber would serialise them: world x one NVLink round trip)
__device__ __forceinline__ double ld_peer(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// grid <= the number of co-resident CTAs (every CTA spins in phase 2 while the others may still be in phase 1)
__global__ void __launch_bounds__(256) k_allreduce_peer(CommArgs A, double* __restrict__ buf, size_t n) {
  __shared__ bool last_s;
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = gridDim.x * (size_t)blockDim.x;
  // ---- phase 1: publish
  double* mine = A.slot[A.rank];
  for (size_t i = tid; i < n; i += nth) mine[i] = buf[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();          // this CTA's copy is visible device-wide before it counts as arrived
    last_s = atomicAdd(A.done, 1u) == gridDim.x - 1;
    __threadfence();          // ... and the last CTA sees every other CTA's copy before it raises the flags
  }
  __syncthreads();
  if (last_s) {
    // every CTA has arrived: one system-scope release per peer publishes the whole slot (cumulativity through the
    // device-scope fences and the arrival counter), then the flag -- this rank's included
    if (threadIdx.x < A.world) st_release_sys(A.flags[threadIdx.x] + A.rank, A.seq);
    if (threadIdx.x == 0) *A.done = 0u;
  }
  // ---- phase 2: wait for all ranks (relaxed polling of LOCAL memory, one acquire fence at the end), then sum in rank order
  if (threadIdx.x < A.world) {
    const unsigned long long* f = A.flags[A.rank] + threadIdx.x;
    while (ld_volatile_u64(f) < A.seq) {
    }
    __threadfence_system();
  }
  __syncthreads();
  for (size_t i = tid; i < n; i += nth) {
    double v[APB_COMM_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < APB_COMM_MAX_RANKS; ++r)
      if (r < A.world) v[r] = ld_peer(A.slot[r] + i);
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < APB_COMM_MAX_RANKS; ++r)
      if (r < A.world) t += v[r];
    buf[i] = t;
  }
}
