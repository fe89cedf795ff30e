// Damped LM solve for large parameter counts (crowded fields): block-sparse preconditioned
// conjugate gradients, one persistent cooperative kernel per solve.
//
// The reference forms the dense P x P matrix  A = H o (I + (1-I)/(1+L)) + L I (1 + diag H)  and calls
// torch.linalg.solve (fit/lm.py:359-371).  For a crowded field H = J^T W J is block-sparse: a model
// only couples to the models its window overlaps (and to the sky), so the matrix is kept as the
// list of <= 8x8 blocks the normal-equation kernels produce -- stored tightly (n_a x n_b doubles per
// block: a galaxy-star block is 7x3, not 8x8) -- and A p is a block-sparse product.  A is symmetric
// positive definite for L > 0, the preconditioner is block-Jacobi (Cholesky of the damped diagonal
// block of every model), which removes the ill-conditioning inside a model (Sersic n / Re / Ie);
// what is left -- overlaps and the sky row -- converges in a few dozen to a few hundred iterations.
//
// One launch runs the whole iteration with TWO grid barriers per iteration:
//   phase 1  q = A p  with  p = z + beta p_old  formed on the fly (the lanes that own a row also store
//            p), and the CTA's share of p.q;
//   phase 2  alpha from the summed shares;  x += alpha p,  r -= alpha q,  z = M^-1 r  per row block, and
//            the CTA's shares of r.z and r.r.
// Dot products are never a pass over the vectors: every CTA writes its share, and after the barrier
// every CTA adds the shares in the same order, so alpha, beta and the stopping decision are
// bit-identical in all CTAs without a broadcast.  Rows whose product is split over several work items
// (the sky row couples to every model) write partial rows that every CTA adds in order:
// no atomics, run-to-run deterministic.  DESIGN.md section 4 has the measurements behind the layout.
#pragma once
#include <cooperative_groups.h>
#include "apb_internal.cuh"
#include "apb_image.cuh"

namespace cg = cooperative_groups;
#ifndef PCG_NT
#define PCG_NT 256     // threads per CTA, and CTAs per SM (the grid barrier's cost grows with the number of CTAs)
#define PCG_MINB 2
#endif
#ifndef PCG_SPIN_NS
#define PCG_SPIN_NS 200
#endif

struct PcgRow {      // row block: free parameters [p0, p0+n) of owner `src`
  int src, p0, n;
  int item0, nitem;  // its work items (nitem > 1: product split over several warps)
  long long doff;    // offset of its diagonal block (n x n, row-major) in bvals
  int slot0, _pad;   // act_off[src] + p0: where its parameter indices start in act_slot (one dependent load less)
};
struct PcgEntry {    // one block contributing to a row block (host side: the packed tables below are built from these)
  long long off;     // offset of the block in bvals (row-major [i of a][j of b], leading dimension ld)
  int ld;
  int transposed;    // 1: the row block is the block's b side
  int n;             // parameters of the other side
  int sl[NB_MAX];    // their indices in x
};
// The product walks the matrix in the order the lanes want it: a warp takes 4, 2 or 1 work items at a time (a "pass":
// 8, 16 or 32 lanes per item, by the number of blocks in the row, so that nearly every row is done in ONE sweep), each
// lane one block of its item per "sweep", and the values of a sweep are stored element-major, lane-minor -- element
// (i, j) of all 32 lanes' blocks in 32 consecutive doubles, ni x nj elements (the largest block of the pass; smaller
// ones are zero-padded), transposition resolved.  Read straight from the tightly packed blocks a
// lane-per-block product touched 32 different lines per load instruction, and the load/store unit's serialisation of
// those was the solver's critical path (20 of 33 us per iteration).  k_pcg_pack refreshes the packed copy from the
// block values before every solve.
struct PcgPass {     // one warp's share of the product: `nit` work items, g lanes each, `nsweep` sweeps of one block per lane
  long long voff;    // its values in `packed`: nsweep x (ni x nj x 32)
  int soff;          // its x indices in `pslots`: nsweep x (nj x 32)
  int ni, nj, nsweep;
  int item0, nit, g, _pad;
};
struct PcgItem {     // work item (one warp): entries [e0, e1) of row block rb
  int rb, e0, e1;
  int multi;         // 0: the only item of its row; 1: first, 2: further item of a split row
  int n, slot0;      // copies of the row's size and slot start: the product does not wait for the row record
};

struct PcgArgs {
  const PcgRow* rows; int n_rows;
  const PcgPass* passes; int n_pass;
  const double* packed; const int* pslots;
  const PcgItem* items; int n_items;
  const int* multi_rows; int n_multi;   // rows split over several items
  const int* act_slot; const int* act_off;
  const double* bvals;     // tightly packed blocks
  const double* diagH;     // P: diagonal of H
  double* fac;             // n_rows x 64: Cholesky factors of the damped diagonal blocks
  const double* b;         // right-hand side (P)
  const double* x0;        // starting point (P) or NULL = 0 (the previous lambda-trial's solution is a good one)
  double* x;               // solution (P)
  double *zp0, *zp1;       // work vectors: (z, p) pairs, double-buffered (2P each, 16-byte aligned)
  double *r, *q;           // ... residual and A p (P)
  double* qpart;           // n_items x 8: partial rows of split rows
  double* part;            // gridDim x 4: per-CTA shares of the dot products
  unsigned int* barrier;   // arrival counter of pcg_barrier, zero at launch
  double* info;            // {iterations, final |r|/|b|}
  int P, max_iter;
  double L, tol;
};

// Grid barrier for the (cooperatively launched, hence co-resident) CTAs of k_pcg.  cooperative_groups' grid.sync() --
// and any acquire fence: __threadfence() compiles to MEMBAR + CCTL.IVALL -- invalidates the whole L1 at every barrier,
// which sends every load of the solver's read-only tables (passes, x indices, the packed matrix: the same ones every
// iteration, ~130 KB per SM) back to L2; the product then was a chain of L2 round trips (ncu: 18 us of a 25 us
// iteration).  Here the arrival is a release-ordered reduction (everything this CTA wrote is in L2 before the count
// moves), the wait is a relaxed poll with NO acquire fence, and every vector that changes hands between CTAs is read with
// ld.cg (L2) after the block barrier that follows the poll: L1 only ever holds data nobody writes during the kernel
// (or that only its own thread writes), so it stays valid.
__device__ __forceinline__ void pcg_barrier(unsigned int* counter, unsigned int& goal) {
  __syncthreads();
  if (threadIdx.x == 0) {
    goal += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
    // (back off between polls: a few hundred CTAs spinning on one L2 line delay the arrivals they are waiting for)
    unsigned int seen;
    for (;;) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (seen >= goal) break;
      __nanosleep(PCG_SPIN_NS);
    }
  }
  __syncthreads();
}

// Sum of NV values over the CTA's threads, result in every thread (same bits: the butterfly's additions commute),
// two block barriers for all NV.
template <int NV>
__device__ __forceinline__ void block_sum_n(double (&v)[NV]) {
  __shared__ double shn[PCG_NT / 32 * NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) shn[(threadIdx.x >> 5) * NV + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < PCG_NT / 32; ++w) t += shn[w * NV + k];
    v[k] = t;
  }
  __syncthreads();
}

// two columns of the per-CTA shares in one pass (r.z and r.r are always wanted together)
__device__ __forceinline__ void pcg_total2(const double* part, int col_a, int col_b, int ncta, double& ta, double& tb) {
  double v[2] = {0.0, 0.0};
  for (int k = threadIdx.x; k < ncta; k += PCG_NT) {
    v[0] += __ldcg(part + 4 * k + col_a);
    v[1] += __ldcg(part + 4 * k + col_b);
  }
  block_sum_n<2>(v);
  ta = v[0];
  tb = v[1];
}

// z = M^-1 r for the diagonal block of one row block (one thread): L L^T z = r.  Returns r.z of the block.
// The factor (<= 36 numbers) is fetched with independent loads BEFORE the substitutions: read element by element inside
// them it was a chain of ~70 dependent L2 round trips, the critical path of every iteration (ncu: half of all warp
// samples waited at the barrier behind it).  Rows beyond the block's size are identity rows.
__device__ __forceinline__ double pcg_precond(const PcgArgs& A, int rb, const PcgRow& row, const double* rv, double* __restrict__ z) {
  const int* sl = A.act_slot + row.slot0;
  const double* F = A.fac + (long long)rb * 64;
  const int n = row.n;
  double Fr[NB_MAX * (NB_MAX + 1) / 2];
  int slr[NB_MAX];
#pragma unroll
  for (int i = 0; i < NB_MAX; ++i) {
    slr[i] = i < n ? sl[i] : 0;
#pragma unroll
    for (int k = 0; k <= i; ++k) Fr[i * (i + 1) / 2 + k] = i < n ? F[i * 8 + k] : (i == k ? 1.0 : 0.0);
  }
  double y[NB_MAX];
#pragma unroll
  for (int i = 0; i < NB_MAX; ++i) {
    double v = i < n ? rv[i] : 0.0;
#pragma unroll
    for (int k = 0; k < i; ++k) v -= Fr[i * (i + 1) / 2 + k] * y[k];
    y[i] = v * Fr[i * (i + 1) / 2 + i];
  }
  double dot = 0.0;
#pragma unroll
  for (int i = NB_MAX - 1; i >= 0; --i) {
    double v = y[i];
#pragma unroll
    for (int k = i + 1; k < NB_MAX; ++k) v -= Fr[k * (k + 1) / 2 + i] * y[k];
    y[i] = v * Fr[i * (i + 1) / 2 + i];
    if (i < n) {
      z[2 * slr[i]] = y[i];
      dot = fma(rv[i], y[i], dot);
    }
  }
  return dot;
}

// packed copy of the block values (see PcgSlice): out[t] = bvals[src[t]], 0 where src[t] < 0
__global__ void __launch_bounds__(256) k_pcg_pack(const double* __restrict__ bvals, const int* __restrict__ src,
                                                  double* __restrict__ out, long long n) {
  for (long long t = blockIdx.x * 256ll + threadIdx.x; t < n; t += gridDim.x * 256ll) {
    const int o = src[t];
    out[t] = o >= 0 ? bvals[o] : 0.0;
  }
}

#ifdef PCG_TIMING   // development builds only: where the time of an iteration goes (globaltimer ns, CTA 0 and the last CTA)
__device__ unsigned long long g_pcg_clk[2][8];
__device__ __forceinline__ unsigned long long pcg_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PCG_MARK(k)                                                                     \
  if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {           \
    const unsigned long long t_ = pcg_now();                                            \
    g_pcg_clk[blockIdx.x == 0 ? 0 : 1][k] += t_ - tmark;                                \
    tmark = t_;                                                                         \
  }
#else
#define PCG_MARK(k)
#endif
__global__ void __launch_bounds__(PCG_NT, PCG_MINB) k_pcg(PcgArgs A) {
  unsigned int goal = 0;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const int ncta = gridDim.x;
  const double inv1L = 1.0 / (1.0 + A.L);

  // ---- setup: Cholesky of every damped diagonal block; x = 0, r = b, z = M^-1 r, p_old = 0.
  //      With a starting point the first pass of the loop below is a PRE-pass: z holds x0, so that phase 1 forms
  //      q = A x0, and phase 2 turns it into r = b - q, z = M^-1 r; the iteration proper starts from there.
  bool pre = A.x0 != nullptr;
  double s_rz = 0.0, s_bb = 0.0;
  // (row blocks are dealt CTA-first -- row rb to thread rb / #CTAs of CTA rb % #CTAs -- so that every SM works on them
  //  and each keeps its own few factors in L1; the same mapping in setup and in phase 2)
  const int rb0 = blockIdx.x + gridDim.x * threadIdx.x;
  for (int rb = rb0; rb < A.n_rows; rb += gsz) {
    const PcgRow row = A.rows[rb];
    const double* V = A.bvals + row.doff;
    double* F = A.fac + (long long)rb * 64;
    double M[NB_MAX][NB_MAX];
    for (int i = 0; i < row.n; ++i)
      for (int j = 0; j <= i; ++j) {
        const double h = V[i * row.n + j];
        M[i][j] = (i == j) ? h + A.L * (1.0 + h) : h * inv1L;
      }
    for (int j = 0; j < row.n; ++j) {
      double d = M[j][j];
      for (int k = 0; k < j; ++k) d -= M[j][k] * M[j][k];
      d = sqrt(fmax(d, 1e-300));
      M[j][j] = d;
      for (int i = j + 1; i < row.n; ++i) {
        double v = M[i][j];
        for (int k = 0; k < j; ++k) v -= M[i][k] * M[j][k];
        M[i][j] = v / d;
      }
    }
    for (int i = 0; i < row.n; ++i)   // (the diagonal is stored inverted: the substitutions multiply)
      for (int j = 0; j <= i; ++j) F[i * 8 + j] = i == j ? 1.0 / M[i][j] : M[i][j];
    const int* sl = A.act_slot + row.slot0;
    double rv[NB_MAX];
    for (int i = 0; i < row.n; ++i) {
      const double bi = A.b[sl[i]];
      rv[i] = bi;
      A.x[sl[i]] = pre ? A.x0[sl[i]] : 0.0;
      A.r[sl[i]] = bi;
      A.zp0[2 * sl[i] + 1] = 0.0;
      s_bb = fma(bi, bi, s_bb);
    }
    if (pre) {
      for (int i = 0; i < row.n; ++i) A.zp0[2 * sl[i]] = A.x0[sl[i]];
    } else {
      s_rz += pcg_precond(A, rb, row, rv, A.zp0);
    }
  }
  {
    double v[2] = {s_rz, s_bb};
    block_sum_n<2>(v);
    if (threadIdx.x == 0) { A.part[4 * blockIdx.x + 1] = v[0]; A.part[4 * blockIdx.x + 2] = v[1]; }
  }
  pcg_barrier(A.barrier, goal);
  double rz, bb_;
  pcg_total2(A.part, 1, 2, ncta, rz, bb_);
  const double bb = bb_;
  double rr = bb, beta = 0.0;
  // (z, p) pairs: the product reads z and the previous direction of a parameter with ONE 16-byte request
  double* pold = A.zp0;
  double* pnew = A.zp1;
  int it = 0;
  PcgRow mrow0 = {0, 0, 0, 0, 0, 0, 0, 0};
  if (A.n_multi > 0) mrow0 = A.rows[A.multi_rows[0]];
#ifdef PCG_TIMING
  unsigned long long tmark = pcg_now();
#endif
  if (bb > 0.0) {
    for (; it < A.max_iter;) {
      // ---- phase 1: p = z + beta p_old (on the fly), q = A p, share of p.q.  A group of lanes per work item.
      // What the phase costs is its chain of dependent loads (pass -> x indices -> p of the other side -> lane
      // reduction -> store), a few microseconds however little arithmetic hangs on it: the grid covers all passes at
      // once (heaviest first, dealt CTA-first so that every SM gets its share), and a pass is one sweep for all but
      // the rows with more than 32 blocks.
      double s_pq = 0.0;
      // (round by round the warps are dealt passes in alternating order: the warp with the heaviest pass of one round
      //  gets the lightest of the next)
      for (int rnd = 0;; ++rnd) {
        const int wic = threadIdx.x >> 5, NW = PCG_NT / 32;
        if (rnd * NW * (int)gridDim.x >= A.n_pass) break;
        const int wp = blockIdx.x + gridDim.x * (rnd * NW + ((rnd & 1) ? NW - 1 - wic : wic));
        if (wp >= A.n_pass) continue;
        const PcgPass ps = A.passes[wp];
        const int G = ps.g, sub = lane / G, gl = lane & (G - 1);
        const int k = ps.item0 + sub;
        PcgItem w = {0, 0, 0, 0, 0, 0};
        if (sub < ps.nit) w = A.items[k];
        // Lane l accumulates its blocks' contribution to all row elements; a shuffle tree over the group adds the lanes
        // (fixed order).
        double racc[NB_MAX];
#pragma unroll
        for (int i = 0; i < NB_MAX; ++i) racc[i] = 0.0;
        const double* V = A.packed + ps.voff + lane;
        const int* SL = A.pslots + ps.soff + lane;
        for (int sw = 0; sw < ps.nsweep; ++sw) {
          // (loads in batches, each batch issued before its first use: written load-by-load the compiler kept ONE value
          //  register and waited out every load's latency in turn.  The first two rows of values are requested before
          //  the directions they multiply have arrived, every further pair of rows under the arithmetic of the previous)
          double pj[NB_MAX];
          int sjv[NB_MAX];
          double2 zp[NB_MAX];
#pragma unroll
          for (int j = 0; j < NB_MAX; ++j) sjv[j] = j < ps.nj ? SL[j * 32] : -1;   // (-1: beyond this lane's block, or no block)
#pragma unroll
          for (int j = 0; j < NB_MAX; ++j) zp[j] = sjv[j] >= 0 ? __ldcg((const double2*)pold + sjv[j]) : make_double2(0.0, 0.0);
          double va[NB_MAX], vb[NB_MAX];
#pragma unroll
          for (int j = 0; j < NB_MAX; ++j) {      // (padding is never fetched: per-lane predicates on the loads)
            va[j] = (sjv[j] >= 0 && 0 < w.n) ? V[j * 32] : 0.0;
            vb[j] = (sjv[j] >= 0 && 1 < w.n) ? V[(ps.nj + j) * 32] : 0.0;
          }
#pragma unroll
          for (int j = 0; j < NB_MAX; ++j) pj[j] = fma(beta, zp[j].y, zp[j].x);
#pragma unroll
          for (int i = 0; i < NB_MAX; i += 2) {
            if (i < ps.ni) {
              double na[NB_MAX], nb[NB_MAX];
#pragma unroll
              for (int j = 0; j < NB_MAX; ++j) {
                na[j] = (i + 2 < NB_MAX && sjv[j] >= 0 && i + 2 < w.n) ? V[((i + 2) * ps.nj + j) * 32] : 0.0;
                nb[j] = (i + 3 < NB_MAX && sjv[j] >= 0 && i + 3 < w.n) ? V[((i + 3) * ps.nj + j) * 32] : 0.0;
              }
              double a = 0.0, b = 0.0;
#pragma unroll
              for (int j = 0; j < NB_MAX; ++j) {
                a = fma(va[j], pj[j], a);
                b = fma(vb[j], pj[j], b);
              }
              racc[i] += a;
              racc[i + 1] += b;
#pragma unroll
              for (int j = 0; j < NB_MAX; ++j) { va[j] = na[j]; vb[j] = nb[j]; }
            }
          }
          V += ps.ni * ps.nj * 32;
          SL += ps.nj * 32;
        }
        __syncwarp();
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < NB_MAX; ++i) {
          double v = racc[i];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
            if (o < G) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (gl == i) acc = v;
        }
        if (gl < w.n) {
          const int sl = A.act_slot[w.slot0 + gl];
          const double2 zp = __ldcg((const double2*)pold + sl);
          const double pi = fma(beta, zp.y, zp.x);
          double v = acc * inv1L;
          if (w.multi < 2) {   // the row's first (or only) item: the damped diagonal, and it stores p
            const double d = A.diagH[sl];
            v += (d + A.L * (1.0 + d) - d * inv1L) * pi;
            pnew[2 * sl + 1] = pi;
          }
          if (!w.multi) A.q[sl] = v;
          else A.qpart[(long long)k * 8 + gl] = v;
          s_pq = fma(pi, v, s_pq);   // (a split row's p.q is the sum of its items' p.(partial row))
        }
      }
      PCG_MARK(0)
      {
        double v[1] = {s_pq};
        block_sum_n<1>(v);
        if (threadIdx.x == 0) A.part[4 * blockIdx.x + 0] = v[0];
      }
      pcg_barrier(A.barrier, goal);
      PCG_MARK(1)
      // ---- alpha.  Split rows: every CTA adds their partial rows (same order everywhere) and stores q; the row's
      //      owner reads the copy its own CTA wrote.  One fused reduction: p.q and up to three row elements at a time.
      double pq = 0.0;
      {
        int m = 0, i = 0;
        for (bool first = true; first || m < A.n_multi; first = false) {
          double v[4] = {0.0, 0.0, 0.0, 0.0};
          int dst[4] = {-1, -1, -1, -1};
          int c = 0;
          if (first) {
            for (int k = threadIdx.x; k < ncta; k += PCG_NT) v[0] += __ldcg(A.part + 4 * k + 0);
            c = 1;
          }
          for (; c < 4 && m < A.n_multi; ++c) {
            const PcgRow row = m == 0 ? mrow0 : A.rows[A.multi_rows[m]];   // (the first -- usually the only -- one is held in registers)
            for (int t = threadIdx.x; t < row.nitem; t += PCG_NT) v[c] += __ldcg(A.qpart + (long long)(row.item0 + t) * 8 + i);
            dst[c] = A.act_slot[row.slot0 + i];
            if (++i == row.n) { i = 0; ++m; }
          }
          block_sum_n<4>(v);
          if (first) pq = v[0];
          if (threadIdx.x == 0) {
#pragma unroll
            for (int c2 = 0; c2 < 4; ++c2)
              if (dst[c2] >= 0) A.q[dst[c2]] = v[c2];
          }
        }
        __syncthreads();
      }
      PCG_MARK(2)
      if (!pre && !(pq > 0.0)) break;
      const double alpha = pre ? 0.0 : rz / pq;
      // ---- phase 2: x += alpha p; r -= alpha q; z = M^-1 r; shares of r.z and r.r
      double s_rz2 = 0.0, s_rr = 0.0;
      for (int rb = rb0; rb < A.n_rows; rb += gsz) {
        const PcgRow row = A.rows[rb];
        const int* sl = A.act_slot + row.slot0;
        double rv[NB_MAX];
        // (all loads of the row issued before the first use)
        int ss[NB_MAX];
        double qv[NB_MAX], pv[NB_MAX], xv[NB_MAX], rr0[NB_MAX];
#pragma unroll
        for (int i = 0; i < NB_MAX; ++i) ss[i] = i < row.n ? sl[i] : -1;
#pragma unroll
        for (int i = 0; i < NB_MAX; ++i) {
          const bool on = ss[i] >= 0;
          qv[i] = on ? __ldcg(A.q + ss[i]) : 0.0;
          pv[i] = (on && !pre) ? __ldcg(pnew + 2 * ss[i] + 1) : 0.0;
          xv[i] = (on && !pre) ? A.x[ss[i]] : 0.0;
          rr0[i] = on ? A.r[ss[i]] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < NB_MAX; ++i) {
          rv[i] = 0.0;
          if (ss[i] < 0) continue;
          double ri;
          if (pre) {
            ri = rr0[i] - qv[i];                        // r = b - A x0
          } else {
            A.x[ss[i]] = fma(alpha, pv[i], xv[i]);
            ri = fma(-alpha, qv[i], rr0[i]);
          }
          A.r[ss[i]] = ri;
          rv[i] = ri;
          s_rr = fma(ri, ri, s_rr);
        }
        s_rz2 += pcg_precond(A, rb, row, rv, pnew);
      }
      PCG_MARK(3)
      {
        double v[2] = {s_rz2, s_rr};
        block_sum_n<2>(v);
        if (threadIdx.x == 0) { A.part[4 * blockIdx.x + 1] = v[0]; A.part[4 * blockIdx.x + 2] = v[1]; }
      }
      pcg_barrier(A.barrier, goal);
      PCG_MARK(4)
      double rz_new;
      pcg_total2(A.part, 1, 2, ncta, rz_new, rr);
      PCG_MARK(5)
      if (pre) {            // the iteration proper starts here: p = z
        pre = false;
        rz = rz_new;
        beta = 0.0;
        if (!(rr > A.tol * A.tol * bb)) break;
        double* t = pold; pold = pnew; pnew = t;
        continue;
      }
      ++it;
      if (!(rr > A.tol * A.tol * bb)) break;
      beta = rz_new / rz;
      rz = rz_new;
      double* t = pold; pold = pnew; pnew = t;
    }
  }
  if (gtid == 0) {
    A.info[0] = (double)it;
    A.info[1] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  }
}
