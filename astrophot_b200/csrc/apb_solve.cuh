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
//   phase 1  q = A p  with  p = z + beta p_old  formed on the fly (the warp that owns a row also stores
//            p), and the CTA's share of p.q;
//   phase 2  alpha from the summed shares;  x += alpha p,  r -= alpha q,  z = M^-1 r  per row block, and
//            the CTA's shares of r.z and r.r.
// Dot products are never a pass over the vectors: every CTA writes its share, and after the barrier
// every CTA adds the shares in the same order, so alpha, beta and the stopping decision are
// bit-identical in all CTAs without a broadcast.  Rows whose product is split over several warps
// (the sky row couples to every model) write per-warp partial rows that their owner adds in order:
// no atomics, run-to-run deterministic.
#pragma once
#include <cooperative_groups.h>
#include "apb_internal.cuh"
#include "apb_image.cuh"

namespace cg = cooperative_groups;
#ifndef PCG_SPIN_NS
#define PCG_SPIN_NS 200
#endif

struct PcgRow {      // row block: free parameters [p0, p0+n) of owner `src`
  int src, p0, n;
  int item0, nitem;  // its work items (nitem > 1: product split over several warps)
  long long doff;    // offset of its diagonal block (n x n, row-major) in bvals
  int slot0, _pad;   // act_off[src] + p0: where its parameter indices start in act_slot (one dependent load less)
};
struct PcgEntry {    // one block contributing to a row block
  long long off;     // offset of the block in bvals (row-major [i of a][j of b], leading dimension ld)
  int ld;
  int transposed;    // 1: the row block is the block's b side
  int n;             // parameters of the other side
  int sl[NB_MAX];    // their indices in x (held here: one dependent load less per product)
};
struct PcgItem {     // work item (one warp): entries [e0, e1) of row block rb
  int rb, e0, e1;
  int multi;         // 0: the only item of its row; 1: first, 2: further item of a split row
  int n, slot0;      // copies of the row's size and slot start: the product does not wait for the row record
};

struct PcgArgs {
  const PcgRow* rows; int n_rows;
  const PcgEntry* entries;
  const PcgItem* items; int n_items;
  const int* multi_rows; int n_multi;   // rows split over several items
  const int* act_slot; const int* act_off;
  const double* bvals;     // tightly packed blocks
  const double* diagH;     // P: diagonal of H
  double* fac;             // n_rows x 64: Cholesky factors of the damped diagonal blocks
  const double* b;         // right-hand side (P)
  const double* x0;        // starting point (P) or NULL = 0 (the previous lambda-trial's solution is a good one)
  double* x;               // solution (P)
  double *r, *z, *pa, *pb, *q;   // work vectors (P)
  double* qpart;           // n_items x 8: partial rows of split rows
  double* part;            // gridDim x 4: per-CTA shares of the dot products
  unsigned int* barrier;   // arrival counter of pcg_barrier, zero at launch
  double* info;            // {iterations, final |r|/|b|}
  int P, max_iter;
  double L, tol;
};

// Grid barrier for the (cooperatively launched, hence co-resident) CTAs of k_pcg.  cooperative_groups' grid.sync()
// invalidates the whole L1 at every barrier, which sends every load of the solver's read-only tables (work items, row
// blocks, block values: the same ones every iteration) back to L2 -- measured: the phases were chains of ~7 dependent
// L2 round trips.  Here the vectors that change hands between CTAs are read with ld.cg (L2) and written through, the
// barrier is a monotonic arrival counter (zeroed before the launch), and the tables stay in L1.
__device__ __forceinline__ void pcg_barrier(unsigned int* counter, unsigned int& goal) {
  __syncthreads();
  if (threadIdx.x == 0) {
    goal += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    // (back off between polls: a few hundred CTAs spinning on one L2 line delay the arrivals they are waiting for)
    while (*(volatile unsigned int*)counter < goal) __nanosleep(PCG_SPIN_NS);
    __threadfence();
  }
  __syncthreads();
}

// sum of one column of the per-CTA shares, same order in every CTA
__device__ __forceinline__ double pcg_total(const double* part, int col, int ncta, double* sh) {
  double v = 0.0;
  for (int k = threadIdx.x; k < ncta; k += 256) v += __ldcg(part + 4 * k + col);
  double r = block_sum<256>(v, sh);
  __shared__ double bc;
  if (threadIdx.x == 0) bc = r;
  __syncthreads();
  r = bc;
  __syncthreads();
  return r;
}

// two columns of the per-CTA shares in one pass (r.z and r.r are always wanted together)
__device__ __forceinline__ void pcg_total2(const double* part, int col_a, int col_b, int ncta, double* sh, double& ta, double& tb) {
  double va = 0.0, vb = 0.0;
  for (int k = threadIdx.x; k < ncta; k += 256) {
    va += __ldcg(part + 4 * k + col_a);
    vb += __ldcg(part + 4 * k + col_b);
  }
  __shared__ double bc[2];
  const double ra = block_sum<256>(va, sh);
  if (threadIdx.x == 0) bc[0] = ra;
  const double rb = block_sum<256>(vb, sh);
  if (threadIdx.x == 0) bc[1] = rb;
  __syncthreads();
  ta = bc[0];
  tb = bc[1];
  __syncthreads();
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
    y[i] = v / Fr[i * (i + 1) / 2 + i];
  }
  double dot = 0.0;
#pragma unroll
  for (int i = NB_MAX - 1; i >= 0; --i) {
    double v = y[i];
#pragma unroll
    for (int k = i + 1; k < NB_MAX; ++k) v -= Fr[k * (k + 1) / 2 + i] * y[k];
    y[i] = v / Fr[i * (i + 1) / 2 + i];
    if (i < n) {
      z[slr[i]] = y[i];
      dot = fma(rv[i], y[i], dot);
    }
  }
  return dot;
}

#ifndef PCG_MINB
#define PCG_MINB 2
#endif
__global__ void __launch_bounds__(256, PCG_MINB) k_pcg(PcgArgs A) {
  unsigned int goal = 0;
  __shared__ double sh[8];
  __shared__ double bc2;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const int gwarp = gtid >> 5, nwarp = gsz >> 5;
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
    for (int i = 0; i < row.n; ++i)
      for (int j = 0; j <= i; ++j) F[i * 8 + j] = M[i][j];
    const int* sl = A.act_slot + row.slot0;
    double rv[NB_MAX];
    for (int i = 0; i < row.n; ++i) {
      const double bi = A.b[sl[i]];
      rv[i] = bi;
      A.x[sl[i]] = pre ? A.x0[sl[i]] : 0.0;
      A.r[sl[i]] = bi;
      A.pa[sl[i]] = 0.0;
      s_bb = fma(bi, bi, s_bb);
    }
    if (pre) {
      for (int i = 0; i < row.n; ++i) A.z[sl[i]] = A.x0[sl[i]];
    } else {
      s_rz += pcg_precond(A, rb, row, rv, A.z);
    }
  }
  {
    const double t0 = block_sum<256>(s_rz, sh), t1 = block_sum<256>(s_bb, sh);
    if (threadIdx.x == 0) { A.part[4 * blockIdx.x + 1] = t0; A.part[4 * blockIdx.x + 2] = t1; }
  }
  pcg_barrier(A.barrier, goal);
  double rz, bb_;
  pcg_total2(A.part, 1, 2, ncta, sh, rz, bb_);
  const double bb = bb_;
  double rr = bb, beta = 0.0;
  double* pold = A.pa;
  double* pnew = A.pb;
  int it = 0;
  if (bb > 0.0) {
    for (; it < A.max_iter;) {
      // ---- phase 1: p = z + beta p_old (on the fly), q = A p, share of p.q.  One warp per work item.
      double s_pq = 0.0;
      for (int k = gwarp; k < A.n_items; k += nwarp) {
        const PcgItem w = A.items[k];
        struct { int n; } row = {w.n};
        // A lane per block: all blocks of a row (up to 32 per sweep) are fetched at once -- the products are tiny, what
        // the phase costs is its chain of dependent L2 loads (item -> row -> entry -> values), so that chain is walked
        // once per sweep, not once per block.  Lane l accumulates its blocks' contribution to all row elements; a
        // shuffle tree adds the lanes (fixed order).
        double racc[NB_MAX];
#pragma unroll
        for (int i = 0; i < NB_MAX; ++i) racc[i] = 0.0;
        for (int e = w.e0 + lane; e < w.e1; e += 32) {
          const PcgEntry en = A.entries[e];
          const double* V = A.bvals + en.off;
          double pj[NB_MAX];
#pragma unroll
          for (int j = 0; j < NB_MAX; ++j) {
            const bool on = j < en.n;
            const int sj = en.sl[on ? j : 0];
            pj[j] = on ? fma(beta, __ldcg(pold + sj), __ldcg(A.z + sj)) : 0.0;
          }
#pragma unroll
          for (int i = 0; i < NB_MAX; ++i) {
            if (i < row.n) {
              double a = 0.0;
#pragma unroll
              for (int j = 0; j < NB_MAX; ++j)
                if (j < en.n) a = fma(en.transposed ? V[j * en.ld + i] : V[i * en.ld + j], pj[j], a);
              racc[i] += a;
            }
          }
        }
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < NB_MAX; ++i) {
          double v = racc[i];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == i) acc = v;
        }
        if (lane < row.n) {
          const int sl = A.act_slot[w.slot0 + lane];
          double v = acc * inv1L;
          if (w.multi < 2) {   // the row's first (or only) item: the damped diagonal, and it stores p
            const double pi = fma(beta, __ldcg(pold + sl), __ldcg(A.z + sl));
            const double d = A.diagH[sl];
            v += (d + A.L * (1.0 + d) - d * inv1L) * pi;
            pnew[sl] = pi;
            if (!w.multi) {
              A.q[sl] = v;
              s_pq = fma(pi, v, s_pq);
            }
          }
          if (w.multi) A.qpart[(long long)k * 8 + lane] = v;
        }
      }
      {
        const double t0 = block_sum<256>(s_pq, sh);
        if (threadIdx.x == 0) A.part[4 * blockIdx.x + 0] = t0;
      }
      pcg_barrier(A.barrier, goal);
      // ---- alpha.  Split rows: every CTA adds their partial rows (same order everywhere), stores q and adds p.q
      double pq = pcg_total(A.part, 0, ncta, sh);
      {
        double extra = 0.0;
        for (int m = 0; m < A.n_multi; ++m) {
          const PcgRow row = A.rows[A.multi_rows[m]];
          const int* sl = A.act_slot + row.slot0;
          for (int i = 0; i < row.n; ++i) {
            double v = 0.0;
            for (int t = threadIdx.x; t < row.nitem; t += 256) v += __ldcg(A.qpart + (long long)(row.item0 + t) * 8 + i);
            const double qi = block_sum<256>(v, sh);
            if (threadIdx.x == 0) {
              A.q[sl[i]] = qi;   // every CTA stores the same value; the row's owner reads the copy its own CTA wrote
              extra = fma(__ldcg(pnew + sl[i]), qi, extra);
            }
          }
        }
        if (threadIdx.x == 0) bc2 = extra;
        __syncthreads();
        pq += bc2;
        __syncthreads();
      }
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
          pv[i] = (on && !pre) ? __ldcg(pnew + ss[i]) : 0.0;
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
        s_rz2 += pcg_precond(A, rb, row, rv, A.z);
      }
      {
        const double t0 = block_sum<256>(s_rz2, sh), t1 = block_sum<256>(s_rr, sh);
        if (threadIdx.x == 0) { A.part[4 * blockIdx.x + 1] = t0; A.part[4 * blockIdx.x + 2] = t1; }
      }
      pcg_barrier(A.barrier, goal);
      double rz_new;
      pcg_total2(A.part, 1, 2, ncta, sh, rz_new, rr);
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
