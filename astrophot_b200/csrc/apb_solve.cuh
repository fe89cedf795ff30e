// Damped LM solve for large parameter counts (crowded fields): block-sparse preconditioned
// conjugate gradients, one persistent cooperative kernel per solve.
//
// The reference forms the dense P x P matrix  A = H o (I + (1-I)/(1+L)) + L I (1 + diag H)  and calls
// torch.linalg.solve (fit/lm.py:359-371).  For a crowded field H = J^T W J is block-sparse: a source
// only couples to the sources its window overlaps (and to the sky), so the matrix is kept as the
// list of <=8x8 blocks k_blocks already produces, and A p is a block-sparse product (a few MB
// instead of P^2 doubles per product).  A is symmetric positive definite for L > 0, the
// preconditioner is block-Jacobi (Cholesky of the damped diagonal block of every source), which
// removes the ill-conditioning inside a source (Sersic n / Re / Ie); what is left -- overlaps and the
// sky row -- converges in a few dozen iterations.
//
// One launch runs the whole iteration: grid = resident CTAs, three grid barriers per iteration,
// every CTA evaluates the (small) dot products redundantly in a fixed order, so alpha / beta and the
// stopping decision are bit-identical in all CTAs without a broadcast.
#pragma once
#include <cooperative_groups.h>
#include "apb_internal.cuh"
#include "apb_image.cuh"

namespace cg = cooperative_groups;

struct PcgRow {      // row block: planes [p0, p0+n) of source `src`
  int src, p0, n, diag_block;   // diag_block: id of its diagonal block
};
struct PcgEntry {    // one block contributing to a row block
  int block;         // block id (values at bvals + 64*block, row-major [i of a][j of b])
  int transposed;    // 1: the row block is the block's b side
  int slot0;         // index into act_slot of the other side's first plane
  int n;             // planes of the other side
};
struct PcgItem {     // work item: entries [e0, e1) of row block rb
  int rb, e0, e1;
  int multi;         // 0: the only item of its row; 1: first, 2: further item of a row split over several
};

struct PcgArgs {
  const PcgRow* rows; int n_rows;
  const PcgEntry* entries;
  const PcgItem* items; int n_items;
  const int* act_slot; const int* act_off;
  const double* bvals;     // n_blocks x 64
  const double* diagH;     // P: diagonal of H
  double* fac;             // n_rows x 64: Cholesky factors of the damped diagonal blocks
  const double* b;         // right-hand side (P)
  double* x;               // solution (P)
  double *r, *z, *p, *q;   // work vectors (P)
  double* info;            // {iterations, final |r|/|b|, 0, 0}
  int P, max_iter;
  double L, tol;
};

// full dot product by one CTA, fixed order (identical in every CTA)
__device__ __forceinline__ double pcg_dot(const double* __restrict__ a, const double* __restrict__ b, int n, double* sh) {
  double v0 = 0.0, v1 = 0.0;
  int i = threadIdx.x;
  for (; i + 256 < n; i += 512) {
    v0 = fma(__ldcg(a + i), __ldcg(b + i), v0);
    v1 = fma(__ldcg(a + i + 256), __ldcg(b + i + 256), v1);
  }
  for (; i < n; i += 256) v0 = fma(__ldcg(a + i), __ldcg(b + i), v0);
  double r = block_sum<256>(v0 + v1, sh);
  __shared__ double bc;
  if (threadIdx.x == 0) bc = r;
  __syncthreads();
  r = bc;
  __syncthreads();
  return r;
}

// z = M^-1 r for the diagonal block of one row block (one thread): L L^T z = r
__device__ __forceinline__ void pcg_precond(const PcgArgs& A, int rb, const double* __restrict__ r, double* __restrict__ z) {
  const PcgRow row = A.rows[rb];
  const int* sl = A.act_slot + A.act_off[row.src] + row.p0;
  const double* F = A.fac + (long long)rb * 64;
  double y[NB_MAX];
  for (int i = 0; i < row.n; ++i) {
    double v = __ldcg(r + sl[i]);
    for (int k = 0; k < i; ++k) v -= F[i * 8 + k] * y[k];
    y[i] = v / F[i * 8 + i];
  }
  for (int i = row.n - 1; i >= 0; --i) {
    double v = y[i];
    for (int k = i + 1; k < row.n; ++k) v -= F[k * 8 + i] * y[k];
    y[i] = v / F[i * 8 + i];
    z[sl[i]] = y[i];
  }
}

// rows whose product is accumulated by several work items (the sky row) are zeroed one phase ahead
__device__ __forceinline__ void pcg_zero_multi(const PcgArgs& A, int gtid, int gsz) {
  for (int k = gtid; k < A.n_items; k += gsz) {
    const PcgItem w = A.items[k];
    if (w.multi == 1) {   // first item of a split row
      const PcgRow row = A.rows[w.rb];
      const int* sl = A.act_slot + A.act_off[row.src] + row.p0;
      for (int i = 0; i < row.n; ++i) A.q[sl[i]] = 0.0;
    }
  }
}

__global__ void __launch_bounds__(256) k_pcg(PcgArgs A) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[8];
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const int gwarp = gtid >> 5, nwarp = gsz >> 5;
  const double inv1L = 1.0 / (1.0 + A.L);

  // ---- setup: Cholesky of every damped diagonal block; x = 0, r = b, z = M^-1 r, p = z
  for (int rb = gtid; rb < A.n_rows; rb += gsz) {
    const PcgRow row = A.rows[rb];
    const double* V = A.bvals + (long long)row.diag_block * 64;
    double* F = A.fac + (long long)rb * 64;
    double M[NB_MAX][NB_MAX];
    for (int i = 0; i < row.n; ++i)
      for (int j = 0; j <= i; ++j) {
        const double h = V[i * 8 + j];
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
    const int* sl = A.act_slot + A.act_off[row.src] + row.p0;
    for (int i = 0; i < row.n; ++i) {
      A.x[sl[i]] = 0.0;
      A.r[sl[i]] = A.b[sl[i]];
    }
    pcg_precond(A, rb, A.b, A.z);
    for (int i = 0; i < row.n; ++i) A.p[sl[i]] = A.z[sl[i]];
  }
  pcg_zero_multi(A, gtid, gsz);
  grid.sync();
  double rz = pcg_dot(A.r, A.z, A.P, sh);
  const double bb = pcg_dot(A.b, A.b, A.P, sh);
  double rr = bb;
  int it = 0;
  if (bb > 0.0) {
    for (; it < A.max_iter; ++it) {
      // ---- q = A p: one warp per work item
      for (int k = gwarp; k < A.n_items; k += nwarp) {
        const PcgItem w = A.items[k];
        const PcgRow row = A.rows[w.rb];
        const int i = lane & 7, jg = lane >> 3;
        double acc = 0.0;
        for (int e = w.e0; e < w.e1; ++e) {
          const PcgEntry en = A.entries[e];
          const double* V = A.bvals + (long long)en.block * 64;
          const int* so = A.act_slot + en.slot0;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * jg + jj;
            if (j < en.n && i < row.n) acc = fma(en.transposed ? V[j * 8 + i] : V[i * 8 + j], __ldcg(A.p + so[j]), acc);
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (lane < row.n) {
          const int sl = A.act_slot[A.act_off[row.src] + row.p0 + lane];
          double v = acc * inv1L;
          if (w.multi < 2) {   // the damped diagonal, once per row
            const double d = A.diagH[sl];
            v += (d + A.L * (1.0 + d) - d * inv1L) * __ldcg(A.p + sl);
          }
          if (w.multi) atomicAdd(A.q + sl, v);
          else A.q[sl] = v;
        }
      }
      grid.sync();
      // ---- alpha; x += alpha p; r -= alpha q; z = M^-1 r
      const double pq = pcg_dot(A.p, A.q, A.P, sh);
      const double alpha = rz / pq;
      for (int rb = gtid; rb < A.n_rows; rb += gsz) {
        const PcgRow row = A.rows[rb];
        const int* sl = A.act_slot + A.act_off[row.src] + row.p0;
        for (int i = 0; i < row.n; ++i) {
          const int s = sl[i];
          A.x[s] = fma(alpha, A.p[s], A.x[s]);
          A.r[s] = fma(-alpha, __ldcg(A.q + s), A.r[s]);
        }
        pcg_precond(A, rb, A.r, A.z);
      }
      grid.sync();
      // ---- beta; p = z + beta p
      const double rz_new = pcg_dot(A.r, A.z, A.P, sh);
      rr = pcg_dot(A.r, A.r, A.P, sh);
      if (!(rr > A.tol * A.tol * bb) || !(pq > 0.0)) { ++it; break; }
      const double beta = rz_new / rz;
      rz = rz_new;
      for (int rb = gtid; rb < A.n_rows; rb += gsz) {
        const PcgRow row = A.rows[rb];
        const int* sl = A.act_slot + A.act_off[row.src] + row.p0;
        for (int i = 0; i < row.n; ++i) A.p[sl[i]] = fma(beta, A.p[sl[i]], A.z[sl[i]]);
      }
      pcg_zero_multi(A, gtid, gsz);
      grid.sync();
    }
  }
  if (gtid == 0) {
    A.info[0] = (double)it;
    A.info[1] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  }
}
