// Dense damped solve for mid-size systems (fit/lm.py:359-371 with 159 < P of a few thousand and a parameter shared
// between sources -- joint multi-band fits, auxiliary PSF models -- where neither the single-CTA solver nor the
// block-sparse PCG applies).  The reference calls torch.linalg.solve (LU); the damped matrix
//     A = H o (I + (1 - I) / (1 + L)) + L I (1 + diag H)
// is symmetric positive definite for L > 0, so it is Cholesky-factored once per (H, L) and the factor serves both solves
// of a lambda-trial (h and the geodesic correction).
//
// k_chol_factor: one persistent cooperative kernel, right-looking blocked Cholesky on 32 x 32 tiles of a workspace copy
//   in global memory (it lives in L2: P = 2000 is 32 MB).  Per block column: every CTA factors the diagonal tile itself
//   in shared memory (cheaper than a third grid barrier), the panel tiles below it are dealt to the CTAs (triangular
//   solve, eight threads per row), grid barrier, the trailing tiles are dealt to the CTAs (rank-32 update), grid barrier.
//   Measured on B200 (profiles/r02_summary.md section 7): 161 / 871 / 2269 us at P = 200 / 1000 / 2000 -- the latency of a
//   block column (~25 us), not the barriers, bandwidth or the FP64 pipe.
// k_chol_solve: forward and backward substitution by one CTA of 1024 threads, row-wise (coalesced) reads of the factor.
#pragma once
#include <cuda_runtime.h>

#define CH_NB 32
#define CH_NT 256

struct CholArgs {
  const double* H;
  double L;
  int P;
  double* W;           // P x P workspace: the factor (lower triangle, row-major) on return
  int* info;           // 0 ok, 1 a pivot was not positive (not finite)
  unsigned int* bar;   // grid barrier counter, zero at launch
  int relaxed;         // poll the counter with relaxed loads (everything read across CTAs goes through ld.cg anyway)
};

// all CTAs resident (cooperative launch).  Release on arrival, acquire on leaving: data written to W by other CTAs before
// the barrier is read after it (with ld.cg: no line of W may be served from this SM's L1).
__device__ __forceinline__ void chol_barrier(unsigned int* counter, unsigned int& goal, int relaxed) {
  __syncthreads();
  if (threadIdx.x == 0) {
    goal += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
    unsigned int seen;
    if (relaxed) {
      // (as apb_solve.cuh's pcg_barrier: no acquire, hence no invalidation of L1 -- no line of W is ever served from L1)
      for (;;) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        if (seen >= goal) break;
      }
    } else {
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        if (seen >= goal) break;
        __nanosleep(32);
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(CH_NT) k_chol_factor(CholArgs A) {
  __shared__ double sD[CH_NB][CH_NB + 1];   // diagonal tile -> its factor
  __shared__ double sA[CH_NB][CH_NB + 1];
  __shared__ double sB[CH_NB][CH_NB + 1];
  __shared__ double sdiag[CH_NB], sdinv[CH_NB];
  __shared__ int s_bad;
  const int P = A.P, nb = (P + CH_NB - 1) / CH_NB;
  const int tid = threadIdx.x;
  double* __restrict__ W = A.W;
  unsigned int goal = 0;
  if (tid == 0) s_bad = 0;
  // the damped matrix (lm.py:359-371)
  const double L = A.L, off = 1.0 / (1.0 + L);
  for (long long q = (long long)blockIdx.x * CH_NT + tid; q < (long long)P * P; q += (long long)gridDim.x * CH_NT) {
    const int i = (int)(q / P), j = (int)(q - (long long)i * P);
    const double hij = A.H[q];
    W[q] = (i == j) ? hij + L * (1.0 + hij) : hij * off;
  }
  chol_barrier(A.bar, goal, A.relaxed);
  for (int kb = 0; kb < nb; ++kb) {
    const int k0 = kb * CH_NB, kn = min(CH_NB, P - k0);
    // ---- diagonal tile, factored by every CTA for itself (identity beyond the matrix)
    for (int q = tid; q < CH_NB * CH_NB; q += CH_NT) {
      const int r = q / CH_NB, c = q - r * CH_NB;
      sD[r][c] = (r < kn && c < kn) ? __ldcg(W + (long long)(k0 + r) * P + (k0 + c)) : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    // (one block barrier per column: the trailing entries are updated with the unscaled column, a_rc a_cc / d_c, and the
    //  columns are scaled by 1 / sqrt(d_c) at the end)
    // (a general fp64 division is a chain of ~50 instructions and sat on the critical path of every column: the pivot's
    //  reciprocal is __drcp_rn, every other division a multiplication by a reciprocal formed once per tile)
    for (int c = 0; c < kn; ++c) {
      const double d = sD[c][c];
      if (tid == 0 && (!(d > 0.0) || !(d < 1.7e308))) s_bad = 1;
      const double dinv = __drcp_rn(d);
#pragma unroll
      for (int q = tid; q < CH_NB * CH_NB; q += CH_NT) {
        const int r = q >> 5, cc = q & 31;
        if (cc > c && r >= cc && r < kn) sD[r][cc] -= sD[r][c] * sD[cc][c] * dinv;
      }
      __syncthreads();
    }
    if (tid < CH_NB) {
      const double sq = sqrt(sD[tid][tid]);
      sdiag[tid] = sq;
      sdinv[tid] = 1.0 / sq;
    }
    __syncthreads();
    for (int q = tid; q < CH_NB * CH_NB; q += CH_NT) {
      const int r = q >> 5, c = q & 31;
      if (r > c) sD[r][c] = sD[r][c] * sdinv[c];
      else if (r == c) sD[r][c] = sdiag[c];
    }
    __syncthreads();
    // ---- panel: X L_kk^T = A_ik for the row tiles below.  Eight threads per row (four rows per warp): column c of X
    //      needs the columns before it, x_rc = (a_rc - sum_{m<c} x_rm l_cm) / l_cc -- the eight share the sum, a warp
    //      barrier per column keeps the row's new entry visible to them
    for (int ib = kb + 1 + blockIdx.x; ib < nb; ib += gridDim.x) {
      const int i0 = ib * CH_NB, in = min(CH_NB, P - i0);
      __syncthreads();
      for (int q = tid; q < CH_NB * CH_NB; q += CH_NT) {
        const int r = q / CH_NB, c = q - r * CH_NB;
        sA[r][c] = (r < in && c < kn) ? __ldcg(W + (long long)(i0 + r) * P + (k0 + c)) : 0.0;
      }
      __syncthreads();
      {
        const int r = tid >> 3, sub = tid & 7;
        for (int c = 0; c < kn; ++c) {
          double part = 0.0;
          for (int m = sub; m < c; m += 8) part = fma(sA[r][m], sD[c][m], part);
          part += __shfl_xor_sync(0xffffffffu, part, 4);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          if (sub == 0) sA[r][c] = (sA[r][c] - part) * sdinv[c];
          __syncwarp();
        }
      }
      __syncthreads();
      for (int q = tid; q < CH_NB * CH_NB; q += CH_NT) {
        const int r = q / CH_NB, c = q - r * CH_NB;
        if (r < in && c < kn) W[(long long)(i0 + r) * P + (k0 + c)] = sA[r][c];
      }
    }
    chol_barrier(A.bar, goal, A.relaxed);
    // (the factored diagonal tile goes back only now: before the barrier a slower CTA may still be loading the tile)
    if (blockIdx.x == 0)
      for (int q = tid; q < kn * kn; q += CH_NT) {
        const int r = q / kn, c = q - r * kn;
        if (c <= r) W[(long long)(k0 + r) * P + (k0 + c)] = sD[r][c];
      }
    // ---- trailing update: A_ij -= A_ik A_jk^T for the tiles kb < jb <= ib, dealt to the CTAs
    const int m = nb - kb - 1;
    const int T = m * (m + 1) / 2;
    for (int t = blockIdx.x; t < T; t += gridDim.x) {
      int ir = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
      while ((ir + 1) * (ir + 2) / 2 <= t) ++ir;
      while (ir * (ir + 1) / 2 > t) --ir;
      const int jr = t - ir * (ir + 1) / 2;
      const int i0 = (kb + 1 + ir) * CH_NB, j0 = (kb + 1 + jr) * CH_NB;
      const int in = min(CH_NB, P - i0), jn = min(CH_NB, P - j0);
      __syncthreads();
      for (int q = tid; q < CH_NB * CH_NB; q += CH_NT) {
        const int r = q / CH_NB, c = q - r * CH_NB;
        sA[r][c] = (r < in && c < kn) ? __ldcg(W + (long long)(i0 + r) * P + (k0 + c)) : 0.0;
        sB[r][c] = (r < jn && c < kn) ? __ldcg(W + (long long)(j0 + r) * P + (k0 + c)) : 0.0;
      }
      __syncthreads();
      {
        // thread (r0, c): rows r0, r0 + 8, r0 + 16, r0 + 24 of column c = lane (sB rows are 33 doubles apart: no bank
        // conflict; sA reads are broadcasts); the four old values are on their way while the products are formed
        const int r0 = tid >> 5, c = tid & 31;
        double old[4], acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = r0 + 8 * i;
          old[i] = (r < in && c < jn) ? __ldcg(W + (long long)(i0 + r) * P + (j0 + c)) : 0.0;
        }
#pragma unroll 8
        for (int k = 0; k < CH_NB; ++k) {
          const double b = sB[c][k];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fma(sA[r0 + 8 * i][k], b, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = r0 + 8 * i;
          if (r < in && c < jn) W[(long long)(i0 + r) * P + (j0 + c)] = old[i] - acc[i];
        }
      }
    }
    chol_barrier(A.bar, goal, A.relaxed);
  }
  if (blockIdx.x == 0 && tid == 0) *A.info = s_bad;
}

// L y = rhs, L^T x = y with the factor of k_chol_factor (lower triangle of W, row-major).  One CTA (CH_SOLVE_NT threads:
// the substitution is a chain of P / 32 dependent steps, each a strip of the factor read once -- what matters is how many
// loads the one SM keeps in flight).  x doubles as the work vector; rhs may alias x.
#define CH_SOLVE_NT 1024
__global__ void __launch_bounds__(CH_SOLVE_NT) k_chol_solve(const double* __restrict__ W, const double* rhs, int P, double* x) {
  __shared__ double sT[CH_NB][CH_NB + 1];
  __shared__ double sy[CH_NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = (P + CH_NB - 1) / CH_NB;
  for (int kb = 0; kb < nb; ++kb) {
    const int k0 = kb * CH_NB, kn = min(CH_NB, P - k0);
    if (warp < kn) {       // a warp per row of the block: rhs_r - sum_{j < k0} L_rj y_j
      const int r = warp;
      const double* row = W + (long long)(k0 + r) * P;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int j = lane;
      for (; j + 96 < k0; j += 128) {
        a0 = fma(row[j], x[j], a0);
        a1 = fma(row[j + 32], x[j + 32], a1);
        a2 = fma(row[j + 64], x[j + 64], a2);
        a3 = fma(row[j + 96], x[j + 96], a3);
      }
      for (; j < k0; j += 32) a0 = fma(row[j], x[j], a0);
      double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) sy[r] = rhs[k0 + r] - acc;
    }
    for (int q = tid; q < CH_NB * CH_NB; q += CH_SOLVE_NT) {
      const int r = q / CH_NB, c = q - r * CH_NB;
      sT[r][c] = (r < kn && c <= r) ? W[(long long)(k0 + r) * P + (k0 + c)] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    if (warp == 0) {
      double v = lane < kn ? sy[lane] : 0.0;
      const double rinv = 1.0 / sT[lane][lane];      // (identity beyond the matrix)
      for (int c = 0; c < kn; ++c) {
        const double yc = __shfl_sync(0xffffffffu, v * rinv, c);
        if (lane == c) v = yc;
        else if (lane > c) v -= sT[lane][c] * yc;
      }
      if (lane < kn) x[k0 + lane] = v;
    }
    __syncthreads();
  }
  for (int kb = nb - 1; kb >= 0; --kb) {
    const int k0 = kb * CH_NB, kn = min(CH_NB, P - k0);
    for (int q = tid; q < CH_NB * CH_NB; q += CH_SOLVE_NT) {
      const int r = q / CH_NB, c = q - r * CH_NB;
      sT[r][c] = (r < kn && c <= r) ? W[(long long)(k0 + r) * P + (k0 + c)] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    if (warp == 0) {
      double v = lane < kn ? x[k0 + lane] : 0.0;
      const double rinv = 1.0 / sT[lane][lane];
      for (int c = kn - 1; c >= 0; --c) {
        const double xc = __shfl_sync(0xffffffffu, v * rinv, c);
        if (lane == c) v = xc;
        else if (lane < c) v -= sT[c][lane] * xc;
      }
      if (lane < kn) x[k0 + lane] = v;
      sy[lane] = lane < kn ? v : 0.0;
    }
    __syncthreads();
    // y_j -= sum_r L_(k0+r, j) x_(k0+r) for the columns before the block: a thread per column, rows read coalesced
    for (int j = tid; j < k0; j += CH_SOLVE_NT) {
      const double* col = W + (long long)k0 * P + j;
      double a0 = 0.0, a1 = 0.0;
      int r = 0;
      for (; r + 1 < kn; r += 2) {
        a0 = fma(col[(long long)r * P], sy[r], a0);
        a1 = fma(col[(long long)(r + 1) * P], sy[r + 1], a1);
      }
      if (r < kn) a0 = fma(col[(long long)r * P], sy[r], a0);
      x[j] -= a0 + a1;
    }
    __syncthreads();
  }
}
