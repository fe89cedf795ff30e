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
//   solve, a thread per row), grid barrier, the trailing tiles are dealt to the CTAs (rank-32 update), grid barrier.
// k_chol_solve: forward and backward substitution by one CTA, row-wise (coalesced) reads of the factor.
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
};

// all CTAs resident (cooperative launch).  Release on arrival, acquire on leaving: data written to W by other CTAs before
// the barrier is read after it (with ld.cg: no line of W may be served from this SM's L1).
__device__ __forceinline__ void chol_barrier(unsigned int* counter, unsigned int& goal) {
  __syncthreads();
  if (threadIdx.x == 0) {
    goal += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
    unsigned int seen;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (seen >= goal) break;
      __nanosleep(64);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(CH_NT) k_chol_factor(CholArgs A) {
  __shared__ double sD[CH_NB][CH_NB + 1];   // diagonal tile -> its factor
  __shared__ double sA[CH_NB][CH_NB + 1];
  __shared__ double sB[CH_NB][CH_NB + 1];
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
  chol_barrier(A.bar, goal);
  for (int kb = 0; kb < nb; ++kb) {
    const int k0 = kb * CH_NB, kn = min(CH_NB, P - k0);
    // ---- diagonal tile, factored by every CTA for itself (identity beyond the matrix)
    for (int q = tid; q < CH_NB * CH_NB; q += CH_NT) {
      const int r = q / CH_NB, c = q - r * CH_NB;
      sD[r][c] = (r < kn && c < kn) ? __ldcg(W + (long long)(k0 + r) * P + (k0 + c)) : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int c = 0; c < kn; ++c) {
      if (tid == 0) {
        double d = sD[c][c];
        if (!(d > 0.0) || !(d < 1.7e308)) { s_bad = 1; d = 1.0; }
        sD[c][c] = sqrt(d);
      }
      __syncthreads();
      const double dinv = 1.0 / sD[c][c];
      if (tid > c && tid < kn) sD[tid][c] *= dinv;
      __syncthreads();
      for (int q = tid; q < kn * kn; q += CH_NT) {
        const int r = q / kn, cc = q - r * kn;
        if (cc > c && r >= cc) sD[r][cc] -= sD[r][c] * sD[cc][c];
      }
      __syncthreads();
    }
    // ---- panel: X L_kk^T = A_ik for the row tiles below, a thread per row
    for (int ib = kb + 1 + blockIdx.x; ib < nb; ib += gridDim.x) {
      const int i0 = ib * CH_NB, in = min(CH_NB, P - i0);
      __syncthreads();
      for (int q = tid; q < CH_NB * CH_NB; q += CH_NT) {
        const int r = q / CH_NB, c = q - r * CH_NB;
        sA[r][c] = (r < in && c < kn) ? __ldcg(W + (long long)(i0 + r) * P + (k0 + c)) : 0.0;
      }
      __syncthreads();
      if (tid < in) {
        for (int c = 0; c < kn; ++c) {
          double v = sA[tid][c];
          for (int m = 0; m < c; ++m) v -= sA[tid][m] * sD[c][m];
          sA[tid][c] = v / sD[c][c];
        }
      }
      __syncthreads();
      for (int q = tid; q < CH_NB * CH_NB; q += CH_NT) {
        const int r = q / CH_NB, c = q - r * CH_NB;
        if (r < in && c < kn) W[(long long)(i0 + r) * P + (k0 + c)] = sA[r][c];
      }
    }
    chol_barrier(A.bar, goal);
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
      for (int q = tid; q < CH_NB * CH_NB; q += CH_NT) {
        const int r = q / CH_NB, c = q - r * CH_NB;      // c = lane: sB rows are 33 doubles apart, no bank conflict
        if (r < in && c < jn) {
          double acc = 0.0;
#pragma unroll 8
          for (int k = 0; k < CH_NB; ++k) acc = fma(sA[r][k], sB[c][k], acc);
          double* w = W + (long long)(i0 + r) * P + (j0 + c);
          *w = __ldcg(w) - acc;
        }
      }
    }
    chol_barrier(A.bar, goal);
  }
  if (blockIdx.x == 0 && tid == 0) *A.info = s_bad;
}

// L y = rhs, L^T x = y with the factor of k_chol_factor (lower triangle of W, row-major).  One CTA of 256 threads.
// x doubles as the work vector; rhs may alias x.
__global__ void __launch_bounds__(CH_NT) k_chol_solve(const double* __restrict__ W, const double* rhs, int P, double* x) {
  __shared__ double sT[CH_NB][CH_NB + 1];
  __shared__ double sy[CH_NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = (P + CH_NB - 1) / CH_NB;
  for (int kb = 0; kb < nb; ++kb) {
    const int k0 = kb * CH_NB, kn = min(CH_NB, P - k0);
    for (int r = warp; r < kn; r += CH_NT / 32) {
      const double* row = W + (long long)(k0 + r) * P;
      double acc = 0.0;
      for (int j = lane; j < k0; j += 32) acc = fma(row[j], x[j], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) sy[r] = rhs[k0 + r] - acc;
    }
    for (int q = tid; q < CH_NB * CH_NB; q += CH_NT) {
      const int r = q / CH_NB, c = q - r * CH_NB;
      sT[r][c] = (r < kn && c <= r) ? W[(long long)(k0 + r) * P + (k0 + c)] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    if (warp == 0) {
      double v = lane < kn ? sy[lane] : 0.0;
      for (int c = 0; c < kn; ++c) {
        const double yc = __shfl_sync(0xffffffffu, v, c) / sT[c][c];
        if (lane == c) v = yc;
        else if (lane > c) v -= sT[lane][c] * yc;
      }
      if (lane < kn) x[k0 + lane] = v;
    }
    __syncthreads();
  }
  for (int kb = nb - 1; kb >= 0; --kb) {
    const int k0 = kb * CH_NB, kn = min(CH_NB, P - k0);
    for (int q = tid; q < CH_NB * CH_NB; q += CH_NT) {
      const int r = q / CH_NB, c = q - r * CH_NB;
      sT[r][c] = (r < kn && c <= r) ? W[(long long)(k0 + r) * P + (k0 + c)] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    if (warp == 0) {
      double v = lane < kn ? x[k0 + lane] : 0.0;
      for (int c = kn - 1; c >= 0; --c) {
        const double xc = __shfl_sync(0xffffffffu, v, c) / sT[c][c];
        if (lane == c) v = xc;
        else if (lane < c) v -= sT[c][lane] * xc;
      }
      if (lane < kn) x[k0 + lane] = v;
      if (lane < CH_NB) sy[lane] = lane < kn ? v : 0.0;
    }
    __syncthreads();
    for (int j = tid; j < k0; j += CH_NT) {
      double acc = 0.0;
      for (int r = 0; r < kn; ++r) acc = fma(W[(long long)(k0 + r) * P + j], sy[r], acc);
      x[j] -= acc;
    }
    __syncthreads();
  }
}
