// Table-driven fp64 exp / log and a Newton reciprocal for the profile kernels.
//
// The profile evaluation (utils/parametric_profiles.py:7-18 `sersic_torch`: pow + exp; moffat_torch: pow) is bound by the
// FP64 pipe, and the CUDA math library's exp / log spend ~17 / ~30 FP64 instructions plus as many integer ones on cases
// that cannot occur here.  These versions are good to ~3e-16 relative (exp) and ~2e-16 absolute / relative (log), well
// inside the 1e-10 parity bar, in 10 FP64 instructions each:
//   exp(x) = 2^e . 2^(j/64) . exp(r),  k = rint(64 x / ln 2) = 64 e + j,  |r| <= ln2/128, degree-5 polynomial
//   log(x) = k ln 2 + log(c_i) + log1p(r),  z = x 2^-k in [0.6875, 1.375),  i = top 7 mantissa bits of z,
//            r = z / c_i - 1 with |r| < 2^-8, degree-6 polynomial (the interval that starts at z = 1 has c = 1 and
//            log c = 0, so log is continuous and exact to rounding around 1)
// Anything outside the fast range (|x| >= 708, non-finite, zero, negative, denormal) goes to the library function.
// The tables (2.5 KB) are computed on the host in long double at plan creation, live in global memory and are copied to
// shared memory by every kernel that evaluates profiles (apb_math_load).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

struct ApbMathTab {
  double lg[128][2];   // {1/c_i rounded, -log(that)}
  double ex[64];       // 2^(j/64)
};

#if defined(__CUDACC__)
#define APB_MHD __host__ __device__ __forceinline__
#else
#define APB_MHD inline
#endif

APB_MHD int apb_hi(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32);
#endif
}
APB_MHD int apb_lo(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int)(u & 0xffffffffu);
#endif
}
APB_MHD double apb_hilo(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x;
#endif
}

#define APB_EXP_FAST_HI 0x40862000   // |x| < 708

// Polynomial coefficients and reduction constants.  On the device they sit in the constant bank: a DFMA takes a
// c[bank][offset] operand for free, while a 64-bit literal costs two move instructions every time it is used.
#define APB_MC_LIST                                                                                             \
  { 0x1.71547652b82fep+6,      /* 0: 64 / ln 2 */                                                                \
    -0x1.62e42fef00000p-7,     /* 1: -(ln 2 / 64) high part, 20 trailing zero bits */                            \
    -0x1.473de6af278edp-40,    /* 2: -(ln 2 / 64) low part */                                                    \
    8.3333333333333332177e-03, /* 3: 1/120 */                                                                    \
    4.1666666666666664354e-02, /* 4: 1/24 */                                                                     \
    1.6666666666666665741e-01, /* 5: 1/6 */                                                                      \
    -1.6666666666666665741e-01,/* 6: -1/6 */                                                                     \
    0.2,                       /* 7 */                                                                           \
    3.3333333333333331483e-01, /* 8: 1/3 */                                                                      \
    0x1.62e42fefa39efp-1,      /* 9: ln 2 */                                                                     \
    4503601774854144.0,        /* 10: 2^52 + 2^31 */                                                             \
    6755399441055744.0 }       /* 11: 1.5 * 2^52 */
#if defined(__CUDACC__)
__constant__ double c_apb_mc[12] = APB_MC_LIST;
#endif
static const double h_apb_mc[12] = APB_MC_LIST;
#if defined(__CUDA_ARCH__)
#define APB_MC(i) c_apb_mc[i]
#else
#define APB_MC(i) h_apb_mc[i]
#endif

// out-of-range arguments: the library functions, kept out of line so that the hot path stays short
#if defined(__CUDACC__)
__host__ __device__ __noinline__ static double apb_exp_slow(double x) { return exp(x); }
__host__ __device__ __noinline__ static double apb_log_slow(double x) { return log(x); }
#else
static double apb_exp_slow(double x) { return exp(x); }
static double apb_log_slow(double x) { return log(x); }
#endif

// fast range of apb_exp_nc / apb_log_nc (outside it their result is garbage, not an error)
APB_MHD bool apb_exp_bad(double x) { return (unsigned)(apb_hi(x) & 0x7fffffff) >= (unsigned)APB_EXP_FAST_HI; }
APB_MHD bool apb_log_bad(double x) { return (unsigned)(apb_hi(x) - 0x00100000) >= 0x7fe00000u; }

// exp(x) for |x| < 708, no range check
APB_MHD double apb_exp_nc(double x, const ApbMathTab& T) {
  const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: adding it rounds to the nearest integer
  double kd = fma(x, APB_MC(0), MAGIC);      // x * 64 / ln 2
  const int ki = apb_lo(kd);
  kd -= MAGIC;
  // ln2/64 = HI + LO, HI with 20 trailing zero bits: kd * HI is exact for |kd| < 2^17
  double r = fma(kd, APB_MC(1), x);
  r = fma(kd, APB_MC(2), r);
  double q = fma(r, APB_MC(3), APB_MC(4));
  q = fma(q, r, APB_MC(5));
  q = fma(q, r, 0.5);
  const double r2 = r * r;
  const double p = fma(q, r2, r);
  const double t = T.ex[ki & 63];
  const double v = fma(t, p, t);
  return apb_hilo(apb_hi(v) + ((ki >> 6) << 20), apb_lo(v));
}

// log(x) for positive normal finite x, no range check
APB_MHD double apb_log_nc(double x, const ApbMathTab& T) {
  const int hx = apb_hi(x);
  const int tmp = hx - 0x3fe60000;
  const int i = (tmp >> 13) & 127;
  const int k = tmp >> 20;
  const double z = apb_hilo(hx - (tmp & (int)0xfff00000), apb_lo(x));
  const double invc = T.lg[i][0], logc = T.lg[i][1];
  const double r = fma(z, invc, -1.0);
  // (double)k without a conversion instruction: 2^52 + 2^31 + k as bits, minus the offset
  const double kd = apb_hilo(0x43300000, k ^ (int)0x80000000) - APB_MC(10);
  double q = fma(r, APB_MC(6), APB_MC(7));
  q = fma(q, r, -0.25);
  q = fma(q, r, APB_MC(8));
  q = fma(q, r, -0.5);
  const double r2 = r * r;
  const double p = fma(q, r2, r);
  return fma(kd, APB_MC(9), logc + p);
}

APB_MHD double apb_exp(double x, const ApbMathTab& T) {
  if (apb_exp_bad(x)) return apb_exp_slow(x);
  return apb_exp_nc(x, T);
}
APB_MHD double apb_log(double x, const ApbMathTab& T) {
  if (apb_log_bad(x)) return apb_log_slow(x);
  return apb_log_nc(x, T);
}

// 1 / b for normal b: hardware seed (20+ bits) and two Newton steps
APB_MHD double apb_rcp(double b) {
#if defined(__CUDA_ARCH__)
  double x0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(b));
  double e = fma(-b, x0, 1.0);
  double x1 = fma(x0, e, x0);
  e = fma(-b, x1, 1.0);
  x1 = fma(x1, e, x1);
  // the seed flushes results below the normal range and cannot represent 1/b for |b| < 2^-1022 or > 2^1022
  const unsigned hb = (unsigned)(__double2hiint(b) & 0x7fffffff);
  return (hb - 0x00200000u < 0x7fc00000u) ? x1 : 1.0 / b;
#else
  return 1.0 / b;
#endif
}

// host: fill the tables (long double; the pair {invc, logc} is consistent: logc = -log(invc as stored))
inline void apb_math_fill(ApbMathTab* T) {
  for (int i = 0; i < 128; ++i) {
    // interval i of z: bits 0x3fe60000 + (i << 13) in the high word
    const int hi0 = 0x3fe60000 + (i << 13), hi1 = 0x3fe60000 + ((i + 1) << 13);
    uint64_t u0 = (uint64_t)(uint32_t)hi0 << 32, u1 = (uint64_t)(uint32_t)hi1 << 32;
    double z0, z1;
    memcpy(&z0, &u0, 8);
    memcpy(&z1, &u1, 8);
    double invc, logc;
    if (z0 == 1.0) {
      invc = 1.0; logc = 0.0;     // r = z - 1 exactly: log is exact to rounding next to 1 from above
    } else if (z1 == 1.0) {
      // last interval below 1: anchor it at 1 too, so that r -> 0 as x -> 1- (|r| <= 2^-8 there)
      invc = 1.0; logc = 0.0;
    } else {
      const long double c = 0.5L * ((long double)z0 + (long double)z1);
      invc = (double)(1.0L / c);
      logc = (double)(-logl((long double)invc));
    }
    T->lg[i][0] = invc;
    T->lg[i][1] = logc;
  }
  for (int j = 0; j < 64; ++j) T->ex[j] = (double)powl(2.0L, (long double)j / 64.0L);
}

#if defined(__CUDACC__)
__device__ ApbMathTab g_mathtab;    // filled by apb_plan_create
__shared__ ApbMathTab s_mathtab;    // per CTA; a kernel that never touches it gets no allocation
// every thread of the CTA must call this before the first profile evaluation
__device__ __forceinline__ void apb_math_load() {
  const double* g = (const double*)&g_mathtab;
  double* s = (double*)&s_mathtab;
  for (int i = threadIdx.x; i < (int)(sizeof(ApbMathTab) / sizeof(double)); i += blockDim.x) s[i] = g[i];
  __syncthreads();
}
#endif
