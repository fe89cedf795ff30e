// FFT convolution for large PSFs (utils/operations.py:9-36 `fft_convolve_torch`, called from
// models/_model_methods.py:233-257).  Hand-written fp64 FFTs, three passes per plane:
//   k_fft_rows     real rows -> half spectra           (two real rows share one complex FFT)
//   k_fft_cols     column FFT . PSF spectrum . inverse column FFT, all in shared memory: the
//                  pointwise multiply never touches HBM and the full 2-D spectrum is never stored
//   k_fft_rows_inv half spectra -> real rows, cropped to the output window
// Each transform is a Stockham autosort FFT whose stages are large-radix (16, 9, 8, 5, 4, 3, 2)
// butterflies held in registers; shared memory is only the exchange between stages (ping-pong, one
// barrier per stage), so a 1152-point transform is three stages (16.8.9) and three exchanges.
// The transform lengths are a cheap 2^a 3^b 5^c >= the padded stamp, not the reference's exact
// image size: the valid region of the circular convolution equals the linear one either way
// (the reference relies on the same fact, model_object.py:313-349).
#pragma once
#include "apb_internal.cuh"

typedef double2 cpx;

#define APB_FFT_MAX_STAGE 14
struct FftDesc {
  int N, nstage;
  int radix[APB_FFT_MAX_STAGE];
  long long tw_off;  // offset (in cpx) of exp(-2 pi i k / N), k = 0..N-1, in the twiddle arena
};

#if defined(__CUDACC__)
#define APB_HD __host__ __device__ __forceinline__
#else
#define APB_HD inline
#endif

APB_HD cpx c_mul(cpx a, cpx b) { return cpx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
APB_HD cpx c_add(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
APB_HD cpx c_sub(cpx a, cpx b) { return cpx{a.x - b.x, a.y - b.y}; }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
APB_HD cpx c_rot(cpx a) { return INV ? cpx{-a.y, a.x} : cpx{a.y, -a.x}; }

// cos, sin of 2 pi m / R for the composite radices (first octant entries, 25 digits)
template <int R>
APB_HD constexpr double root_cos(int m) {
  if (R == 16) {
    constexpr double t[16] = {1.0, 0.9238795325112867561281832, 0.7071067811865475244008444, 0.38268343236508977172846,
                              0.0, -0.38268343236508977172846, -0.7071067811865475244008444, -0.9238795325112867561281832,
                              -1.0, -0.9238795325112867561281832, -0.7071067811865475244008444, -0.38268343236508977172846,
                              0.0, 0.38268343236508977172846, 0.7071067811865475244008444, 0.9238795325112867561281832};
    return t[m];
  } else if (R == 9) {
    constexpr double t[9] = {1.0, 0.7660444431189780352023927, 0.1736481776669303488517166, -0.5,
                             -0.9396926207859083840541093, -0.9396926207859083840541093, -0.5,
                             0.1736481776669303488517166, 0.7660444431189780352023927};
    return t[m];
  } else {
    constexpr double t[8] = {1.0, 0.7071067811865475244008444, 0.0, -0.7071067811865475244008444,
                             -1.0, -0.7071067811865475244008444, 0.0, 0.7071067811865475244008444};
    return t[m];
  }
}
template <int R>
APB_HD constexpr double root_sin(int m) {
  if (R == 16) {
    constexpr double t[16] = {0.0, 0.38268343236508977172846, 0.7071067811865475244008444, 0.9238795325112867561281832,
                              1.0, 0.9238795325112867561281832, 0.7071067811865475244008444, 0.38268343236508977172846,
                              0.0, -0.38268343236508977172846, -0.7071067811865475244008444, -0.9238795325112867561281832,
                              -1.0, -0.9238795325112867561281832, -0.7071067811865475244008444, -0.38268343236508977172846};
    return t[m];
  } else if (R == 9) {
    constexpr double t[9] = {0.0, 0.6427876096865393263226434, 0.984807753012208059366743, 0.8660254037844386467637232,
                             0.3420201433256687330440996, -0.3420201433256687330440996, -0.8660254037844386467637232,
                             -0.984807753012208059366743, -0.6427876096865393263226434};
    return t[m];
  } else {
    constexpr double t[8] = {0.0, 0.7071067811865475244008444, 1.0, 0.7071067811865475244008444,
                             0.0, -0.7071067811865475244008444, -1.0, -0.7071067811865475244008444};
    return t[m];
  }
}

// a * exp(-+ 2 pi i m / R) with m a compile-time constant after unrolling
template <int R, bool INV>
APB_HD cpx mul_root(cpx a, int m) {
  m %= R;
  if (m == 0) return a;
  if (2 * m == R) return cpx{-a.x, -a.y};
  if (4 * m == R) return c_rot<INV>(a);
  if (4 * m == 3 * R) return c_rot<!INV>(a);
  const double c = root_cos<R>(m), s = root_sin<R>(m);
  return INV ? cpx{a.x * c - a.y * s, a.x * s + a.y * c} : cpx{a.x * c + a.y * s, a.y * c - a.x * s};
}

// in-register DFT of R values: v[k] <- sum_r v[r] exp(-+ 2 pi i r k / R)
template <int R, bool INV>
struct FftReg;
template <bool INV>
struct FftReg<2, INV> {
  static APB_HD void run(cpx* v) {
    const cpx a = v[0], b = v[1];
    v[0] = c_add(a, b);
    v[1] = c_sub(a, b);
  }
};
template <bool INV>
struct FftReg<4, INV> {
  static APB_HD void run(cpx* v) {
    const cpx t0 = c_add(v[0], v[2]), t1 = c_sub(v[0], v[2]), t2 = c_add(v[1], v[3]), t3 = c_rot<INV>(c_sub(v[1], v[3]));
    v[0] = c_add(t0, t2);
    v[1] = c_add(t1, t3);
    v[2] = c_sub(t0, t2);
    v[3] = c_sub(t1, t3);
  }
};
template <bool INV>
struct FftReg<3, INV> {
  static APB_HD void run(cpx* v) {
    const double s3 = 0.86602540378443864676;
    const cpx t1 = c_add(v[1], v[2]);
    const cpx m = cpx{v[0].x - 0.5 * t1.x, v[0].y - 0.5 * t1.y};
    const cpx dd = c_sub(v[1], v[2]);
    const cpx d = c_rot<INV>(cpx{s3 * dd.x, s3 * dd.y});
    v[0] = c_add(v[0], t1);
    v[1] = c_add(m, d);
    v[2] = c_sub(m, d);
  }
};
template <bool INV>
struct FftReg<5, INV> {
  static APB_HD void run(cpx* v) {
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
    const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
    const cpx v0 = v[0];
    const cpx a1 = c_add(v[1], v[4]), a2 = c_add(v[2], v[3]), b1 = c_sub(v[1], v[4]), b2 = c_sub(v[2], v[3]);
    const cpx p1 = cpx{v0.x + c1 * a1.x + c2 * a2.x, v0.y + c1 * a1.y + c2 * a2.y};
    const cpx p2 = cpx{v0.x + c2 * a1.x + c1 * a2.x, v0.y + c2 * a1.y + c1 * a2.y};
    const cpx q1 = c_rot<INV>(cpx{s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y});
    const cpx q2 = c_rot<INV>(cpx{s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y});
    v[0] = c_add(v0, c_add(a1, a2));
    v[1] = c_add(p1, q1);
    v[2] = c_add(p2, q2);
    v[3] = c_sub(p2, q2);
    v[4] = c_sub(p1, q1);
  }
};
// R = R1 * R2 (Cooley-Tukey in registers): r = R2 a + b, k = c + R1 d
template <int R, int R1, int R2, bool INV>
APB_HD void fft_reg_composite(cpx* v) {
  cpx t[R2][R1];
#pragma unroll
  for (int b = 0; b < R2; ++b) {
#pragma unroll
    for (int a = 0; a < R1; ++a) t[b][a] = v[R2 * a + b];
    FftReg<R1, INV>::run(t[b]);
#pragma unroll
    for (int c = 1; c < R1; ++c) t[b][c] = mul_root<R, INV>(t[b][c], b * c);
  }
#pragma unroll
  for (int c = 0; c < R1; ++c) {
    cpx u[R2];
#pragma unroll
    for (int b = 0; b < R2; ++b) u[b] = t[b][c];
    FftReg<R2, INV>::run(u);
#pragma unroll
    for (int d = 0; d < R2; ++d) v[c + R1 * d] = u[d];
  }
}
template <bool INV>
struct FftReg<8, INV> {
  static APB_HD void run(cpx* v) { fft_reg_composite<8, 4, 2, INV>(v); }
};
template <bool INV>
struct FftReg<9, INV> {
  static APB_HD void run(cpx* v) { fft_reg_composite<9, 3, 3, INV>(v); }
};
template <bool INV>
struct FftReg<16, INV> {
  static APB_HD void run(cpx* v) { fft_reg_composite<16, 4, 4, INV>(v); }
};

// Shared-memory sequences are skewed by one element every 16 (index i lives at i + i/16): the
// first stage of a radix-16 transform writes with stride 16 elements, which without the skew puts
// every lane of a quarter-warp on the same 16-byte bank (8-way conflict).
#define FPAD(i) ((i) + ((i) >> 4))

// powers w^1 .. w^(R-1) of the stage twiddle from ONE table read: products arranged as a tree of
// depth <= 4 (error <= ~5 ulp), which trades the 15 strided, bank-conflicting table reads of a
// radix-16 butterfly for ~80 flops on an FP64 pipe that is otherwise mostly idle
template <int R>
APB_HD void twiddle_powers(cpx w1, cpx* w) {
  w[1] = w1;
#pragma unroll
  for (int r = 2; r < R; ++r) {
    const int h = (r & (r - 1)) == 0 ? r / 2 : (r & -r);   // power of two: square; else split off the low bit
    w[r] = c_mul(w[r - h], w[h]);
  }
}

// One radix-R butterfly of a Stockham autosort stage: loads its R inputs of the length-N sequence
// `in`, applies the stage twiddles and the in-register DFT.  Ns = product of the radices of the
// stages already done, j in [0, N/R).  Returns the index of output 0; output r goes to o + r*Ns
// (both through FPAD).  tw = exp(-2 pi i k / N) table (global memory, read through L1).
template <int R, bool INV, bool PLAIN = false>
APB_HD int fft_bfly(const cpx* __restrict__ in, int N, int Ns, int j, const cpx* __restrict__ tw, cpx* v) {
  const int nb = N / R;
  int k, jq;
  if ((Ns & (Ns - 1)) == 0) {
    k = j & (Ns - 1);
    jq = j - k;
  } else {
    jq = (j / Ns) * Ns;
    k = j - jq;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = in[PLAIN ? (j + r * nb) : FPAD(j + r * nb)];
  if (k) {
    cpx w1 = tw[k * (N / (Ns * R))];
    if (INV) w1.y = -w1.y;
    cpx w[R];
    twiddle_powers<R>(w1, w);
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = c_mul(v[r], w[r]);
  }
  FftReg<R, INV>::run(v);
  return jq * R + k;
}

#if defined(__CUDACC__)
// asynchronous global -> shared copies (LDGSTS): the tile loads of a CTA are all issued before any
// is waited for, instead of one dependent load per loop trip; src_bytes = 0 zero-fills
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ---- bulk asynchronous copies (TMA, SASS UBLKCP) completing on an mbarrier -------------------------------------
// One thread arms the barrier with the byte count and issues the copies; every thread of the CTA waits on the barrier
// before it reads the landing buffer.  Global source, shared destination and size are multiples of 16 bytes.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}

// One stage for nf sequences: in + f*ld -> out + f*ld, one barrier.  A thread keeps one butterfly
// (R complex values) in registers at a time.  noinline on purpose: inlined into the stage loop the
// compiler hoists every radix's fp64 constants out of the loop and spills ~1 KB per thread; as
// separate functions each radix gets its own allocation (<= 112 registers, no spills).
// PLAIN: the input sequences are stored without the skew (the landing buffer of a bulk copy); the output always has it.
template <int R, bool INV, bool PLAIN = false>
__device__ __noinline__ void fft_stage(const cpx* __restrict__ in, cpx* __restrict__ out, int nf, int ld, int N, int Ns,
                                       const cpx* __restrict__ tw) {
  const int nb = N / R, total = nf * nb;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int f = idx / nb, j = idx - f * nb;
    cpx v[R];
    const int o = fft_bfly<R, INV, PLAIN>(in + f * ld, N, Ns, j, tw, v);
#pragma unroll
    for (int r = 0; r < R; ++r) out[f * ld + FPAD(o + r * Ns)] = v[r];
  }
  __syncthreads();
}

// nf FFTs of length D.N at a + f*ld (ping-pong partner b), unnormalised.  All threads of the CTA
// take part; returns the buffer that holds the result.
template <bool INV>
__device__ __forceinline__ cpx* fft_run(cpx* a, cpx* b, int nf, int ld, const FftDesc& D, const cpx* __restrict__ tw,
                                        bool plain_first = false) {
  int Ns = 1;
  const int N = D.N;
#pragma unroll 1
  for (int st = 0; st < D.nstage; ++st) {
    const int R = D.radix[st];
    if (st == 0 && plain_first) {
      switch (R) {
        case 16: fft_stage<16, INV, true>(a, b, nf, ld, N, Ns, tw); break;
        case 9: fft_stage<9, INV, true>(a, b, nf, ld, N, Ns, tw); break;
        case 8: fft_stage<8, INV, true>(a, b, nf, ld, N, Ns, tw); break;
        case 5: fft_stage<5, INV, true>(a, b, nf, ld, N, Ns, tw); break;
        case 4: fft_stage<4, INV, true>(a, b, nf, ld, N, Ns, tw); break;
        case 3: fft_stage<3, INV, true>(a, b, nf, ld, N, Ns, tw); break;
        default: fft_stage<2, INV, true>(a, b, nf, ld, N, Ns, tw); break;
      }
      cpx* t = a;
      a = b;
      b = t;
      Ns *= R;
      continue;
    }
    switch (R) {
      case 16: fft_stage<16, INV>(a, b, nf, ld, N, Ns, tw); break;
      case 9: fft_stage<9, INV>(a, b, nf, ld, N, Ns, tw); break;
      case 8: fft_stage<8, INV>(a, b, nf, ld, N, Ns, tw); break;
      case 5: fft_stage<5, INV>(a, b, nf, ld, N, Ns, tw); break;
      case 4: fft_stage<4, INV>(a, b, nf, ld, N, Ns, tw); break;
      case 3: fft_stage<3, INV>(a, b, nf, ld, N, Ns, tw); break;
      default: fft_stage<2, INV>(a, b, nf, ld, N, Ns, tw); break;
    }
    cpx* t = a;
    a = b;
    b = t;
    Ns *= R;
  }
  return a;
}

// ----------------------------------------------------------------------------
// pass A: real rows -> half spectra.  work: {src, code, row0, nrows};  code >= 0: plane `code` of
// the source's stamp (evaluation region, zero-padded to N);  code < 0: shifted PSF plane -1-code,
// rotated so that its centre sits at column 0.
// spectra layout (cpx), COLUMN-major so that the column pass moves whole columns with bulk copies:
//                       image planes  specA + ((plane*nxh + kx) * eh + row)
//                       PSF planes    specK + ((k*nxh + kx) * sph + a)
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k_fft_rows(const DevSrc* __restrict__ src, const FftDesc* __restrict__ descs,
                                                  const cpx* __restrict__ twid, const int4* __restrict__ work, int mode,
                                                  const double* __restrict__ stamp, const double* __restrict__ psfst,
                                                  cpx* __restrict__ spec) {
  extern __shared__ cpx fsm[];
  __shared__ FftDesc D;
  const int4 wk = work[blockIdx.x];
  const DevSrc& s = src[wk.x];
  const Geo& g = s.geo[mode];
  if (threadIdx.x == 0) D = descs[s.fftx];
  __syncthreads();
  const int N = D.N, nxh = N / 2 + 1;
  const int nf = (wk.w + 1) / 2;
  const int ld = FPAD(N) + 1;
  const cpx* tw = twid + D.tw_off;
  cpx* a = fsm;
  cpx* b = a + nf * ld;
  const bool is_psf = wk.y < 0;
  const double* base;
  int rstride, valid_w, xoff, col_len;
  cpx* dst;
  if (!is_psf) {
    base = stamp + s.stamp_off + (long long)wk.y * s.plane_stride + (long long)(g.ey0 - g.my0) * g.mw + (g.ex0 - g.mx0);
    rstride = g.mw;
    valid_w = g.ew;
    xoff = 0;
    col_len = g.eh;
    dst = spec + s.specA_off + ((long long)wk.y * nxh) * g.eh;
  } else {
    const int kp = -1 - wk.y;
    base = psfst + s.psf_off + (long long)kp * s.spw * s.sph;
    rstride = s.spw;
    valid_w = s.spw;
    xoff = (s.spw - 1) / 2;
    col_len = s.sph;
    dst = spec + s.specK_off + ((long long)kp * nxh) * s.sph;
  }
  for (int idx = threadIdx.x; idx < nf * N; idx += blockDim.x) {
    const int f = idx / N, x = idx - f * N;
    int sx = x + xoff;
    if (sx >= N) sx -= N;
    const int r0 = wk.z + 2 * f;
    const bool v0 = sx < valid_w, v1 = v0 && (2 * f + 1 < wk.w);
    cpx* dst = &a[f * ld + FPAD(x)];
    cp_async8(&dst->x, v0 ? base + (long long)r0 * rstride + sx : base, v0 ? 8 : 0);
    cp_async8(&dst->y, v1 ? base + (long long)(r0 + 1) * rstride + sx : base, v1 ? 8 : 0);
  }
  cp_async_wait_all();
  __syncthreads();
  const cpx* r = fft_run<false>(a, b, nf, ld, D, tw);
  // thread (k, f), f fastest: the two rows of transform f are neighbours in column k, so a warp writes runs of
  // 2 nf x 16 bytes (ld is odd: the strided shared-memory reads hit distinct banks)
  for (int idx = threadIdx.x; idx < nf * nxh; idx += blockDim.x) {
    const int k = idx / nf, f = idx - k * nf;
    const int kn = k ? N - k : 0;
    const cpx zk = r[f * ld + FPAD(k)], zn = r[f * ld + FPAD(kn)];
    const int r0 = wk.z + 2 * f;
    cpx* col = dst + (long long)k * col_len;
    col[r0] = cpx{0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y)};
    if (2 * f + 1 < wk.w) col[r0 + 1] = cpx{0.5 * (zk.y + zn.y), -0.5 * (zk.x - zn.x)};
  }
}

// ----------------------------------------------------------------------------
// pass B: columns.  jobs: {src, in_plane | -1-k (PSF plane k), kernel, out_plane};
// work: {job, kx0, ncols, 0}.  Image job: forward column FFT of `ncols` columns, multiply by the
// PSF spectrum, inverse column FFT, keep the rows of the output window:
//     specB + ((out_plane*nxh + kx) * oh + y)
// PSF job: forward column FFT only:  specKT + ((k*nxh + kx) * Ny + ky)
// Every column is contiguous in global memory (pass A writes column-major), so a tile is moved by bulk asynchronous
// copies (TMA): one thread issues them -- the input columns on one mbarrier, the PSF-spectrum columns on a second --
// and the PSF spectrum arrives under the forward transform instead of stalling the multiply.
// shared memory: A = landing buffer (plain layout) and ping buffer, B = pong buffer, KT = PSF-spectrum tile.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k_fft_cols(const DevSrc* __restrict__ src, const FftDesc* __restrict__ descs,
                                                  const cpx* __restrict__ twid, const int4* __restrict__ jobs,
                                                  const int4* __restrict__ work, int mode, cpx* __restrict__ spec) {
  extern __shared__ cpx fsm[];
  __shared__ FftDesc D;
  __shared__ __align__(8) unsigned long long bar_in, bar_kt;
  const int4 wk = work[blockIdx.x];
  const int4 jb = jobs[wk.x];
  const DevSrc& s = src[jb.x];
  const Geo& g = s.geo[mode];
  if (threadIdx.x == 0) {
    D = descs[s.ffty];
    mbar_init(&bar_in, 1);
    mbar_init(&bar_kt, 1);
    mbar_init_fence();
  }
  __syncthreads();
  const int N = D.N, nxh = s.nxh;
  const int nc = wk.z, kx0 = wk.y;
  const int ld = s.fft_ld;  // skewed length + pad
  const cpx* tw = twid + D.tw_off;
  cpx* a = fsm;
  cpx* b = a + s.fft_nc * ld;
  cpx* ktile = b + s.fft_nc * ld;     // nc x N, plain
  const bool is_psf = jb.y < 0;
  const cpx* in;
  int rows_valid, yoff;
  if (!is_psf) {
    in = spec + s.specA_off + ((long long)jb.y * nxh + kx0) * g.eh;
    rows_valid = g.eh;
    yoff = 0;
  } else {
    in = spec + s.specK_off + ((long long)(-1 - jb.y) * nxh + kx0) * s.sph;
    rows_valid = s.sph;
    yoff = (s.sph - 1) / 2;
  }
  // sequence element y <- column element (y + yoff) mod N: elements [yoff, rows_valid) land at [0, rows_valid - yoff),
  // elements [0, yoff) at [N - yoff, N); everything between is the zero padding
  if (threadIdx.x == 0) {
    const unsigned n_hi = (unsigned)(rows_valid - yoff) * 16u, n_lo = (unsigned)yoff * 16u;
    mbar_expect_tx(&bar_in, (unsigned)nc * (n_hi + n_lo));
    for (int c = 0; c < nc; ++c) {
      tma_load_1d(a + c * ld, in + (long long)c * rows_valid + yoff, n_hi, &bar_in);
      if (n_lo) tma_load_1d(a + c * ld + (N - yoff), in + (long long)c * rows_valid, n_lo, &bar_in);
    }
    if (!is_psf) {
      const cpx* kt = spec + s.specKT_off + ((long long)jb.z * nxh + kx0) * N;
      mbar_expect_tx(&bar_kt, (unsigned)nc * (unsigned)N * 16u);
      for (int c = 0; c < nc; ++c) tma_load_1d(ktile + c * N, kt + (long long)c * N, (unsigned)N * 16u, &bar_kt);
    }
  }
  {
    const int z0 = rows_valid - yoff, nz = N - rows_valid;
    for (int idx = threadIdx.x; idx < nc * nz; idx += blockDim.x) {
      const int c = idx / nz, y = z0 + (idx - c * nz);
      a[c * ld + y] = cpx{0.0, 0.0};
    }
  }
  __syncthreads();          // the zero padding is in place
  mbar_wait(&bar_in, 0);    // the columns have landed
  cpx* r = fft_run<false>(a, b, nc, ld, D, tw, true);
  if (is_psf) {
    cpx* kt = spec + s.specKT_off + ((long long)(-1 - jb.y) * nxh + kx0) * N;
    for (int idx = threadIdx.x; idx < nc * N; idx += blockDim.x) {
      const int c = idx / N, y = idx - c * N;
      kt[(long long)c * N + y] = r[c * ld + FPAD(y)];
    }
    return;
  }
  mbar_wait(&bar_kt, 0);
  const double scale = 1.0 / ((double)N * (double)s.fft_nx);
#pragma unroll 4
  for (int idx = threadIdx.x; idx < nc * N; idx += blockDim.x) {
    const int c = idx / N, y = idx - c * N;
    const cpx v = c_mul(r[c * ld + FPAD(y)], ktile[c * N + y]);
    r[c * ld + FPAD(y)] = cpx{v.x * scale, v.y * scale};
  }
  __syncthreads();
  const cpx* z = fft_run<true>(r, r == a ? b : a, nc, ld, D, tw);
  // (foh, fow: the output window on the source's grid -- oh, ow unless the PSF is super-sampled, DevSrc::up)
  cpx* out = spec + s.specB_off + ((long long)jb.w * nxh + kx0) * s.foh;
  for (int idx = threadIdx.x; idx < nc * s.foh; idx += blockDim.x) {
    const int c = idx / s.foh, y = idx - c * s.foh;
    out[(long long)c * s.foh + y] = z[c * ld + FPAD(y + s.by)];
  }
}

// ----------------------------------------------------------------------------
// pass C: half spectra -> real rows of the output window.  work: {src, out_plane, row0, nrows}
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k_fft_rows_inv(const DevSrc* __restrict__ src, const FftDesc* __restrict__ descs,
                                                      const cpx* __restrict__ twid, const int4* __restrict__ work,
                                                      const cpx* __restrict__ spec, double* __restrict__ outar) {
  extern __shared__ cpx fsm[];
  __shared__ FftDesc D;
  const int4 wk = work[blockIdx.x];
  const DevSrc& s = src[wk.x];
  if (threadIdx.x == 0) D = descs[s.fftx];
  __syncthreads();
  const int N = D.N, nxh = s.nxh;
  const int nf = (wk.w + 1) / 2;
  const int ld = FPAD(N) + 1;
  const cpx* tw = twid + D.tw_off;
  cpx* a = fsm;
  cpx* b = a + nf * ld;
  const cpx* in = spec + s.specB_off + ((long long)wk.y * nxh) * s.foh;
  // thread (k, f), f fastest: a warp reads runs of 2 nf x 16 bytes of column kk (column-major spectra)
#pragma unroll 4
  for (int idx = threadIdx.x; idx < nf * N; idx += blockDim.x) {
    const int k = idx / nf, f = idx - k * nf;
    const int r0 = wk.z + 2 * f;
    const bool second = 2 * f + 1 < wk.w;
    const bool lo = 2 * k <= N;
    const int kk = lo ? k : N - k;
    const cpx* col = in + (long long)kk * s.foh;
    const cpx x1 = col[r0];
    const cpx x2 = second ? col[r0 + 1] : cpx{0.0, 0.0};
    a[f * ld + FPAD(k)] = lo ? cpx{x1.x - x2.y, x1.y + x2.x} : cpx{x1.x + x2.y, -x1.y + x2.x};
  }
  __syncthreads();
  const cpx* z = fft_run<true>(a, b, nf, ld, D, tw);
  double* o = outar + s.fine_off + (long long)wk.y * s.fow * s.foh;
  for (int idx = threadIdx.x; idx < nf * s.fow; idx += blockDim.x) {
    const int f = idx / s.fow, x = idx - f * s.fow;
    const int r0 = wk.z + 2 * f;
    const cpx v = z[f * ld + FPAD(x + s.bx)];
    o[(long long)r0 * s.fow + x] = v.x;
    if (2 * f + 1 < wk.w) o[(long long)(r0 + 1) * s.fow + x] = v.y;
  }
}
#endif  // __CUDACC__
