// FFT convolution for large PSFs (utils/operations.py:9-36 `fft_convolve_torch`, called from
// models/_model_methods.py:233-257).  Hand-written fp64 mixed-radix (2/3/4/5) Stockham FFTs in
// shared memory, three passes per plane:
//   k_fft_rows     real rows -> half spectra           (two real rows share one complex FFT)
//   k_fft_cols     column FFT . PSF spectrum . inverse column FFT, all in shared memory: the
//                  pointwise multiply never touches HBM and the full 2-D spectrum is never stored
//   k_fft_rows_inv half spectra -> real rows, cropped to the output window
// The transform lengths are the next 2^a 3^b 5^c >= the padded stamp, not the reference's exact
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
template <bool INV>
APB_HD cpx c_tw(const cpx* __restrict__ tw, int idx) {
  cpx w = tw[idx];
  if (INV) w.y = -w.y;
  return w;
}

// One radix-R butterfly of a Stockham autosort stage.  `in`/`out` hold one length-N sequence,
// Ns = product of the radices of the stages already done, j in [0, N/R).
template <bool INV>
APB_HD void fft_butterfly(const cpx* __restrict__ in, cpx* __restrict__ out, int N, int R, int Ns, int j,
                          const cpx* __restrict__ tw) {
  const int nb = N / R;
  int k, jq;
  if ((Ns & (Ns - 1)) == 0) {
    k = j & (Ns - 1);
    jq = j - k;
  } else {
    jq = (j / Ns) * Ns;
    k = j - jq;
  }
  const int o = jq * R + k;
  const int step = k * (N / (Ns * R));  // twiddle index of w^(k) for this stage
  if (R == 4) {
    cpx v0 = in[j], v1 = in[j + nb], v2 = in[j + 2 * nb], v3 = in[j + 3 * nb];
    if (k) {
      v1 = c_mul(v1, c_tw<INV>(tw, step));
      v2 = c_mul(v2, c_tw<INV>(tw, 2 * step));
      v3 = c_mul(v3, c_tw<INV>(tw, 3 * step));
    }
    const cpx t0 = c_add(v0, v2), t1 = c_sub(v0, v2), t2 = c_add(v1, v3), t3 = c_rot<INV>(c_sub(v1, v3));
    out[o] = c_add(t0, t2);
    out[o + Ns] = c_add(t1, t3);
    out[o + 2 * Ns] = c_sub(t0, t2);
    out[o + 3 * Ns] = c_sub(t1, t3);
  } else if (R == 2) {
    cpx v0 = in[j], v1 = in[j + nb];
    if (k) v1 = c_mul(v1, c_tw<INV>(tw, step));
    out[o] = c_add(v0, v1);
    out[o + Ns] = c_sub(v0, v1);
  } else if (R == 3) {
    cpx v0 = in[j], v1 = in[j + nb], v2 = in[j + 2 * nb];
    if (k) {
      v1 = c_mul(v1, c_tw<INV>(tw, step));
      v2 = c_mul(v2, c_tw<INV>(tw, 2 * step));
    }
    const double s3 = 0.86602540378443864676;
    const cpx t1 = c_add(v1, v2);
    const cpx m = cpx{v0.x - 0.5 * t1.x, v0.y - 0.5 * t1.y};
    const cpx dd = c_sub(v1, v2);
    const cpx d = c_rot<INV>(cpx{s3 * dd.x, s3 * dd.y});
    out[o] = c_add(v0, t1);
    out[o + Ns] = c_add(m, d);
    out[o + 2 * Ns] = c_sub(m, d);
  } else {  // R == 5
    cpx v0 = in[j], v1 = in[j + nb], v2 = in[j + 2 * nb], v3 = in[j + 3 * nb], v4 = in[j + 4 * nb];
    if (k) {
      v1 = c_mul(v1, c_tw<INV>(tw, step));
      v2 = c_mul(v2, c_tw<INV>(tw, 2 * step));
      v3 = c_mul(v3, c_tw<INV>(tw, 3 * step));
      v4 = c_mul(v4, c_tw<INV>(tw, 4 * step));
    }
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
    const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
    const cpx a1 = c_add(v1, v4), a2 = c_add(v2, v3), b1 = c_sub(v1, v4), b2 = c_sub(v2, v3);
    const cpx p1 = cpx{v0.x + c1 * a1.x + c2 * a2.x, v0.y + c1 * a1.y + c2 * a2.y};
    const cpx p2 = cpx{v0.x + c2 * a1.x + c1 * a2.x, v0.y + c2 * a1.y + c1 * a2.y};
    const cpx q1 = c_rot<INV>(cpx{s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y});
    const cpx q2 = c_rot<INV>(cpx{s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y});
    out[o] = c_add(v0, c_add(a1, a2));
    out[o + Ns] = c_add(p1, q1);
    out[o + 2 * Ns] = c_add(p2, q2);
    out[o + 3 * Ns] = c_sub(p2, q2);
    out[o + 4 * Ns] = c_sub(p1, q1);
  }
}

#if defined(__CUDACC__)
// nf FFTs of length D.N living at a + f*ld (ping-pong partner b).  All threads of the CTA
// take part; returns the buffer that holds the (unnormalised) result.
template <bool INV>
__device__ __forceinline__ cpx* fft_run(cpx* a, cpx* b, int nf, int ld, const FftDesc& D, const cpx* __restrict__ tw) {
  int Ns = 1;
  const int N = D.N;
  for (int st = 0; st < D.nstage; ++st) {
    const int R = D.radix[st];
    const int nb = N / R;
    const int total = nf * nb;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int f = idx / nb, j = idx - f * nb;
      fft_butterfly<INV>(a + f * ld, b + f * ld, N, R, Ns, j, tw);
    }
    __syncthreads();
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
// spectra layout (cpx): image planes  specA + ((plane*eh + row) * nxp + kx)
//                       PSF planes    specK + ((k*sph + a) * nxp + kx)
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fft_rows(const DevSrc* __restrict__ src, const FftDesc* __restrict__ descs,
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
  const int N = D.N, nxh = N / 2 + 1, nxp = s.nxp;
  const int nf = (wk.w + 1) / 2;
  cpx* tw = fsm;
  cpx* a = tw + N;
  cpx* b = a + nf * N;
  for (int q = threadIdx.x; q < N; q += blockDim.x) tw[q] = twid[D.tw_off + q];
  const bool is_psf = wk.y < 0;
  const double* base;
  int rstride, valid_w, xoff;
  cpx* dst;
  if (!is_psf) {
    base = stamp + s.stamp_off + (long long)wk.y * s.plane_stride + (long long)(g.ey0 - g.my0) * g.mw + (g.ex0 - g.mx0);
    rstride = g.mw;
    valid_w = g.ew;
    xoff = 0;
    dst = spec + s.specA_off + ((long long)wk.y * g.eh) * nxp;
  } else {
    const int kp = -1 - wk.y;
    base = psfst + s.psf_off + (long long)kp * s.spw * s.sph;
    rstride = s.spw;
    valid_w = s.spw;
    xoff = (s.spw - 1) / 2;
    dst = spec + s.specK_off + ((long long)kp * s.sph) * nxp;
  }
  for (int idx = threadIdx.x; idx < nf * N; idx += blockDim.x) {
    const int f = idx / N, x = idx - f * N;
    int sx = x + xoff;
    if (sx >= N) sx -= N;
    const int r0 = wk.z + 2 * f;
    double re = 0.0, im = 0.0;
    if (sx < valid_w) {
      re = base[(long long)r0 * rstride + sx];
      if (2 * f + 1 < wk.w) im = base[(long long)(r0 + 1) * rstride + sx];
    }
    a[idx] = cpx{re, im};
  }
  __syncthreads();
  const cpx* r = fft_run<false>(a, b, nf, N, D, tw);
  for (int idx = threadIdx.x; idx < nf * nxh; idx += blockDim.x) {
    const int f = idx / nxh, k = idx - f * nxh;
    const cpx zk = r[f * N + k], zn = r[f * N + (k ? N - k : 0)];
    const int r0 = wk.z + 2 * f;
    dst[(long long)r0 * nxp + k] = cpx{0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y)};
    if (2 * f + 1 < wk.w) dst[(long long)(r0 + 1) * nxp + k] = cpx{0.5 * (zk.y + zn.y), -0.5 * (zk.x - zn.x)};
  }
}

// ----------------------------------------------------------------------------
// pass B: columns.  jobs: {src, in_plane | -1-k (PSF plane k), kernel, out_plane};
// work: {job, kx0, ncols, 0}.  Image job: forward column FFT of `ncols` columns, multiply by the
// PSF spectrum, inverse column FFT, keep the rows of the output window:
//     specB + ((out_plane*oh + y) * nxp + kx)
// PSF job: forward column FFT only, stored column-major:  specKT + ((k*nxh + kx) * Ny + ky)
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fft_cols(const DevSrc* __restrict__ src, const FftDesc* __restrict__ descs,
                                                  const cpx* __restrict__ twid, const int4* __restrict__ jobs,
                                                  const int4* __restrict__ work, int mode, cpx* __restrict__ spec) {
  extern __shared__ cpx fsm[];
  __shared__ FftDesc D;
  const int4 wk = work[blockIdx.x];
  const int4 jb = jobs[wk.x];
  const DevSrc& s = src[jb.x];
  const Geo& g = s.geo[mode];
  if (threadIdx.x == 0) D = descs[s.ffty];
  __syncthreads();
  const int N = D.N, nxp = s.nxp, nxh = s.nxh;
  const int nc = wk.z, kx0 = wk.y;
  const int ld = N + 8 / s.fft_nc;  // pad: the transposing tile load/store is bank-conflict free
  cpx* tw = fsm;
  cpx* a = tw + N;
  cpx* b = a + s.fft_nc * ld;
  for (int q = threadIdx.x; q < N; q += blockDim.x) tw[q] = twid[D.tw_off + q];
  const bool is_psf = jb.y < 0;
  const cpx* in;
  int rows_valid, yoff;
  if (!is_psf) {
    in = spec + s.specA_off + ((long long)jb.y * g.eh) * nxp + kx0;
    rows_valid = g.eh;
    yoff = 0;
  } else {
    in = spec + s.specK_off + ((long long)(-1 - jb.y) * s.sph) * nxp + kx0;
    rows_valid = s.sph;
    yoff = (s.sph - 1) / 2;
  }
  for (int idx = threadIdx.x; idx < nc * N; idx += blockDim.x) {
    const int y = idx / nc, c = idx - y * nc;
    int sy = y + yoff;
    if (sy >= N) sy -= N;
    a[c * ld + y] = sy < rows_valid ? in[(long long)sy * nxp + c] : cpx{0.0, 0.0};
  }
  __syncthreads();
  cpx* r = fft_run<false>(a, b, nc, ld, D, tw);
  if (is_psf) {
    cpx* kt = spec + s.specKT_off + ((long long)(-1 - jb.y) * nxh + kx0) * N;
    for (int idx = threadIdx.x; idx < nc * N; idx += blockDim.x) {
      const int c = idx / N, y = idx - c * N;
      kt[(long long)c * N + y] = r[c * ld + y];
    }
    return;
  }
  const cpx* kt = spec + s.specKT_off + ((long long)jb.z * nxh + kx0) * N;
  const double scale = 1.0 / ((double)N * (double)s.fft_nx);
  for (int idx = threadIdx.x; idx < nc * N; idx += blockDim.x) {
    const int c = idx / N, y = idx - c * N;
    const cpx v = c_mul(r[c * ld + y], kt[(long long)c * N + y]);
    r[c * ld + y] = cpx{v.x * scale, v.y * scale};
  }
  __syncthreads();
  cpx* other = (r == a) ? b : a;
  const cpx* z = fft_run<true>(r, other, nc, ld, D, tw);
  cpx* out = spec + s.specB_off + ((long long)jb.w * s.oh) * nxp + kx0;
  for (int idx = threadIdx.x; idx < nc * s.oh; idx += blockDim.x) {
    const int y = idx / nc, c = idx - y * nc;
    out[(long long)y * nxp + c] = z[c * ld + y + s.by];
  }
}

// ----------------------------------------------------------------------------
// pass C: half spectra -> real rows of the output window.  work: {src, out_plane, row0, nrows}
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fft_rows_inv(const DevSrc* __restrict__ src, const FftDesc* __restrict__ descs,
                                                      const cpx* __restrict__ twid, const int4* __restrict__ work,
                                                      const cpx* __restrict__ spec, double* __restrict__ outar) {
  extern __shared__ cpx fsm[];
  __shared__ FftDesc D;
  const int4 wk = work[blockIdx.x];
  const DevSrc& s = src[wk.x];
  if (threadIdx.x == 0) D = descs[s.fftx];
  __syncthreads();
  const int N = D.N, nxp = s.nxp;
  const int nf = (wk.w + 1) / 2;
  cpx* tw = fsm;
  cpx* a = tw + N;
  cpx* b = a + nf * N;
  for (int q = threadIdx.x; q < N; q += blockDim.x) tw[q] = twid[D.tw_off + q];
  const cpx* in = spec + s.specB_off + ((long long)wk.y * s.oh) * nxp;
  for (int idx = threadIdx.x; idx < nf * N; idx += blockDim.x) {
    const int f = idx / N, k = idx - f * N;
    const int r0 = wk.z + 2 * f;
    const bool second = 2 * f + 1 < wk.w;
    const bool lo = 2 * k <= N;
    const int kk = lo ? k : N - k;
    const cpx x1 = in[(long long)r0 * nxp + kk];
    const cpx x2 = second ? in[(long long)(r0 + 1) * nxp + kk] : cpx{0.0, 0.0};
    a[idx] = lo ? cpx{x1.x - x2.y, x1.y + x2.x} : cpx{x1.x + x2.y, -x1.y + x2.x};
  }
  __syncthreads();
  const cpx* z = fft_run<true>(a, b, nf, N, D, tw);
  double* o = outar + s.out_off + (long long)wk.y * s.ow * s.oh;
  for (int idx = threadIdx.x; idx < nf * s.ow; idx += blockDim.x) {
    const int f = idx / s.ow, x = idx - f * s.ow;
    const int r0 = wk.z + 2 * f;
    const cpx v = z[f * N + x + s.bx];
    o[(long long)r0 * s.ow + x] = v.x;
    if (2 * f + 1 < wk.w) o[(long long)(r0 + 1) * s.ow + x] = v.y;
  }
}
#endif  // __CUDACC__
