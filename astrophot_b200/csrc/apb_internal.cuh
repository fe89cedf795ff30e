// Internal device structures and the per-point profile evaluation shared by the
// kernels of libastrophot_b200.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/astrophot_b200.h"
#include "apb_math.cuh"

#define APB_LN10 2.302585092994045684
#define APB_PI 3.141592653589793238

// -----------------------------------------------------------------------------
// static per-source table (built once per plan)
// -----------------------------------------------------------------------------
struct Geo {
  int ex0, ey0, ew, eh;  // evaluation region: out window +- psf border (image pixels)
  int rx0, ry0, rw, rh;  // working region: working window +- psf border
  int mx0, my0, mw, mh;  // stamp region: evaluation region + 1 px ring, clipped to working region
  int tile0, ntile;      // this source's tiles in the mode's tile list
  int chunk0, nchunk;    // mean-reference partial sums (REF_MEAN only)
};

struct DevSrc {
  int kind, flags, image, n_elem;
  int ox, oy, ow, oh;
  Geo geo[2];  // 0 = forward sampling, 1 = jacobian
  int slot[APB_MAX_ELEM];
  double cval[APB_MAX_ELEM];
  int plane[APB_MAX_ELEM];  // derivative plane (1..n_act) of each element, 0 = none
  int n_act;
  // auxiliary PSF model: its n_pp free parameters are pseudo-elements n_elem .. n_elem_all-1 of this source
  // (slot / plane filled, never seen by the profile evaluation); psf_src = the source that samples the PSF
  int n_elem_all, psf_src, n_pp;
  // APB_F_AMP: pseudo-element (slot / cval / plane filled, value and chain factor in DevDyn::el / chain) holding the
  // log10 amplitude applied by k_amp after sampling and normalisation; -1 = none
  int amp_elem;
  int n_prof;
  double prof[APB_MAX_PROF];
  int sampling_mode, quad_init, integrate_mode, quad_level, gridding, max_depth, ref_mode;
  double tol, soft2;
  // sub-pixel re-gridding (utils/operations.py:94-102,150-247), precomputed by the host with the arithmetic the kernels
  // used to repeat per cell (divisions: ~60 instructions each on the device):
  //   gsc[d], gasc[d]: edge / area of a depth-d cell in pixel units (gsc[1] = 1, gsc[d+1] = gsc[d] / G, gasc[d+1] = gasc[d] / G^2)
  //   goff[i]: offset of child column / row i of a cell, in units of the cell's edge: -(G-1)/(2G) + i/G
  //   g2d = G^2 (the error threshold grows by it per level);  gmagic: (ch * gmagic) >> 16 == ch / G for ch < G^2
  double gsc[APB_MAX_DEPTH + 2], gasc[APB_MAX_DEPTH + 2], goff[16], g2d;
  int gmagic;
  int psf, psf_shift, bx, by, pw, ph;  // pw, ph: raw PSF size
  long long stamp_off;     // doubles, into the stamp arena; plane p at stamp_off + p*plane_stride
  long long plane_stride;  // >= mw*mh of both modes
  long long out_off;       // PSF sources: offset of cropped planes (ow*oh each) in the out arena; else -1
  // super-sampled PSF (apb_source_t.upscale = up > 1): S, Sinv, rij, area, geo[], bx, by and the PSF stamps live on the
  // fine grid (pixels 1 / up of the image's); the convolution (k_point) writes the fine output window (fox, foy, fow, foh
  // = up x the output window) at fine_off, k_reduce_up block-sums it into the planes at out_off that everything
  // downstream reads.  up == 1: the fine window is the output window and fine_off == out_off
  int up, fox, foy, fow, foh;
  long long fine_off;
  long long psf_off;       // offset of this source's shifted PSF stamps (3 x spw*sph) or -1
  int spw, sph;            // shifted stamp size
  double S[4], Sinv[4], rij[2], rxy[2], area;
  const unsigned char* mask;            // the model's own mask (non-zero: no contribution) or NULL
  int mask_x0, mask_y0, mask_w, mask_h; // image pixel of its element (0, 0), its shape
  int same_geo;            // forward and jacobian geometry identical
  // FFT convolution (apb_fft.cuh); conv_fft = 0: tiled direct convolution
  int conv_fft;
  int fftx, ffty;          // FftDesc index of the row / column transform
  int fft_nx, nxh, nxp;    // row length, nx/2+1, padded spectrum row stride (cpx)
  int fft_nf, fft_nc;      // complex row FFTs per CTA, columns per CTA
  int fft_ld;              // column tile leading dimension in shared memory (Ny + pad)
  long long specA_off, specB_off, specK_off, specKT_off;  // cpx offsets into the spectrum arena
};

// per-call, per-source values
struct DevDyn {
  double el[APB_MAX_ELEM];
  double chain[APB_MAX_ELEM];  // d value / d x  (0 for locked)
  double c, s, qinv;           // rotation by -(PA - pi/2), 1/q
  double k[8];                 // per-kind constants (see prep kernel)
  double sx, sy;               // sub-pixel shift of the centre (PSF sources)
  int rx, ry;                  // rounded centre pixel
  double thr[2];               // tolerance * reference per mode
  double spl_m[APB_MAX_PROF];  // spline slopes
  double rcut;                 // plane-unit radius beyond which the profile cannot change the mean reference (inf: none)
};

// view of one output plane of a source, in the output window
struct PlaneView {
  const double* p;
  int stride;
};

__device__ __forceinline__ PlaneView out_plane(const DevSrc& s, int mode, int plane, const double* stamp,
                                               const double* outar) {
  PlaneView v;
  if (s.out_off >= 0) {
    v.p = outar + s.out_off + (long long)plane * s.ow * s.oh;
    v.stride = s.ow;
  } else {
    const Geo& g = s.geo[mode];
    v.p = stamp + s.stamp_off + plane * s.plane_stride + (long long)(s.oy - g.my0) * g.mw + (s.ox - g.mx0);
    v.stride = g.mw;
  }
  return v;
}

// model_object.py:370-371: working_image.data * logical_not(self.mask), applied where a source's planes are consumed
__device__ __forceinline__ bool src_masked(const DevSrc& s, int x, int y) {
  if (!s.mask) return false;
  const int mx = x - s.mask_x0, my = y - s.mask_y0;
  return mx >= 0 && mx < s.mask_w && my >= 0 && my < s.mask_h && s.mask[(long long)my * s.mask_w + mx] != 0;
}

// -----------------------------------------------------------------------------
// quadrature tables
// -----------------------------------------------------------------------------
struct QuadTab {
  double a[APB_MAX_QUAD + 1][APB_MAX_QUAD];  // abscissae / 2  (unit pixel offsets)
  double w[APB_MAX_QUAD + 1][APB_MAX_QUAD];  // weights / 2
};
__constant__ QuadTab c_quad;

// -----------------------------------------------------------------------------
// profile evaluation
// -----------------------------------------------------------------------------
__device__ __forceinline__ double sersic_b(double n) {
  double i = 1.0 / n;
  return 2 * n - 1.0 / 3 + i * (4.0 / 405 + i * (46.0 / 25515 + i * (131.0 / 1148175 - i * (2194697.0 / 30690717750.0))));
}
__device__ __forceinline__ double sersic_db(double n) {
  double i = 1.0 / n;
  double i2 = i * i;
  return 2 - i2 * (4.0 / 405 + i * (92.0 / 25515 + i * (393.0 / 1148175 - i * (4 * 2194697.0 / 30690717750.0))));
}

template <int KIND>
struct KindInfo {};
template <> struct KindInfo<APB_SERSIC> { static constexpr int NE = 7; };
template <> struct KindInfo<APB_PLANE_SKY> { static constexpr int NE = 5; };
template <> struct KindInfo<APB_EXPONENTIAL> { static constexpr int NE = 6; };
template <> struct KindInfo<APB_GAUSSIAN> { static constexpr int NE = 6; };
template <> struct KindInfo<APB_MOFFAT> { static constexpr int NE = 7; };
template <> struct KindInfo<APB_SPLINE> { static constexpr int NE = APB_MAX_ELEM; };

// Arithmetic of the profile kernels: fp64 (table-driven exp / log, apb_math.cuh) or fp32 (the reference's
// AP_config.ap_dtype = float32: the CUDA single-precision functions, 1-2 ulp).
template <typename T> struct MT;
template <> struct MT<double> {
  static __device__ __forceinline__ double exp(double x) { return apb_exp(x, s_mathtab); }
  static __device__ __forceinline__ double log(double x) { return apb_log(x, s_mathtab); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double rcp(double x) { return apb_rcp(x); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
};
template <> struct MT<float> {
  static __device__ __forceinline__ float exp(float x) { return expf(x); }
  static __device__ __forceinline__ float log(float x) { return logf(x); }
  static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float rcp(float x) { return 1.0f / x; }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
};

// Per-source values of a profile evaluation, held in registers while a thread works on one source (the kernels used to
// re-read them from the DevSrc / DevDyn tables at every node: a third of the instructions of an evaluation).
template <typename T>
struct PCtxT {
  T c, s, qinv, soft2;   // rotation by -(PA - pi/2), 1/q, softening^2
  T k0;                  // amplitude constant times the sub-pixel area factor of the cell being integrated
  T k1, k2, k3, k4, k5, k6;
  T q;                   // axis ratio (d/dPA)
  int radial;
};
typedef PCtxT<double> PCtx;
template <typename T>
__device__ __forceinline__ void pctx_load(PCtxT<T>& c, const DevSrc& s, const DevDyn& d, double ascale) {
  c.c = (T)d.c; c.s = (T)d.s; c.qinv = (T)d.qinv; c.soft2 = (T)s.soft2;
  c.k0 = (T)(ascale * d.k[0]);
  c.k1 = (T)d.k[1]; c.k2 = (T)d.k[2]; c.k3 = (T)d.k[3]; c.k4 = (T)d.k[4]; c.k5 = (T)d.k[5]; c.k6 = (T)d.k[6];
  c.q = (T)d.el[2];
  c.radial = s.flags & APB_F_RADIAL;
}

// softened squared radius: ONE expression for every kernel, so that all of them see the same bits
template <typename T>
__device__ __forceinline__ T r2_of(T xp, T yp, T soft2) { return MT<T>::fma(xp, xp, MT<T>::fma(yp, yp, soft2)); }

// rotated, axis-ratio-scaled offsets (_shared_methods.py:274-307, coordinates.py:5-13).  PSF models are circular: k_prep
// gives them c = 1, s = 0, 1/q = 1, for which these products are exact.
template <typename T>
__device__ __forceinline__ void rot_coords(const PCtxT<T>& c, T X, T Y, T& xp, T& yp) {
  xp = c.c * X - c.s * Y;
  yp = (c.s * X + c.c * Y) * c.qinv;
}

// Value of an analytic profile at squared radius R2 WITHOUT range checks in exp / log: `bad` is raised instead when an
// argument leaves their fast range (then the value is garbage and the caller redoes it with eval_rot).  Three of these
// run side by side per thread in the integration kernels; one predicate and one branch serve all of them.
// (fp32: the library functions cover every argument, `bad` stays false.)
template <int KIND>
__device__ __forceinline__ double prof_fast(const PCtxT<double>& c, double R2, bool& bad) {
  if (KIND == APB_SERSIC) {
    const double z = R2 * c.k1;
    const double L2 = apb_log_nc(z, s_mathtab);
    const double t = L2 * c.k2;
    const double u = apb_exp_nc(t, s_mathtab);
    const double v = fma(-c.k3, u, c.k3);
    bad = bad || apb_log_bad(z) || apb_exp_bad(t) || apb_exp_bad(v);
    return c.k0 * apb_exp_nc(v, s_mathtab);
  } else if (KIND == APB_EXPONENTIAL) {
    const double R = sqrt(R2);
    const double v = -c.k2 * (R * c.k1 - 1.0);
    bad = bad || apb_exp_bad(v);
    return c.k0 * apb_exp_nc(v, s_mathtab);
  } else if (KIND == APB_GAUSSIAN) {
    const double v = -0.5 * R2 * c.k1;
    bad = bad || apb_exp_bad(v);
    return c.k0 * apb_exp_nc(v, s_mathtab);
  } else {   // APB_MOFFAT
    const double t = 1.0 + R2 * c.k1;
    const double lt = apb_log_nc(t, s_mathtab);
    const double v = -c.k2 * lt;
    bad = bad || apb_log_bad(t) || apb_exp_bad(v);
    return c.k0 * apb_exp_nc(v, s_mathtab);
  }
}
template <int KIND>
__device__ __forceinline__ float prof_fast(const PCtxT<float>& c, float R2, bool& bad) {
  if (KIND == APB_SERSIC) {
    const float u = expf(logf(R2 * c.k1) * c.k2);
    return c.k0 * expf(fmaf(-c.k3, u, c.k3));
  } else if (KIND == APB_EXPONENTIAL) {
    return c.k0 * expf(-c.k2 * (sqrtf(R2) * c.k1 - 1.0f));
  } else if (KIND == APB_GAUSSIAN) {
    return c.k0 * expf(-0.5f * R2 * c.k1);
  } else {   // APB_MOFFAT
    return c.k0 * expf(-c.k2 * logf(1.0f + R2 * c.k1));
  }
}

// Brightness at rotated offsets (xp, yp) -- c.k0 carries the sub-pixel area factor -- and optionally dI/d(element) for
// every element (natural units).  dI must hold the kind's element count.
template <int KIND, bool GRAD, typename T>
__device__ __forceinline__ T eval_rot(const PCtxT<T>& c, const DevSrc& s, const DevDyn& d, T xp, T yp,
                                      T* __restrict__ dI) {
  typedef MT<T> M;
  const T LN10 = (T)APB_LN10;
  const T R2 = r2_of<T>(xp, yp, c.soft2);
  T I, dIdR_over_R;  // (dI/dR)/R : multiply by xp, yp pieces to get dI/dX
  if (KIND == APB_SERSIC) {
    // k0 = area*10^Ie, k1 = 1/Re^2, k2 = 1/(2n), k3 = b_n, k4 = b'_n, k5 = 1/n, k6 = 1/Re
    const T L2 = M::log(R2 * c.k1);  // 2 ln(R/Re)
    const T u = M::exp(L2 * c.k2);
    I = c.k0 * M::exp(M::fma(-c.k3, u, c.k3));
    if (GRAD) {
      const T bu = c.k3 * u;
      dIdR_over_R = -I * bu * c.k5 * M::rcp(R2);
      dI[4] = I * (-c.k4 * (u - T(1.0)) + bu * (T(0.5) * L2) * c.k5 * c.k5);
      dI[5] = I * bu * c.k5 * c.k6;
      dI[6] = LN10 * I;
    }
  } else if (KIND == APB_EXPONENTIAL) {
    // k0 = area*10^Ie, k1 = 1/Re, k2 = b_1
    const T R = M::sqrt(R2);
    I = c.k0 * M::exp(-c.k2 * (R * c.k1 - T(1.0)));
    if (GRAD) {
      dIdR_over_R = -I * c.k2 * c.k1 / R;
      dI[4] = I * c.k2 * R * c.k1 * c.k1;
      dI[5] = LN10 * I;
    }
  } else if (KIND == APB_GAUSSIAN) {
    // k0 = area*10^flux/sqrt(2 pi sigma^2), k1 = 1/sigma^2, k2 = 1/sigma
    I = c.k0 * M::exp(T(-0.5) * R2 * c.k1);
    if (GRAD) {
      dIdR_over_R = -I * c.k1;
      dI[4] = I * (-c.k2 + R2 * c.k1 * c.k2);
      dI[5] = LN10 * I;
    }
  } else if (KIND == APB_MOFFAT) {
    // k0 = area*10^I0, k1 = 1/Rd^2, k2 = n, k3 = 1/Rd
    const T t = T(1.0) + R2 * c.k1;
    const T lt = M::log(t);
    I = c.k0 * M::exp(-c.k2 * lt);
    if (GRAD) {
      dIdR_over_R = -I * c.k2 * T(2.0) * c.k1 / t;
      dI[4] = -I * lt;
      dI[5] = I * c.k2 * T(2.0) * R2 * c.k1 * c.k3 / t;
      dI[6] = LN10 * I;
    }
  } else {  // APB_SPLINE: k0 = area
    const T R = M::sqrt(R2);
    const int K = s.n_prof;
    // idx = searchsorted(prof[:-1], R, left) - 1, wrapping -1 -> K-1 (utils/interpolate.py:53)
    int idx = -1;
    for (int k = 0; k < K - 1; ++k)
      if ((T)s.prof[k] < R) idx = k;
    const double* v = d.el + 4;
    T sv, dsdR;
    if (GRAD)
      for (int k = 0; k < K; ++k) dI[4 + k] = T(0.0);
    if (R > (T)s.prof[K - 1]) {
      const T h = (T)(s.prof[K - 1] - s.prof[K - 2]);
      const T f = (R - (T)s.prof[K - 2]) / h;
      sv = (T)v[K - 2] + (R - (T)s.prof[K - 2]) * ((T)(v[K - 1] - v[K - 2]) / h);
      dsdR = (T)(v[K - 1] - v[K - 2]) / h;
      I = c.k0 * M::exp(LN10 * sv);
      if (GRAD) {
        dI[4 + K - 2] = LN10 * I * (T(1.0) - f);
        dI[4 + K - 1] = LN10 * I * f;
      }
    } else {
      const int i0 = idx < 0 ? K - 1 : idx;
      const int i1 = idx + 1;
      const T dx = (T)(s.prof[i1] - s.prof[i0]);
      const T t = (R - (T)s.prof[i0]) / dx;
      const T t2 = t * t, t3 = t2 * t;
      const T h00 = 1 - 3 * t2 + 2 * t3, h10 = t - 2 * t2 + t3, h01 = 3 * t2 - 2 * t3, h11 = t3 - t2;
      const T v0 = (T)v[i0], v1 = (T)v[i1], m0 = (T)d.spl_m[i0], m1 = (T)d.spl_m[i1];
      sv = h00 * v0 + h10 * m0 * dx + h01 * v1 + h11 * m1 * dx;
      I = c.k0 * M::exp(LN10 * sv);
      if (GRAD) {
        dsdR = ((-6 * t + 6 * t2) * v0 + (1 - 4 * t + 3 * t2) * m0 * dx + (6 * t - 6 * t2) * v1 +
                (-2 * t + 3 * t2) * m1 * dx) / dx;
        const T g = LN10 * I;
        dI[4 + i0] += g * h00;
        dI[4 + i1] += g * h01;
        // slopes: m_0 = D_0, m_k = (D_{k-1}+D_k)/2, m_{K-1} = D_{K-2},  D_k = (v_{k+1}-v_k)/h_k
        const int ms[2] = {i0, i1};
        const T mw_[2] = {g * h10 * dx, g * h11 * dx};
        for (int q = 0; q < 2; ++q) {
          const int m = ms[q];
          const T w = mw_[q];
          if (m == 0) {
            const T ih = (T)(1.0 / (s.prof[1] - s.prof[0]));
            dI[4 + 1] += w * ih;
            dI[4 + 0] -= w * ih;
          } else if (m == K - 1) {
            const T ih = (T)(1.0 / (s.prof[K - 1] - s.prof[K - 2]));
            dI[4 + K - 1] += w * ih;
            dI[4 + K - 2] -= w * ih;
          } else {
            const T ia = (T)(0.5 / (s.prof[m] - s.prof[m - 1]));
            const T ib = (T)(0.5 / (s.prof[m + 1] - s.prof[m]));
            dI[4 + m] += w * (ia - ib);
            dI[4 + m - 1] -= w * ia;
            dI[4 + m + 1] += w * ib;
          }
        }
      }
    }
    if (GRAD) dIdR_over_R = LN10 * I * dsdR / R;
  }
  if (GRAD) {
    // dR/dX * R = xp*c + yp*s/q ; dR/dY * R = -xp*s + yp*c/q
    if (c.radial) {
      dI[0] = -dIdR_over_R * xp;
      dI[1] = -dIdR_over_R * yp;
      dI[2] = T(0.0);
      dI[3] = T(0.0);
    } else {
      dI[0] = -dIdR_over_R * (xp * c.c + yp * c.s * c.qinv);
      dI[1] = -dIdR_over_R * (-xp * c.s + yp * c.c * c.qinv);
      dI[2] = dIdR_over_R * (-(yp * yp) * c.qinv);
      dI[3] = dIdR_over_R * (xp * yp * (c.q - c.qinv));
    }
  }
  return I;
}

// Evaluate brightness I at plane offset (X, Y) from the centre, scaled by `ascale`
// (sub-pixel area factor), and optionally dI/d(element) for every element.
// NE = compile-time element bound for the kind.  dI must hold n_elem doubles.  T: arithmetic of the evaluation.
template <int KIND, bool GRAD, typename T = double>
__device__ __forceinline__ double eval_point(const DevSrc& s, const DevDyn& d, double X, double Y, double ascale,
                                             double* __restrict__ dI) {
  if constexpr (KIND == APB_PLANE_SKY) {
    // planesky_model.py:65-74:  pixel_area F + X dx + Y dy  (natural flux units; only ever sampled at pixel centres)
    if (GRAD) {
      dI[0] = -d.el[3] * ascale;
      dI[1] = -d.el[4] * ascale;
      dI[2] = s.area * ascale;
      dI[3] = X * ascale;
      dI[4] = Y * ascale;
    }
    return ascale * (s.area * d.el[2] + X * d.el[3] + Y * d.el[4]);
  } else {
    PCtxT<T> c;
    pctx_load(c, s, d, ascale);
    T xp, yp;
    rot_coords<T>(c, (T)X, (T)Y, xp, yp);
    if constexpr (sizeof(T) == sizeof(double)) {
      return eval_rot<KIND, GRAD, T>(c, s, d, xp, yp, (T*)dI);
    } else {
      T g[GRAD ? KindInfo<KIND>::NE : 1];
      const T I = eval_rot<KIND, GRAD, T>(c, s, d, xp, yp, g);
      if (GRAD) {
        const int ne = (KIND == APB_SPLINE) ? s.n_elem : KindInfo<KIND>::NE;
        for (int e = 0; e < ne; ++e) dI[e] = (double)g[e];
      }
      return (double)I;
    }
  }
}

// deterministic block reduction (fixed tree), result valid in thread 0
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    r = lane < NT / 32 ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;
}
