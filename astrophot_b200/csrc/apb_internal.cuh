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

// Per-source values of a profile evaluation, held in registers while a thread works on one source (the kernels used to
// re-read them from the DevSrc / DevDyn tables at every node: a third of the instructions of an evaluation).
struct PCtx {
  double c, s, qinv, soft2;   // rotation by -(PA - pi/2), 1/q, softening^2
  double k0;                  // amplitude constant times the sub-pixel area factor of the cell being integrated
  double k1, k2, k3, k4, k5, k6;
  double q;                   // axis ratio (d/dPA)
  int radial;
};
__device__ __forceinline__ void pctx_load(PCtx& c, const DevSrc& s, const DevDyn& d, double ascale) {
  c.c = d.c; c.s = d.s; c.qinv = d.qinv; c.soft2 = s.soft2;
  c.k0 = ascale * d.k[0];
  c.k1 = d.k[1]; c.k2 = d.k[2]; c.k3 = d.k[3]; c.k4 = d.k[4]; c.k5 = d.k[5]; c.k6 = d.k[6];
  c.q = d.el[2];
  c.radial = s.flags & APB_F_RADIAL;
}

// softened squared radius: ONE expression for every kernel, so that all of them see the same bits
__device__ __forceinline__ double r2_of(double xp, double yp, double soft2) { return fma(xp, xp, fma(yp, yp, soft2)); }

// rotated, axis-ratio-scaled offsets (_shared_methods.py:274-307, coordinates.py:5-13).  PSF models are circular: k_prep
// gives them c = 1, s = 0, 1/q = 1, for which these products are exact.
__device__ __forceinline__ void rot_coords(const PCtx& c, double X, double Y, double& xp, double& yp) {
  xp = c.c * X - c.s * Y;
  yp = (c.s * X + c.c * Y) * c.qinv;
}

// Value of an analytic profile at squared radius R2 WITHOUT range checks in exp / log: `bad` is raised instead when an
// argument leaves their fast range (then the value is garbage and the caller redoes it with eval_rot).  Three of these
// run side by side per thread in the integration kernels; one predicate and one branch serve all of them.
template <int KIND>
__device__ __forceinline__ double prof_fast(const PCtx& c, double R2, bool& bad) {
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

// Brightness at rotated offsets (xp, yp) -- c.k0 carries the sub-pixel area factor -- and optionally dI/d(element) for
// every element (natural units).  dI must hold the kind's element count.
template <int KIND, bool GRAD>
__device__ __forceinline__ double eval_rot(const PCtx& c, const DevSrc& s, const DevDyn& d, double xp, double yp,
                                           double* __restrict__ dI) {
  const double R2 = r2_of(xp, yp, c.soft2);
  double I, dIdR_over_R;  // (dI/dR)/R : multiply by xp, yp pieces to get dI/dX
  if (KIND == APB_SERSIC) {
    // k0 = area*10^Ie, k1 = 1/Re^2, k2 = 1/(2n), k3 = b_n, k4 = b'_n, k5 = 1/n, k6 = 1/Re
    const double L2 = apb_log(R2 * c.k1, s_mathtab);  // 2 ln(R/Re)
    const double u = apb_exp(L2 * c.k2, s_mathtab);
    I = c.k0 * apb_exp(fma(-c.k3, u, c.k3), s_mathtab);
    if (GRAD) {
      const double bu = c.k3 * u;
      dIdR_over_R = -I * bu * c.k5 * apb_rcp(R2);
      dI[4] = I * (-c.k4 * (u - 1.0) + bu * (0.5 * L2) * c.k5 * c.k5);
      dI[5] = I * bu * c.k5 * c.k6;
      dI[6] = APB_LN10 * I;
    }
  } else if (KIND == APB_EXPONENTIAL) {
    // k0 = area*10^Ie, k1 = 1/Re, k2 = b_1
    const double R = sqrt(R2);
    I = c.k0 * apb_exp(-c.k2 * (R * c.k1 - 1.0), s_mathtab);
    if (GRAD) {
      dIdR_over_R = -I * c.k2 * c.k1 / R;
      dI[4] = I * c.k2 * R * c.k1 * c.k1;
      dI[5] = APB_LN10 * I;
    }
  } else if (KIND == APB_GAUSSIAN) {
    // k0 = area*10^flux/sqrt(2 pi sigma^2), k1 = 1/sigma^2, k2 = 1/sigma
    I = c.k0 * apb_exp(-0.5 * R2 * c.k1, s_mathtab);
    if (GRAD) {
      dIdR_over_R = -I * c.k1;
      dI[4] = I * (-c.k2 + R2 * c.k1 * c.k2);
      dI[5] = APB_LN10 * I;
    }
  } else if (KIND == APB_MOFFAT) {
    // k0 = area*10^I0, k1 = 1/Rd^2, k2 = n, k3 = 1/Rd
    const double t = 1.0 + R2 * c.k1;
    const double lt = apb_log(t, s_mathtab);
    I = c.k0 * apb_exp(-c.k2 * lt, s_mathtab);
    if (GRAD) {
      dIdR_over_R = -I * c.k2 * 2.0 * c.k1 / t;
      dI[4] = -I * lt;
      dI[5] = I * c.k2 * 2.0 * R2 * c.k1 * c.k3 / t;
      dI[6] = APB_LN10 * I;
    }
  } else {  // APB_SPLINE: k0 = area
    const double R = sqrt(R2);
    const int K = s.n_prof;
    // idx = searchsorted(prof[:-1], R, left) - 1, wrapping -1 -> K-1 (utils/interpolate.py:53)
    int idx = -1;
    for (int k = 0; k < K - 1; ++k)
      if (s.prof[k] < R) idx = k;
    const double* v = d.el + 4;
    double sv, dsdR;
    if (GRAD)
      for (int k = 0; k < K; ++k) dI[4 + k] = 0.0;
    if (R > s.prof[K - 1]) {
      const double h = s.prof[K - 1] - s.prof[K - 2];
      const double f = (R - s.prof[K - 2]) / h;
      sv = v[K - 2] + (R - s.prof[K - 2]) * ((v[K - 1] - v[K - 2]) / h);
      dsdR = (v[K - 1] - v[K - 2]) / h;
      I = c.k0 * apb_exp(APB_LN10 * sv, s_mathtab);
      if (GRAD) {
        dI[4 + K - 2] = APB_LN10 * I * (1.0 - f);
        dI[4 + K - 1] = APB_LN10 * I * f;
      }
    } else {
      const int i0 = idx < 0 ? K - 1 : idx;
      const int i1 = idx + 1;
      const double dx = s.prof[i1] - s.prof[i0];
      const double t = (R - s.prof[i0]) / dx;
      const double t2 = t * t, t3 = t2 * t;
      const double h00 = 1 - 3 * t2 + 2 * t3, h10 = t - 2 * t2 + t3, h01 = 3 * t2 - 2 * t3, h11 = t3 - t2;
      sv = h00 * v[i0] + h10 * d.spl_m[i0] * dx + h01 * v[i1] + h11 * d.spl_m[i1] * dx;
      I = c.k0 * apb_exp(APB_LN10 * sv, s_mathtab);
      if (GRAD) {
        dsdR = ((-6 * t + 6 * t2) * v[i0] + (1 - 4 * t + 3 * t2) * d.spl_m[i0] * dx + (6 * t - 6 * t2) * v[i1] +
                (-2 * t + 3 * t2) * d.spl_m[i1] * dx) / dx;
        const double g = APB_LN10 * I;
        dI[4 + i0] += g * h00;
        dI[4 + i1] += g * h01;
        // slopes: m_0 = D_0, m_k = (D_{k-1}+D_k)/2, m_{K-1} = D_{K-2},  D_k = (v_{k+1}-v_k)/h_k
        const int ms[2] = {i0, i1};
        const double mw_[2] = {g * h10 * dx, g * h11 * dx};
        for (int q = 0; q < 2; ++q) {
          const int m = ms[q];
          const double w = mw_[q];
          if (m == 0) {
            const double ih = 1.0 / (s.prof[1] - s.prof[0]);
            dI[4 + 1] += w * ih;
            dI[4 + 0] -= w * ih;
          } else if (m == K - 1) {
            const double ih = 1.0 / (s.prof[K - 1] - s.prof[K - 2]);
            dI[4 + K - 1] += w * ih;
            dI[4 + K - 2] -= w * ih;
          } else {
            const double ia = 0.5 / (s.prof[m] - s.prof[m - 1]);
            const double ib = 0.5 / (s.prof[m + 1] - s.prof[m]);
            dI[4 + m] += w * (ia - ib);
            dI[4 + m - 1] -= w * ia;
            dI[4 + m + 1] += w * ib;
          }
        }
      }
    }
    if (GRAD) dIdR_over_R = APB_LN10 * I * dsdR / R;
  }
  if (GRAD) {
    // dR/dX * R = xp*c + yp*s/q ; dR/dY * R = -xp*s + yp*c/q
    if (c.radial) {
      dI[0] = -dIdR_over_R * xp;
      dI[1] = -dIdR_over_R * yp;
      dI[2] = 0.0;
      dI[3] = 0.0;
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
// NE = compile-time element bound for the kind.  dI must hold n_elem doubles.
template <int KIND, bool GRAD>
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
    PCtx c;
    pctx_load(c, s, d, ascale);
    double xp, yp;
    rot_coords(c, X, Y, xp, yp);
    return eval_rot<KIND, GRAD>(c, s, d, xp, yp, dI);
  }
}

template <int KIND>
struct KindInfo {};
template <> struct KindInfo<APB_SERSIC> { static constexpr int NE = 7; };
template <> struct KindInfo<APB_PLANE_SKY> { static constexpr int NE = 5; };
template <> struct KindInfo<APB_EXPONENTIAL> { static constexpr int NE = 6; };
template <> struct KindInfo<APB_GAUSSIAN> { static constexpr int NE = 6; };
template <> struct KindInfo<APB_MOFFAT> { static constexpr int NE = 7; };
template <> struct KindInfo<APB_SPLINE> { static constexpr int NE = APB_MAX_ELEM; };

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
