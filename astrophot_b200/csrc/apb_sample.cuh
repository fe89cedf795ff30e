// Sampling kernels: parameter prep, first pass, curvature select + ballot compaction,
// work-queue refinement, reduction, scatter, PSF-model normalisation.
#pragma once
#include <cooperative_groups.h>
#include "apb_internal.cuh"

// work queues of the adaptive integration (one set per refinement depth)
struct Level {
  int* src;      // source of each entry
  double* x;     // plane offset from the centre
  double* y;
  int* parent;   // depth 1: pixel index in the stamp; deeper: entry index one level up
  int* child;    // first child entry one level down, or -1
  double* res;   // [entry * NVp + plane]
};

struct Queues {
  Level lv[APB_MAX_DEPTH + 1];  // index by depth (1-based)
  int* count;                   // [APB_MAX_DEPTH + 2] entries per depth, device
  int* overflow;                // device flag
  unsigned long long* cum;      // [2][APB_MAX_DEPTH + 2] entries per depth summed over all earlier passes (value-only / derivative)
  int* last_kind;               // 0 / 1: kind of the pass whose counts `count` holds
  int cap[APB_MAX_DEPTH + 2];   // entries allocated per depth
  int NVp;                      // planes carried per entry in this pass
};

// ----------------------------------------------------------------------------
// prep: x -> per-source natural values, chain factors, constants
// ----------------------------------------------------------------------------
// One warp per source: lanes transform the elements in parallel, then five lanes compute the
// independent groups of constants (each a serial chain of fp64 transcendentals) side by side.
// Every sampling pass waits on this kernel, so its latency, not its throughput, is what counts.
__global__ void __launch_bounds__(128) k_prep(const DevSrc* __restrict__ src, DevDyn* __restrict__ dyn, int n_src,
                                              const apb_param_t* __restrict__ par, const double* __restrict__ x, int as_rep,
                                              int* qcount, double* skyJ, int write_skyJ,
                                              unsigned long long* qcum, int* qlast) {
  const int si = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (si == 0 && qcount) {
    // the counts of the previous pass move into the running totals (apb_plan_stats), then the queues are empty again
    const int prev = *qlast;
    __syncwarp();
    if (lane < APB_MAX_DEPTH + 2) {
      qcum[prev * (APB_MAX_DEPTH + 2) + lane] += (unsigned int)qcount[lane];
      qcount[lane] = 0;
    }
    if (lane == 0) *qlast = write_skyJ ? 1 : 0;
  }
  if (si >= n_src) return;
  const DevSrc& s = src[si];
  DevDyn& d = dyn[si];
  // (the amplitude pseudo-element of an APB_F_AMP source sits right after the profile's elements)
  for (int e = lane; e < s.n_elem + (s.amp_elem >= 0 ? 1 : 0); e += 32) {
    const int sl = s.slot[e];
    double v = s.cval[e], ch = 0.0;
    if (sl >= 0) {
      const double r = x[sl];
      v = r;
      ch = 1.0;
      if (as_rep) {
        const apb_param_t p = par[sl];
        if (p.transform == APB_TR_LOWER) {
          const double dd = r - p.lo, rt = sqrt(dd * dd + 4.0);
          v = 0.5 * (r + p.lo + rt);
          ch = 0.5 + 0.5 * dd / rt;
        } else if (p.transform == APB_TR_UPPER) {
          const double dd = r - p.hi, rt = sqrt(dd * dd + 4.0);
          v = 0.5 * (r + p.hi - rt);
          ch = 0.5 - 0.5 * dd / rt;
        } else if (p.transform == APB_TR_BOTH) {
          v = (atan(r) + APB_PI / 2) * (p.hi - p.lo) / APB_PI + p.lo;
          ch = (p.hi - p.lo) / (APB_PI * (r * r + 1.0));
        } else if (p.transform == APB_TR_CYCLIC) {
          const double per = p.hi - p.lo;
          double m = fmod(r - p.lo, per);
          if (m < 0) m += per;
          v = p.lo + m;
        }
      }
    }
    d.el[e] = v;
    d.chain[e] = ch;
  }
  __syncwarp();
  const int kind = s.kind;
  const bool profile = kind != APB_FLAT_SKY && kind != APB_POINT;
  if (lane == 0) {
    // pixel position of the centre, sub-pixel shift (model_object.py:320-323, point_source.py:149-150)
    const double cx = d.el[0], cy = d.el[1];
    const double pcx = s.Sinv[0] * (cx - s.rxy[0]) + s.Sinv[1] * (cy - s.rxy[1]) + s.rij[0];
    const double pcy = s.Sinv[2] * (cx - s.rxy[0]) + s.Sinv[3] * (cy - s.rxy[1]) + s.rij[1];
    const double rx = rint(pcx), ry = rint(pcy);
    d.rx = (int)rx;
    d.ry = (int)ry;
    d.sx = pcx - rx;
    d.sy = pcy - ry;
  } else if (lane == 1) {
    double sn = 0.0, cs = 1.0, qi = 1.0;
    if (profile && !(s.flags & APB_F_RADIAL)) {
      sincos(-(d.el[3] - APB_PI / 2), &sn, &cs);
      qi = 1.0 / d.el[2];
    }
    d.c = cs;
    d.s = sn;
    d.qinv = qi;
  } else if (lane == 2) {
    // amplitude
    if (kind == APB_FLAT_SKY) {
      const double a = s.area * exp10(d.el[2]);
      d.k[0] = a;
      if (write_skyJ) skyJ[si] = APB_LN10 * a * d.chain[2];
    } else if (kind == APB_POINT) {
      d.k[0] = exp10(d.el[2]);
    } else if (kind == APB_SERSIC || kind == APB_MOFFAT) {
      d.k[0] = s.area * exp10(d.el[6]);
    } else if (kind == APB_EXPONENTIAL) {
      d.k[0] = s.area * exp10(d.el[5]);
    } else if (kind == APB_GAUSSIAN) {
      const double sg = d.el[4];
      d.k[0] = s.area * exp10(d.el[5]) / sqrt(2 * APB_PI * sg * sg);
    } else {
      d.k[0] = s.area;
    }
  } else if (lane == 3) {
    double t0 = 0.0, t1 = 0.0;
    if (kind == APB_SERSIC && s.ref_mode == APB_REF_SERSIC_FLUX) {
      // total flux / numel of the working image (sersic_model.py:87-89, conversions/functions.py:168-190)
      const double n = d.el[4], Re = d.el[5], bn = sersic_b(n);
      const double flux = 2 * APB_PI * exp10(d.el[6]) * Re * Re * d.el[2] * n * (exp(bn) * pow(bn, -2 * n)) * exp(lgamma(2 * n));
      t0 = s.tol * (flux / ((double)s.geo[0].rw * (double)s.geo[0].rh));
      t1 = s.tol * (flux / ((double)s.geo[1].rw * (double)s.geo[1].rh));
    }
    d.thr[0] = t0;
    d.thr[1] = t1;
  } else if (lane == 4) {
    if (kind == APB_SERSIC) {
      const double n = d.el[4], Re = d.el[5];
      d.k[1] = 1.0 / (Re * Re);
      d.k[2] = 0.5 / n;
      d.k[3] = sersic_b(n);
      d.k[4] = sersic_db(n);
      d.k[5] = 1.0 / n;
      d.k[6] = 1.0 / Re;
    } else if (kind == APB_EXPONENTIAL) {
      d.k[1] = 1.0 / d.el[4];
      d.k[2] = sersic_b(1.0);
    } else if (kind == APB_GAUSSIAN) {
      const double sg = d.el[4];
      d.k[1] = 1.0 / (sg * sg);
      d.k[2] = 1.0 / sg;
    } else if (kind == APB_MOFFAT) {
      d.k[1] = 1.0 / (d.el[5] * d.el[5]);
      d.k[2] = d.el[4];
      d.k[3] = 1.0 / d.el[5];
    } else if (kind == APB_SPLINE) {
      const int K = s.n_prof;
      const double* v = d.el + 4;
      for (int k = 0; k < K; ++k) {
        double m;
        if (k == 0) m = (v[1] - v[0]) / (s.prof[1] - s.prof[0]);
        else if (k == K - 1) m = (v[K - 1] - v[K - 2]) / (s.prof[K - 1] - s.prof[K - 2]);
        else m = ((v[k] - v[k - 1]) / (s.prof[k] - s.prof[k - 1]) + (v[k + 1] - v[k]) / (s.prof[k + 1] - s.prof[k])) / 2;
        d.spl_m[k] = m;
      }
    }
  } else if (lane == 5) {
    // Mean reference over a working region much larger than the source (a model inside a big group window,
    // _model_methods.py:151-152 with group_model_object.py:211-227): beyond rcut every pixel is below 1e-22 / N_pix of
    // the central value, i.e. far below the rounding of the sum, and is skipped by k_mean_partial.  Elliptical radius
    // >= Euclidean distance (q <= 1), so a Euclidean cut is conservative.  Power-law profiles get no cut.
    double rc = INFINITY;
    if (profile && s.integrate_mode == APB_INTEGRATE_THRESHOLD && s.ref_mode == APB_REF_MEAN &&
        (s.flags & APB_F_RADIAL ? true : d.el[2] <= 1.0)) {
      const double npx = fmax((double)s.geo[0].rw * s.geo[0].rh, (double)s.geo[1].rw * s.geo[1].rh);
      const double lg = log(npx * 1e22);
      if (kind == APB_EXPONENTIAL) rc = d.el[4] * lg / sersic_b(1.0);
      else if (kind == APB_GAUSSIAN) rc = d.el[4] * sqrt(2.0 * lg);
      else if (kind == APB_SPLINE) {
        const int K = s.n_prof;
        const double* v = d.el + 4;
        const double m = (v[K - 1] - v[K - 2]) / (s.prof[K - 1] - s.prof[K - 2]);   // log10 per unit radius beyond the last node
        if (m < 0.0) rc = s.prof[K - 1] + fmax(0.0, (-lg / APB_LN10 + fmin(v[0], v[1]) - v[K - 1]) / m);
      }
    }
    d.rcut = rc;
  }
}

// plane coordinates (relative to the centre) of image pixel (pi, pj)
__device__ __forceinline__ void pix_coords(const DevSrc& s, const DevDyn& d, double pi, double pj, double& X,
                                           double& Y) {
  if (s.psf >= 0 && s.psf_shift != APB_SHIFT_NONE) {
    // grid re-centred on the source: X = S.(pix - round(pc))   (model_object.py:320-323)
    const double di = pi - d.rx, dj = pj - d.ry;
    X = s.S[0] * di + s.S[1] * dj;
    Y = s.S[2] * di + s.S[3] * dj;
  } else {
    const double di = pi - s.rij[0], dj = pj - s.rij[1];
    X = (s.S[0] * di + s.S[1] * dj + s.rxy[0]) - d.el[0];
    Y = (s.S[2] * di + s.S[3] * dj + s.rxy[1]) - d.el[1];
  }
}

template <bool GRAD, int NE>
struct Acc {
  double v[(GRAD ? NE : 0) + 1];
};

// Gauss-Legendre nodes k0, k0 + kstep, ... of the n x n rule (n = s.quad_level unless given) over the cell centred
// (X, Y) whose edges are the pixel's times `scale`; acc.v[0] = weighted sum, acc.v[1+e] = derivative wrt element e
// (natural units); returns the value at the centre node (0 for lanes that do not own it).
// The rotation (_shared_methods.py:274-307) is linear, so it is applied once to the cell centre and once to the two
// pixel edges; node (ax, ay) (pixel units) then sits at
//     xp = fma(ax, uxi, fma(ay, uxj, xp0)),   yp = fma(ax, vyi, fma(ay, vyj, yp0))
// -- the SAME expressions in every kernel and for every split of the nodes over lanes, so that all integration kernels
// see bit-identical node values (their refinement decisions must agree).
template <int KIND, bool GRAD, typename T = double>
__device__ __forceinline__ double gl_nodes_n(const DevSrc& s, const DevDyn& d, double X, double Y, int n, double scale,
                                             double ascale, int k0, int kstep, Acc<GRAD, KindInfo<KIND>::NE>& acc) {
  constexpr int NE = KindInfo<KIND>::NE;
  const int ne = (KIND == APB_SPLINE) ? s.n_elem : NE;
  PCtxT<T> c;
  pctx_load(c, s, d, ascale);
  T xp0, yp0, uxi, vyi, uxj, vyj;
  rot_coords<T>(c, (T)X, (T)Y, xp0, yp0);
  rot_coords<T>(c, (T)(s.S[0] * scale), (T)(s.S[2] * scale), uxi, vyi);
  rot_coords<T>(c, (T)(s.S[1] * scale), (T)(s.S[3] * scale), uxj, vyj);
  const int nn = n * n, mid = nn / 2;
  if constexpr (!GRAD && KIND != APB_SPLINE) {
    if (n == 3 && kstep == 1) {
      // the default 3x3 rule, all nodes on one lane: a row of three at a time with unchecked exp / log, so that three
      // independent chains are in flight and one range test serves the row (kept rolled: three inlined evaluations
      // are ~250 instructions; nine would crowd the other phases of the integration kernels out of the instruction cache)
      const T a0 = (T)c_quad.a[3][0], a2 = (T)c_quad.a[3][2];          // a[3][1] == 0
      const T w0 = (T)c_quad.w[3][0], w1 = (T)c_quad.w[3][1], w2 = (T)c_quad.w[3][2];
      T tot = T(0.0), centre = T(0.0);
#pragma unroll 1
      for (int ky = 0; ky < 3; ++ky) {
        const T ay = (T)c_quad.a[3][ky];
        const T tx = MT<T>::fma(ay, uxj, xp0), ty = MT<T>::fma(ay, vyj, yp0);
        const T xa = MT<T>::fma(a0, uxi, tx), ya = MT<T>::fma(a0, vyi, ty);
        const T xb = MT<T>::fma(a2, uxi, tx), yb = MT<T>::fma(a2, vyi, ty);
        bool bad = false;
        T Ia = prof_fast<KIND>(c, r2_of<T>(xa, ya, c.soft2), bad);
        T Im = prof_fast<KIND>(c, r2_of<T>(tx, ty, c.soft2), bad);
        T Ib = prof_fast<KIND>(c, r2_of<T>(xb, yb, c.soft2), bad);
        if (bad) {   // an argument outside the fast range of exp / log (extreme or non-finite parameters)
          Ia = eval_rot<KIND, false, T>(c, s, d, xa, ya, nullptr);
          Im = eval_rot<KIND, false, T>(c, s, d, tx, ty, nullptr);
          Ib = eval_rot<KIND, false, T>(c, s, d, xb, yb, nullptr);
        }
        if (ky == 1) centre = Im;
        tot = MT<T>::fma((T)c_quad.w[3][ky], MT<T>::fma(w2, Ib, MT<T>::fma(w1, Im, w0 * Ia)), tot);
      }
      acc.v[0] = (double)tot;
      return (double)centre;
    }
  }
  T av[(GRAD ? NE : 0) + 1];
  av[0] = T(0.0);
  if (GRAD)
    for (int e = 0; e < ne; ++e) av[1 + e] = T(0.0);
  T centre = T(0.0);
  T dI[GRAD ? NE : 1];
  for (int k = k0; k < nn; k += kstep) {
    const int kx = k % n, ky = k / n;
    const T ax = (T)c_quad.a[n][kx], ay = (T)c_quad.a[n][ky];
    const T w = (T)c_quad.w[n][kx] * (T)c_quad.w[n][ky];
    const T xp = MT<T>::fma(ax, uxi, MT<T>::fma(ay, uxj, xp0)), yp = MT<T>::fma(ax, vyi, MT<T>::fma(ay, vyj, yp0));
    const T I = eval_rot<KIND, GRAD, T>(c, s, d, xp, yp, dI);
    if (k == mid) centre = I;
    av[0] += I * w;
    if (GRAD)
      for (int e = 0; e < ne; ++e) av[1 + e] += dI[e] * w;
  }
  acc.v[0] = (double)av[0];
  if (GRAD)
    for (int e = 0; e < ne; ++e) acc.v[1 + e] = (double)av[1 + e];
  return (double)centre;
}

template <int KIND, bool GRAD, typename T = double>
__device__ __forceinline__ double gl_nodes(const DevSrc& s, const DevDyn& d, double X, double Y, double scale, double ascale,
                                           int k0, int kstep, Acc<GRAD, KindInfo<KIND>::NE>& acc) {
  return gl_nodes_n<KIND, GRAD, T>(s, d, X, Y, s.quad_level, scale, ascale, k0, kstep, acc);
}

// Gauss-Legendre n x n over one (sub)pixel, all nodes on the calling thread.
// acc[0] = integral, acc[1+e] = derivative wrt element e (natural units); returns the centre node.
template <int KIND, bool GRAD, typename T = double>
__device__ __forceinline__ double gl_integrate(const DevSrc& s, const DevDyn& d, double X, double Y, int n,
                                               double scale, double ascale, double* __restrict__ acc) {
  constexpr int NE = KindInfo<KIND>::NE;
  const int ne = (KIND == APB_SPLINE) ? s.n_elem : NE;
  Acc<GRAD, NE> a;
  const double centre = gl_nodes_n<KIND, GRAD, T>(s, d, X, Y, n, scale, ascale, 0, 1, a);
  acc[0] = a.v[0];
  if (GRAD)
    for (int e = 0; e < ne; ++e) acc[1 + e] = a.v[1 + e];
  return centre;
}

// ----------------------------------------------------------------------------
// first pass over the stamp region (one 32x32 tile per CTA, four rows per thread)
// ----------------------------------------------------------------------------
template <int KIND, bool GRAD, typename T = double>
__device__ __forceinline__ void first_pass_pixel(const DevSrc& s, const DevDyn& d, const Geo& g, int i, int j,
                                                 double* __restrict__ stamp, int err_plane) {
  constexpr int NE = KindInfo<KIND>::NE;
  const int ne = (KIND == APB_SPLINE) ? s.n_elem : NE;
  const int pi = g.mx0 + i, pj = g.my0 + j;
  double X, Y;
  pix_coords(s, d, (double)pi, (double)pj, X, Y);
  double* base = stamp + s.stamp_off + (long long)j * g.mw + i;
  const bool in_e = pi >= g.ex0 && pi < g.ex0 + g.ew && pj >= g.ey0 && pj < g.ey0 + g.eh;
  double acc[(GRAD ? NE : 0) + 1];
  if (s.sampling_mode == APB_SAMPLE_MIDPOINT) {
    acc[0] = eval_point<KIND, GRAD, T>(s, d, X, Y, 1.0, acc + 1);
  } else if (s.sampling_mode == APB_SAMPLE_TRAPEZOID) {
    // mean of the four pixel-corner values (_model_methods.py:124-143); error proxy = curvature, as for midpoint
    double dI[GRAD ? NE : 1];
    acc[0] = 0.0;
    if (GRAD)
      for (int e = 0; e < ne; ++e) acc[1 + e] = 0.0;
    for (int a = -1; a <= 1; a += 2)
      for (int b = -1; b <= 1; b += 2) {
        const double ox = 0.5 * b, oy = 0.5 * a;
        const double I = eval_point<KIND, GRAD, T>(s, d, X + (s.S[0] * ox + s.S[1] * oy), Y + (s.S[2] * ox + s.S[3] * oy),
                                                1.0, dI);
        acc[0] += 0.25 * I;
        if (GRAD)
          for (int e = 0; e < ne; ++e) acc[1 + e] += 0.25 * dI[e];
      }
  } else if (s.sampling_mode == APB_SAMPLE_QUAD) {
    const double centre = gl_integrate<KIND, GRAD, T>(s, d, X, Y, s.quad_init, 1.0, 1.0, acc);
    base[(long long)err_plane * s.plane_stride] = fabs(acc[0] - centre);
  } else {  // simpsons: 3x3 half-pixel lattice, weights [1 4 1; 4 16 4; 1 4 1]/36 (_model_methods.py:99-109)
    double dI[GRAD ? NE : 1];
    acc[0] = 0.0;
    if (GRAD)
      for (int e = 0; e < ne; ++e) acc[1 + e] = 0.0;
    double midv = 0.0;
    for (int a = -1; a <= 1; ++a)
      for (int b = -1; b <= 1; ++b) {
        const double w = ((a == 0 ? 4.0 : 1.0) * (b == 0 ? 4.0 : 1.0)) / 36.0;
        const double ox = 0.5 * b, oy = 0.5 * a;
        const double I = eval_point<KIND, GRAD, T>(s, d, X + (s.S[0] * ox + s.S[1] * oy), Y + (s.S[2] * ox + s.S[3] * oy),
                                                1.0, dI);
        if (a == 0 && b == 0) midv = I;
        acc[0] += w * I;
        if (GRAD)
          for (int e = 0; e < ne; ++e) acc[1 + e] += w * dI[e];
      }
    base[(long long)err_plane * s.plane_stride] = fabs(acc[0] - midv);
  }
  base[0] = acc[0];
  if (GRAD && in_e) {
    for (int e = 0; e < ne; ++e) {
      const int p = s.plane[e];
      if (p > 0) base[(long long)p * s.plane_stride] = acc[1 + e] * d.chain[e];
    }
  }
}

// Four midpoint evaluations of one thread (rows ly, ly+8, ly+16, ly+24 of a 32x32 tile) side by side: one register
// context, four independent exp / log chains, one range test.  Same arithmetic as eval_point (bit-identical values:
// the mean reference may re-evaluate pixels through that path).
#define FIRST_ROWS 4
template <int KIND, typename T>
__device__ __forceinline__ void first_pass_fast(const DevSrc& s, const DevDyn& d, const Geo& g, int i, int j0,
                                                double* __restrict__ stamp) {
  PCtxT<T> c;
  pctx_load(c, s, d, 1.0);
  T xp[FIRST_ROWS], yp[FIRST_ROWS], I[FIRST_ROWS];
  bool bad = false;
#pragma unroll
  for (int r = 0; r < FIRST_ROWS; ++r) {
    double X, Y;
    pix_coords(s, d, (double)(g.mx0 + i), (double)(g.my0 + j0 + 8 * r), X, Y);
    rot_coords<T>(c, (T)X, (T)Y, xp[r], yp[r]);
  }
#pragma unroll
  for (int r = 0; r < FIRST_ROWS; ++r) I[r] = prof_fast<KIND>(c, r2_of<T>(xp[r], yp[r], c.soft2), bad);
  if (bad) {
#pragma unroll
    for (int r = 0; r < FIRST_ROWS; ++r) I[r] = eval_rot<KIND, false, T>(c, s, d, xp[r], yp[r], nullptr);
  }
  double* base = stamp + s.stamp_off + (long long)j0 * g.mw + i;
#pragma unroll
  for (int r = 0; r < FIRST_ROWS; ++r)
    if (j0 + 8 * r < g.mh) base[(long long)(8 * r) * g.mw] = (double)I[r];
}

// tiles are 32 x 32 pixels of the stamp region; a thread owns four rows of one column
template <bool GRAD, typename T>
__global__ void __launch_bounds__(256, GRAD ? 2 : 4) k_first(const DevSrc* __restrict__ src, const DevDyn* __restrict__ dyn,
                                               const int4* __restrict__ tiles, int mode, double* __restrict__ stamp,
                                               int err_plane_unused) {
  apb_math_load();
  const int4 t = tiles[blockIdx.x];
  const DevSrc& s = src[t.x];
  const DevDyn& d = dyn[t.x];
  const Geo& g = s.geo[mode];
  const int i = t.y + (threadIdx.x & 31), j0 = t.z + (threadIdx.x >> 5);
  if (i >= g.mw || j0 >= g.mh) return;
  if constexpr (!GRAD) {
    if (s.sampling_mode == APB_SAMPLE_MIDPOINT) {
      switch (s.kind) {
        case APB_SERSIC: first_pass_fast<APB_SERSIC, T>(s, d, g, i, j0, stamp); return;
        case APB_EXPONENTIAL: first_pass_fast<APB_EXPONENTIAL, T>(s, d, g, i, j0, stamp); return;
        case APB_GAUSSIAN: first_pass_fast<APB_GAUSSIAN, T>(s, d, g, i, j0, stamp); return;
        case APB_MOFFAT: first_pass_fast<APB_MOFFAT, T>(s, d, g, i, j0, stamp); return;
        default: break;
      }
    }
  }
  const int errp = s.n_act + 1;
  for (int j = j0; j < g.mh && j < j0 + 8 * FIRST_ROWS; j += 8) {
    switch (s.kind) {
      case APB_SERSIC: first_pass_pixel<APB_SERSIC, GRAD, T>(s, d, g, i, j, stamp, errp); break;
      case APB_EXPONENTIAL: first_pass_pixel<APB_EXPONENTIAL, GRAD, T>(s, d, g, i, j, stamp, errp); break;
      case APB_GAUSSIAN: first_pass_pixel<APB_GAUSSIAN, GRAD, T>(s, d, g, i, j, stamp, errp); break;
      case APB_MOFFAT: first_pass_pixel<APB_MOFFAT, GRAD, T>(s, d, g, i, j, stamp, errp); break;
      case APB_SPLINE: first_pass_pixel<APB_SPLINE, GRAD, T>(s, d, g, i, j, stamp, errp); break;
      case APB_PLANE_SKY: first_pass_pixel<APB_PLANE_SKY, GRAD>(s, d, g, i, j, stamp, errp); break;
      default: break;
    }
  }
}

// ----------------------------------------------------------------------------
// mean reference (default _integrate_reference = mean of the first-pass image over the
// WORKING region, _model_methods.py:151-152).  Two stages, fixed order => deterministic.
// chunk list: {src, first pixel, n pixels, slot}
// ----------------------------------------------------------------------------
template <int KIND, typename T = double>
__device__ __forceinline__ double first_value(const DevSrc& s, const DevDyn& d, double X, double Y) {
  double acc[1];
  if (s.sampling_mode == APB_SAMPLE_MIDPOINT) return eval_point<KIND, false, T>(s, d, X, Y, 1.0, acc);
  if (s.sampling_mode == APB_SAMPLE_TRAPEZOID) {
    double tot = 0.0;
    for (int a = -1; a <= 1; a += 2)
      for (int b = -1; b <= 1; b += 2) {
        const double ox = 0.5 * b, oy = 0.5 * a;
        tot += 0.25 * eval_point<KIND, false, T>(s, d, X + (s.S[0] * ox + s.S[1] * oy), Y + (s.S[2] * ox + s.S[3] * oy), 1.0, acc);
      }
    return tot;
  }
  if (s.sampling_mode == APB_SAMPLE_QUAD) {
    gl_integrate<KIND, false, T>(s, d, X, Y, s.quad_init, 1.0, 1.0, acc);
    return acc[0];
  }
  double tot = 0.0;
  for (int a = -1; a <= 1; ++a)
    for (int b = -1; b <= 1; ++b) {
      const double w = ((a == 0 ? 4.0 : 1.0) * (b == 0 ? 4.0 : 1.0)) / 36.0;
      const double ox = 0.5 * b, oy = 0.5 * a;
      tot += w * eval_point<KIND, false, T>(s, d, X + (s.S[0] * ox + s.S[1] * oy), Y + (s.S[2] * ox + s.S[3] * oy), 1.0, acc);
    }
  return tot;
}

template <typename T>
__global__ void __launch_bounds__(256) k_mean_partial(const DevSrc* __restrict__ src, const DevDyn* __restrict__ dyn,
                                                      const int4* __restrict__ chunks, int mode,
                                                      const double* __restrict__ stamp, double* __restrict__ part) {
  __shared__ double sh[8];
  apb_math_load();
  const int4 c = chunks[blockIdx.x];
  const DevSrc& s = src[c.x];
  const DevDyn& d = dyn[c.x];
  const Geo& g = s.geo[mode];
  const bool from_stamp = (g.mx0 == g.rx0 && g.my0 == g.ry0 && g.mw == g.rw && g.mh == g.rh);
  // this chunk's rows, clipped to the box outside which the profile is negligible (k_prep: rcut) when the centre
  // lies inside the working region
  int i0 = 0, i1 = g.rw, j0 = c.y, j1 = c.y + c.z;
  if (!from_stamp && isfinite(d.rcut)) {
    const double cx = d.el[0], cy = d.el[1];
    const double pcx = s.Sinv[0] * (cx - s.rxy[0]) + s.Sinv[1] * (cy - s.rxy[1]) + s.rij[0];
    const double pcy = s.Sinv[2] * (cx - s.rxy[0]) + s.Sinv[3] * (cy - s.rxy[1]) + s.rij[1];
    if (pcx >= g.rx0 && pcx < g.rx0 + g.rw && pcy >= g.ry0 && pcy < g.ry0 + g.rh) {
      const double hx = d.rcut * sqrt(s.Sinv[0] * s.Sinv[0] + s.Sinv[1] * s.Sinv[1]) + 3.0;
      const double hy = d.rcut * sqrt(s.Sinv[2] * s.Sinv[2] + s.Sinv[3] * s.Sinv[3]) + 3.0;
      i0 = max(i0, (int)fmax(floor(pcx - hx) - g.rx0, -1.0e9));
      i1 = min(i1, (int)fmin(ceil(pcx + hx) - g.rx0 + 1.0, 1.0e9));
      j0 = max(j0, (int)fmax(floor(pcy - hy) - g.ry0, -1.0e9));
      j1 = min(j1, (int)fmin(ceil(pcy + hy) - g.ry0 + 1.0, 1.0e9));
    }
  }
  const int nc = max(i1 - i0, 0), nr = max(j1 - j0, 0);
  const long long npx = (long long)nc * nr;
  double v = 0.0;
  for (long long q = threadIdx.x; q < npx; q += 256) {
    const int i = i0 + (int)(q % nc), j = j0 + (int)(q / nc);
    if (from_stamp) {
      v += stamp[s.stamp_off + (long long)j * g.rw + i];
    } else {
      double X, Y;
      pix_coords(s, d, (double)(g.rx0 + i), (double)(g.ry0 + j), X, Y);
      switch (s.kind) {
        case APB_SERSIC: v += first_value<APB_SERSIC, T>(s, d, X, Y); break;
        case APB_EXPONENTIAL: v += first_value<APB_EXPONENTIAL, T>(s, d, X, Y); break;
        case APB_GAUSSIAN: v += first_value<APB_GAUSSIAN, T>(s, d, X, Y); break;
        case APB_MOFFAT: v += first_value<APB_MOFFAT, T>(s, d, X, Y); break;
        case APB_SPLINE: v += first_value<APB_SPLINE, T>(s, d, X, Y); break;
        default: break;
      }
    }
  }
  const double tot = block_sum<256>(v, sh);
  if (threadIdx.x == 0) part[c.w] = tot;
}

__global__ void k_mean_final(const DevSrc* __restrict__ src, DevDyn* __restrict__ dyn, const int* __restrict__ list,
                             int n, int mode, const double* __restrict__ part) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int si = list[q];
  const DevSrc& s = src[si];
  const Geo& g = s.geo[mode];
  double tot = 0.0;
  for (int k = 0; k < g.nchunk; ++k) tot += part[g.chunk0 + k];
  dyn[si].thr[mode] = s.tol * (tot / ((double)g.rw * (double)g.rh));
}

// ----------------------------------------------------------------------------
// select: curvature test + warp-ballot compaction into the depth-1 queue
// ----------------------------------------------------------------------------
__device__ __forceinline__ void select_tile(const DevSrc* __restrict__ src, const DevDyn* __restrict__ dyn, const int4 t, int mode,
                                            const double* __restrict__ stamp, const Queues& q) {
  const DevSrc& s = src[t.x];
  if (s.integrate_mode != APB_INTEGRATE_THRESHOLD) return;
  const DevDyn& d = dyn[t.x];
  const Geo& g = s.geo[mode];
  const int i = t.y + (threadIdx.x & 31);
  const int lane = threadIdx.x & 31;
  // tiles are 32 x 32: a warp walks rows ly, ly+8, ly+16, ly+24 (warp-uniform trip count: the ballots stay well defined)
#pragma unroll 1
  for (int r = 0; r < FIRST_ROWS; ++r) {
    const int j = t.z + (threadIdx.x >> 5) + 8 * r;
    bool sel = false;
    double X = 0, Y = 0;
    if (i < g.mw && j < g.mh) {
      const int pi = g.mx0 + i, pj = g.my0 + j;
      if (pi >= g.ex0 && pi < g.ex0 + g.ew && pj >= g.ey0 && pj < g.ey0 + g.eh) {
        const double* m = stamp + s.stamp_off;
        double err;
        if (s.sampling_mode == APB_SAMPLE_MIDPOINT || s.sampling_mode == APB_SAMPLE_TRAPEZOID) {
          if (g.rw >= 3 && g.rh >= 3) {
            // 3x3 Laplacian, replicate-padded over the working region (_model_methods.py:87-98)
            int ic = min(max(pi, g.rx0 + 1), g.rx0 + g.rw - 2) - g.mx0;
            int jc = min(max(pj, g.ry0 + 1), g.ry0 + g.rh - 2) - g.my0;
            ic = min(max(ic, 1), g.mw - 2);
            jc = min(max(jc, 1), g.mh - 2);
            const double* c = m + (long long)jc * g.mw + ic;
            err = fabs(c[-g.mw] + c[-1] + c[1] + c[g.mw] - 4.0 * c[0]);
          } else {
            err = 0.0;
          }
        } else {
          err = m[(long long)(s.n_act + 1) * s.plane_stride + (long long)j * g.mw + i];
        }
        sel = err > d.thr[mode];
        if (sel) pix_coords(s, d, (double)pi, (double)pj, X, Y);
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, sel);
    if (bal == 0) continue;
    int base = 0;
    if (lane == 0) base = atomicAdd(&q.count[1], __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (sel) {
      const int e = base + __popc(bal & ((1u << lane) - 1));
      if (e < q.cap[1]) {
        const Level& L = q.lv[1];
        L.src[e] = t.x;
        L.x[e] = X;
        L.y[e] = Y;
        L.parent[e] = j * g.mw + i;
        L.child[e] = -1;
      } else {
        *q.overflow = 1;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_select(const DevSrc* __restrict__ src, const DevDyn* __restrict__ dyn,
                                                const int4* __restrict__ tiles, int mode,
                                                const double* __restrict__ stamp, Queues q) {
  select_tile(src, dyn, tiles[blockIdx.x], mode, stamp, q);
}

// ----------------------------------------------------------------------------
// refinement: one thread per queued (sub)pixel, Gauss-Legendre, ballot-compacted re-queue
// (utils/operations.py:123-247)
// ----------------------------------------------------------------------------
template <int KIND, bool GRAD>
__device__ __forceinline__ bool refine_entry(const DevSrc& s, const DevDyn& d, int mode, int depth, double X, double Y,
                                             double* __restrict__ res, int NVp) {
  constexpr int NE = KindInfo<KIND>::NE;
  const int ne = (KIND == APB_SPLINE) ? s.n_elem : NE;
  double acc[(GRAD ? NE : 0) + 1];
  const double scale = s.gsc[depth], ascale = s.gasc[depth];
  double thr = d.thr[mode];
  for (int k = 1; k < depth; ++k) thr *= s.g2d;
  const double centre = gl_integrate<KIND, GRAD>(s, d, X, Y, s.quad_level, scale, ascale, acc);
  if (depth < s.max_depth && fabs(acc[0] - centre) > thr) return true;
  res[0] = acc[0];
  if (GRAD)
    for (int e = 0; e < ne; ++e) {
      const int p = s.plane[e];
      if (p > 0) res[p] = acc[1 + e] * d.chain[e];
    }
  return false;
}

template <bool GRAD>
__global__ void __launch_bounds__(128) k_refine(const DevSrc* __restrict__ src, const DevDyn* __restrict__ dyn,
                                                int mode, int depth, Queues q) {
  apb_math_load();
  const int n = min(q.count[depth], q.cap[depth]);
  const Level& L = q.lv[depth];
  const int lane = threadIdx.x & 31;
  // whole warps iterate together so the ballot below is well defined
  for (int base_t = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base_t < n; base_t += gridDim.x * blockDim.x) {
    const int t = base_t + lane;
    bool split = false;
    int si = 0;
    double X = 0, Y = 0;
    if (t < n) {
      si = L.src[t];
      X = L.x[t];
      Y = L.y[t];
      const DevSrc& s = src[si];
      const DevDyn& d = dyn[si];
      double* res = L.res + (long long)t * q.NVp;
      switch (s.kind) {
        case APB_SERSIC: split = refine_entry<APB_SERSIC, GRAD>(s, d, mode, depth, X, Y, res, q.NVp); break;
        case APB_EXPONENTIAL: split = refine_entry<APB_EXPONENTIAL, GRAD>(s, d, mode, depth, X, Y, res, q.NVp); break;
        case APB_GAUSSIAN: split = refine_entry<APB_GAUSSIAN, GRAD>(s, d, mode, depth, X, Y, res, q.NVp); break;
        case APB_MOFFAT: split = refine_entry<APB_MOFFAT, GRAD>(s, d, mode, depth, X, Y, res, q.NVp); break;
        case APB_SPLINE: split = refine_entry<APB_SPLINE, GRAD>(s, d, mode, depth, X, Y, res, q.NVp); break;
        default: break;
      }
    }
    // re-queue only the pixels that failed the test: gridding^2 children each
    const int nchild = split ? src[si].gridding * src[si].gridding : 0;
    int pre = nchild;  // inclusive warp scan of child counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += v;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    if (total == 0) continue;
    int wbase = 0;
    if (lane == 31) wbase = atomicAdd(&q.count[depth + 1], total);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    if (split) {
      const DevSrc& s = src[si];
      const int first = wbase + pre - nchild;
      if (first + nchild <= q.cap[depth + 1]) {
        L.child[t] = first;
        const Level& C = q.lv[depth + 1];
        const int G = s.gridding;
        const double scale = s.gsc[depth];
        for (int k = 0; k < nchild; ++k) {
          // displacement_grid: linspace(-(G-1)/(2G), (G-1)/(2G), G)  (utils/operations.py:94-102)
          const double dx = s.goff[k % G] * scale;
          const double dy = s.goff[k / G] * scale;
          const int c = first + k;
          C.src[c] = si;
          C.x[c] = X + (s.S[0] * dx + s.S[1] * dy);
          C.y[c] = Y + (s.S[2] * dx + s.S[3] * dy);
          C.parent[c] = t;
          C.child[c] = -1;
        }
      } else {
        *q.overflow = 1;
        L.child[t] = -1;
        double* res = L.res + (long long)t * q.NVp;
        for (int p = 0; p < q.NVp; ++p) res[p] = 0.0;
        // The slots this entry reserved below the capacity are inside the range the next depth processes
        // (min(count, cap)): fill them with inert copies of the parent -- left as they are they hold whatever the
        // allocation contained, and a garbage source index is an illegal access one launch later.  (Their results
        // are never read: the parent has no child list, and the host repeats the pass with larger queues anyway.)
        const Level& C = q.lv[depth + 1];
        for (int c = first; c < q.cap[depth + 1] && c < first + nchild; ++c) {
          C.src[c] = si;
          C.x[c] = X;
          C.y[c] = Y;
          C.parent[c] = t;
          C.child[c] = -1;
        }
      }
    }
  }
}

// children -> parent sums, in child order (operations.py:245), deepest level first
__global__ void k_reduce_level(const DevSrc* __restrict__ src, int depth, Queues q) {
  const int n = min(q.count[depth], q.cap[depth]);
  const Level& L = q.lv[depth];
  const Level& C = q.lv[depth + 1];
  const long long total = (long long)n * q.NVp;
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(w / q.NVp), p = (int)(w % q.NVp);
    const int first = L.child[t];
    if (first < 0) continue;
    const int G = src[L.src[t]].gridding;
    double acc = 0.0;
    for (int k = 0; k < G * G; ++k) acc += C.res[(long long)(first + k) * q.NVp + p];
    L.res[w] = acc;
  }
}

// depth-1 results -> stamp planes
__global__ void k_scatter(const DevSrc* __restrict__ src, Queues q, double* __restrict__ stamp, int grad) {
  const int n = min(q.count[1], q.cap[1]);
  const Level& L = q.lv[1];
  const long long total = (long long)n * q.NVp;
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(w / q.NVp), p = (int)(w % q.NVp);
    const DevSrc& s = src[L.src[t]];
    if (p > (grad ? s.n_act : 0)) continue;
    stamp[s.stamp_off + (long long)p * s.plane_stride + L.parent[t]] = L.res[w];
  }
}

// ----------------------------------------------------------------------------
// Fused adaptive integration (utils/operations.py:123-247) in ONE launch after the ballot-compacted
// depth-1 queue is built: no per-depth launches, no deeper queues, nothing that can overflow.
//  * depth 1: L lanes (power of two >= quad_level^2, <= 32) share one queue entry, one Gauss-Legendre
//    node each; partial sums are combined with an xor-shuffle tree (fixed order => deterministic).
//  * an entry that fails |GL - centre| > thr is subdivided on the spot by the whole warp, depth
//    first: lane c integrates child c (gridding^2 children, 32 per sweep) with its own GL rule and
//    error test; children that fail again are taken one at a time (ballot + ffs), broadcast to the
//    warp and subdivided the same way; child sums are combined with the same shuffle tree.
// Queue entries are dealt round-robin to the warps of the grid so that the expensive entries, which
// cluster at the source centre, spread over all SMs.
// ----------------------------------------------------------------------------
template <int KIND, bool GRAD>
__device__ __forceinline__ void acc_shuffle_sum(const DevSrc& s, Acc<GRAD, KindInfo<KIND>::NE>& acc, unsigned mask, int width) {
  constexpr int NE = KindInfo<KIND>::NE;
  const int ne = (KIND == APB_SPLINE) ? s.n_elem : NE;
  for (int o = width >> 1; o > 0; o >>= 1) {
    acc.v[0] += __shfl_xor_sync(mask, acc.v[0], o);
    if (GRAD)
      for (int e = 0; e < ne; ++e) acc.v[1 + e] += __shfl_xor_sync(mask, acc.v[1 + e], o);
  }
}

// Whole warp: integral of the cell centred (X, Y) at `depth` (already known to need subdivision)
// as the sum of its gridding^2 children at depth+1.  Arguments are warp-uniform.  Result in every lane.
template <int KIND, bool GRAD, int LEVELS_LEFT, typename T = double>
__device__ __forceinline__ void split_cell(const DevSrc& s, const DevDyn& d, int mode, int depth, double X, double Y,
                                           Acc<GRAD, KindInfo<KIND>::NE>& out, int* __restrict__ qcount) {
  constexpr int NE = KindInfo<KIND>::NE;
  const int ne = (KIND == APB_SPLINE) ? s.n_elem : NE;
  const int lane = threadIdx.x & 31;
  const int G = s.gridding, nchild = G * G, cd = depth + 1;
  const double scale = s.gsc[cd], ascale = s.gasc[cd], pscale = s.gsc[depth];
  double thr = d.thr[mode];
  for (int k = 1; k < cd; ++k) thr *= s.g2d;
  if (lane == 0) atomicAdd(&qcount[cd], nchild);
  out.v[0] = 0.0;
  if (GRAD)
    for (int e = 0; e < ne; ++e) out.v[1 + e] = 0.0;
  for (int c0 = 0; c0 < nchild; c0 += 32) {
    const int c = c0 + lane;
    const bool have = c < nchild;
    Acc<GRAD, NE> ca;
    ca.v[0] = 0.0;
    if (GRAD)
      for (int e = 0; e < ne; ++e) ca.v[1 + e] = 0.0;
    double cx = 0.0, cy = 0.0;
    bool again = false;
    if (have) {
      // displacement_grid: linspace(-(G-1)/(2G), (G-1)/(2G), G) of the parent cell (utils/operations.py:94-102)
      const int cyi = (c * s.gmagic) >> 16, cxi = c - cyi * G;
      const double dx = s.goff[cxi] * pscale;
      const double dy = s.goff[cyi] * pscale;
      cx = X + (s.S[0] * dx + s.S[1] * dy);
      cy = Y + (s.S[2] * dx + s.S[3] * dy);
      const double centre = gl_nodes<KIND, GRAD, T>(s, d, cx, cy, scale, ascale, 0, 1, ca);
      again = cd < s.max_depth && fabs(ca.v[0] - centre) > thr;
    }
    if constexpr (LEVELS_LEFT > 0) {
      unsigned bal = __ballot_sync(0xffffffffu, again);
      while (bal) {
        const int b = __ffs(bal) - 1;
        bal &= bal - 1;
        const double bx = __shfl_sync(0xffffffffu, cx, b), by = __shfl_sync(0xffffffffu, cy, b);
        Acc<GRAD, NE> sub;
        split_cell<KIND, GRAD, LEVELS_LEFT - 1, T>(s, d, mode, cd, bx, by, sub, qcount);
        if (lane == b) ca = sub;
      }
    }
    acc_shuffle_sum<KIND, GRAD>(s, ca, 0xffffffffu, 32);
    out.v[0] += ca.v[0];
    if (GRAD)
      for (int e = 0; e < ne; ++e) out.v[1 + e] += ca.v[1 + e];
  }
}

// depth 1 of one queue entry by the L lanes of a group.  Stores the result unless the entry must be
// subdivided; returns that decision (same in every lane of the group).
template <int KIND, bool GRAD, typename T = double>
__device__ __forceinline__ bool depth1_entry(const DevSrc& s, const DevDyn& d, int mode, double X, double Y, int parent,
                                             double* __restrict__ stamp, int L, int gl, unsigned gmask, bool valid) {
  constexpr int NE = KindInfo<KIND>::NE;
  const int ne = (KIND == APB_SPLINE) ? s.n_elem : NE;
  Acc<GRAD, NE> acc;
  double centre = 0.0;
  if (valid) {
    centre = gl_nodes<KIND, GRAD, T>(s, d, X, Y, 1.0, 1.0, gl, L, acc);
  } else {
    acc.v[0] = 0.0;
    if (GRAD)
      for (int e = 0; e < ne; ++e) acc.v[1 + e] = 0.0;
  }
  for (int o = L >> 1; o > 0; o >>= 1) centre += __shfl_xor_sync(gmask, centre, o);
  acc_shuffle_sum<KIND, GRAD>(s, acc, gmask, L);
  if (!valid) return false;
  if (1 < s.max_depth && fabs(acc.v[0] - centre) > d.thr[mode]) return true;
  if (gl == 0) {
    double* base = stamp + s.stamp_off + parent;
    base[0] = acc.v[0];
    if (GRAD)
      for (int e = 0; e < ne; ++e) {
        const int p = s.plane[e];
        if (p > 0) base[(long long)p * s.plane_stride] = acc.v[1 + e] * d.chain[e];
      }
  }
  return false;
}

// whole warp: subdivide one depth-1 entry and store its integral
template <int KIND, bool GRAD, typename T = double>
__device__ __forceinline__ void split_and_store(const DevSrc& s, const DevDyn& d, int mode, double X, double Y, int parent,
                                                double* __restrict__ stamp, int* __restrict__ qcount) {
  constexpr int NE = KindInfo<KIND>::NE;
  const int ne = (KIND == APB_SPLINE) ? s.n_elem : NE;
  Acc<GRAD, NE> out;
  split_cell<KIND, GRAD, APB_MAX_DEPTH - 2, T>(s, d, mode, 1, X, Y, out, qcount);
  if ((threadIdx.x & 31) == 0) {
    double* base = stamp + s.stamp_off + parent;
    base[0] = out.v[0];
    if (GRAD)
      for (int e = 0; e < ne; ++e) {
        const int p = s.plane[e];
        if (p > 0) base[(long long)p * s.plane_stride] = out.v[1 + e] * d.chain[e];
      }
  }
}

// spline sources carry up to 24 elements per accumulator: kept out of line so that their local
// arrays do not inflate the register allocation of the analytic profiles
template <bool GRAD, typename T = double>
__device__ __noinline__ bool depth1_entry_spline(const DevSrc& s, const DevDyn& d, int mode, double X, double Y, int parent,
                                                 double* __restrict__ stamp, int L, int gl, unsigned gmask, bool valid) {
  return depth1_entry<APB_SPLINE, GRAD, T>(s, d, mode, X, Y, parent, stamp, L, gl, gmask, valid);
}
template <bool GRAD, typename T = double>
__device__ __noinline__ void split_and_store_spline(const DevSrc& s, const DevDyn& d, int mode, double X, double Y, int parent,
                                                    double* __restrict__ stamp, int* __restrict__ qcount) {
  split_and_store<APB_SPLINE, GRAD, T>(s, d, mode, X, Y, parent, stamp, qcount);
}

template <bool GRAD, typename T>
__global__ void __launch_bounds__(128, GRAD ? 3 : 4) k_integrate(const DevSrc* __restrict__ src, const DevDyn* __restrict__ dyn, int mode,
                                                      double* __restrict__ stamp, Queues q, int L, int n_max) {
  const int n = min(q.count[1], q.cap[1]);
  if (n >= n_max) return;   // long queues belong to k_integrate_pool
  apb_math_load();
  const Level& Lv = q.lv[1];
  const int lane = threadIdx.x & 31;
  const int epw = 32 / L, g = lane / L, gl = lane - g * L;
  const unsigned gmask = L == 32 ? 0xffffffffu : (((1u << L) - 1u) << (g * L));
  // persistent warps draw tasks (32/L queue entries) from a global counter: a warp that lands on
  // entries needing subdivision simply draws fewer tasks (q.count[0] is zeroed by k_prep)
  for (;;) {
    int task = 0;
    if (lane == 0) task = atomicAdd(&q.count[0], 1);
    task = __shfl_sync(0xffffffffu, task, 0);
    const int base = task * epw;
    if (base >= n) break;
    const int t = base + g;
    const bool valid = t < n;
    int si = 0, parent = 0;
    double X = 0, Y = 0;
    if (valid) {
      si = Lv.src[t];
      X = Lv.x[t];
      Y = Lv.y[t];
      parent = Lv.parent[t];
    }
    bool split = false;
    {
      const DevSrc& s = src[si];
      const DevDyn& d = dyn[si];
      switch (s.kind) {
        case APB_SERSIC: split = depth1_entry<APB_SERSIC, GRAD, T>(s, d, mode, X, Y, parent, stamp, L, gl, gmask, valid); break;
        case APB_EXPONENTIAL: split = depth1_entry<APB_EXPONENTIAL, GRAD, T>(s, d, mode, X, Y, parent, stamp, L, gl, gmask, valid); break;
        case APB_GAUSSIAN: split = depth1_entry<APB_GAUSSIAN, GRAD, T>(s, d, mode, X, Y, parent, stamp, L, gl, gmask, valid); break;
        case APB_MOFFAT: split = depth1_entry<APB_MOFFAT, GRAD, T>(s, d, mode, X, Y, parent, stamp, L, gl, gmask, valid); break;
        case APB_SPLINE: split = depth1_entry_spline<GRAD, T>(s, d, mode, X, Y, parent, stamp, L, gl, gmask, valid); break;
        default: break;
      }
    }
    __syncwarp();
    // entries that failed the error test: the whole warp subdivides them one after the other
    unsigned need = __ballot_sync(0xffffffffu, split && gl == 0);
    while (need) {
      const int b = __ffs(need) - 1;
      need &= need - 1;
      const int sb = __shfl_sync(0xffffffffu, si, b), pb = __shfl_sync(0xffffffffu, parent, b);
      const double Xb = __shfl_sync(0xffffffffu, X, b), Yb = __shfl_sync(0xffffffffu, Y, b);
      const DevSrc& s = src[sb];
      const DevDyn& d = dyn[sb];
      switch (s.kind) {
        case APB_SERSIC: split_and_store<APB_SERSIC, GRAD, T>(s, d, mode, Xb, Yb, pb, stamp, q.count); break;
        case APB_EXPONENTIAL: split_and_store<APB_EXPONENTIAL, GRAD, T>(s, d, mode, Xb, Yb, pb, stamp, q.count); break;
        case APB_GAUSSIAN: split_and_store<APB_GAUSSIAN, GRAD, T>(s, d, mode, Xb, Yb, pb, stamp, q.count); break;
        case APB_MOFFAT: split_and_store<APB_MOFFAT, GRAD, T>(s, d, mode, Xb, Yb, pb, stamp, q.count); break;
        case APB_SPLINE: split_and_store_spline<GRAD, T>(s, d, mode, Xb, Yb, pb, stamp, q.count); break;
        default: break;
      }
    }
  }
}

// ----------------------------------------------------------------------------
// Throughput form of the same integration, for queues far larger than the grid (crowded fields:
// 10^5..10^8 entries).  k_integrate spends 16 lanes on the 9 nodes of a depth-1 entry and 32 lanes on
// the 25 children of a subdivided one; here a CTA draws 128 entries at a time and keeps every lane on
// its own cell:
//   phase 1  one thread per depth-1 entry (all its nodes); entries that pass are stored, the others
//            are listed in shared memory;
//   phase 2  the children of all listed entries are pooled: thread c takes child c % G^2 of listed
//            entry c / G^2 (full lanes whatever the number of failing entries); child integrals go
//            to shared memory, children that fail again are listed;
//   phase 3  each listed child is subdivided by a whole warp, depth first (split_cell, as above);
//   phase 4  one thread per (entry, plane) adds the child integrals in child order and stores.
// Same nodes, same decisions, same child values as k_integrate; only the order in which the
// children of an entry are added differs (sequential instead of a shuffle tree).  k_integrate and
// k_integrate_pool are launched back to back and pick by the queue length, which only the device knows.
// ----------------------------------------------------------------------------
#define POOL_B 128        // depth-1 entries per CTA batch (= CTA size)
#ifndef POOL_MINB
#define POOL_MINB 5       // resident CTAs per SM the value-only kernel is compiled for (register budget)
#endif
#define POOL_CSUM 3200    // child integrals held per CTA (doubles)
#define POOL_C2 800       // grandchild integrals held per CTA (value-only kernel, depth-3 phase)

template <int KIND, bool GRAD, typename T = double>
__device__ __forceinline__ bool pool_child(const DevSrc& s, const DevDyn& d, int mode, double X, double Y, int ch,
                                           double* __restrict__ out, int nv) {
  constexpr int NE = KindInfo<KIND>::NE;
  const int ne = (KIND == APB_SPLINE) ? s.n_elem : NE;
  const int G = s.gridding;
  const double scale = s.gsc[2], ascale = s.gasc[2], thr = d.thr[mode] * s.g2d;
  const int cyi = (ch * s.gmagic) >> 16, cxi = ch - cyi * G;
  const double dx = s.goff[cxi], dy = s.goff[cyi];
  const double cx = X + (s.S[0] * dx + s.S[1] * dy);
  const double cy = Y + (s.S[2] * dx + s.S[3] * dy);
  Acc<GRAD, NE> ca;
  const double centre = gl_nodes<KIND, GRAD, T>(s, d, cx, cy, scale, ascale, 0, 1, ca);
  out[0] = ca.v[0];
  if (GRAD)
    for (int e = 0; e < ne && 1 + e < nv; ++e) out[1 + e] = ca.v[1 + e];
  return 2 < s.max_depth && fabs(ca.v[0] - centre) > thr;
}

template <int KIND, bool GRAD, typename T = double>
__device__ __forceinline__ void pool_split_child(const DevSrc& s, const DevDyn& d, int mode, double X, double Y, int ch,
                                                 double* __restrict__ out, int nv, int* __restrict__ qcount) {
  constexpr int NE = KindInfo<KIND>::NE;
  const int ne = (KIND == APB_SPLINE) ? s.n_elem : NE;
  const int G = s.gridding;
  const int cyi = (ch * s.gmagic) >> 16, cxi = ch - cyi * G;
  const double dx = s.goff[cxi], dy = s.goff[cyi];
  const double cx = X + (s.S[0] * dx + s.S[1] * dy);
  const double cy = Y + (s.S[2] * dx + s.S[3] * dy);
  Acc<GRAD, NE> sub;
  split_cell<KIND, GRAD, APB_MAX_DEPTH - 3, T>(s, d, mode, 2, cx, cy, sub, qcount);
  if ((threadIdx.x & 31) == 0) {
    out[0] = sub.v[0];
    if (GRAD)
      for (int e = 0; e < ne && 1 + e < nv; ++e) out[1 + e] = sub.v[1 + e];
  }
}
// Integral of grandchild `gc` of child `ch` of the depth-1 entry at (X, Y): a depth-3 cell, final when max_depth == 3
// (no further test, utils/operations.py:196-199).  Same centre arithmetic as pool_split_child -> split_cell.
template <int KIND, typename T = double>
__device__ __forceinline__ double pool_grandchild(const DevSrc& s, const DevDyn& d, double X, double Y, int ch, int gc) {
  const int G = s.gridding;
  const int cyi = (ch * s.gmagic) >> 16, cxi = ch - cyi * G;
  const double dx = s.goff[cxi], dy = s.goff[cyi];
  const double cx = X + (s.S[0] * dx + s.S[1] * dy);
  const double cy = Y + (s.S[2] * dx + s.S[3] * dy);
  const int gyi = (gc * s.gmagic) >> 16, gxi = gc - gyi * G;
  const double ex = s.goff[gxi] * s.gsc[2], ey = s.goff[gyi] * s.gsc[2];
  const double gx = cx + (s.S[0] * ex + s.S[1] * ey);
  const double gy = cy + (s.S[2] * ex + s.S[3] * ey);
  Acc<false, KindInfo<KIND>::NE> ca;
  gl_nodes<KIND, false, T>(s, d, gx, gy, s.gsc[3], s.gasc[3], 0, 1, ca);
  return ca.v[0];
}
template <typename T = double>
__device__ __noinline__ double pool_grandchild_spline(const DevSrc& s, const DevDyn& d, double X, double Y, int ch, int gc) {
  return pool_grandchild<APB_SPLINE, T>(s, d, X, Y, ch, gc);
}

template <bool GRAD, typename T = double>
__device__ __noinline__ bool pool_child_spline(const DevSrc& s, const DevDyn& d, int mode, double X, double Y, int ch,
                                               double* __restrict__ out, int nv) {
  return pool_child<APB_SPLINE, GRAD, T>(s, d, mode, X, Y, ch, out, nv);
}
template <bool GRAD, typename T = double>
__device__ __noinline__ void pool_split_child_spline(const DevSrc& s, const DevDyn& d, int mode, double X, double Y, int ch,
                                                     double* __restrict__ out, int nv, int* __restrict__ qcount) {
  pool_split_child<APB_SPLINE, GRAD, T>(s, d, mode, X, Y, ch, out, nv, qcount);
}

// g2 = largest gridding^2 of the plan, nv = values per child (1, or 1 + largest element count)
template <bool GRAD, typename T>
__global__ void __launch_bounds__(POOL_B, GRAD ? 3 : POOL_MINB) k_integrate_pool(const DevSrc* __restrict__ src, const DevDyn* __restrict__ dyn,
                                                                          int mode, double* __restrict__ stamp, Queues q,
                                                                          int n_min, int g2, int nv, int pmax) {
  __shared__ double csum[POOL_CSUM];
  __shared__ double csum2[POOL_C2];             // grandchild integrals of the pooled depth-3 phase (value-only kernel)
  __shared__ double eX[POOL_B], eY[POOL_B];
  __shared__ int eS[POOL_B], eP[POOL_B], list1[POOL_B];
  __shared__ unsigned short list2[POOL_CSUM];   // child index < POOL_CSUM
  __shared__ int s_base, s_nf1, s_nf2;
  const int n = min(q.count[1], q.cap[1]);
  if (n < n_min) return;
  apb_math_load();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per_chunk = max(1, POOL_CSUM / (g2 * nv));   // listed entries whose children fit in csum
  const int step_pi = POOL_B / g2, step_ch = POOL_B - step_pi * g2;
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      s_base = atomicAdd(&q.count[0], POOL_B);
      s_nf1 = 0;
      s_nf2 = 0;
    }
    __syncthreads();
    const int base = s_base;
    if (base >= n) break;
    // ---- phase 1
    {
      const int t = base + tid;
      const bool valid = t < n;
      int si = 0, parent = 0;
      double X = 0, Y = 0;
      if (valid) {
        si = q.lv[1].src[t];
        X = q.lv[1].x[t];
        Y = q.lv[1].y[t];
        parent = q.lv[1].parent[t];
      }
      eS[tid] = si; eP[tid] = parent; eX[tid] = X; eY[tid] = Y;
      const DevSrc& s = src[si];
      const DevDyn& d = dyn[si];
      bool split = false;
      const unsigned me = 1u << lane;
      switch (s.kind) {
        case APB_SERSIC: split = depth1_entry<APB_SERSIC, GRAD, T>(s, d, mode, X, Y, parent, stamp, 1, 0, me, valid); break;
        case APB_EXPONENTIAL: split = depth1_entry<APB_EXPONENTIAL, GRAD, T>(s, d, mode, X, Y, parent, stamp, 1, 0, me, valid); break;
        case APB_GAUSSIAN: split = depth1_entry<APB_GAUSSIAN, GRAD, T>(s, d, mode, X, Y, parent, stamp, 1, 0, me, valid); break;
        case APB_MOFFAT: split = depth1_entry<APB_MOFFAT, GRAD, T>(s, d, mode, X, Y, parent, stamp, 1, 0, me, valid); break;
        case APB_SPLINE: split = depth1_entry_spline<GRAD, T>(s, d, mode, X, Y, parent, stamp, 1, 0, me, valid); break;
        default: break;
      }
      if (split) list1[atomicAdd(&s_nf1, 1)] = tid;
    }
    __syncthreads();
    const int nf1 = s_nf1;
    for (int c0 = 0; c0 < nf1; c0 += per_chunk) {
      const int nb = min(per_chunk, nf1 - c0);
      if (tid == 0) s_nf2 = 0;
      __syncthreads();
      // ---- phase 2
      int nchild_sum = 0;
      int pi = tid / g2, ch = tid - pi * g2;          // (entry, child) of c = tid, advanced by POOL_B per trip
      for (int c = tid; c < nb * g2; c += POOL_B, pi += step_pi, ch += step_ch) {
        if (ch >= g2) { ch -= g2; ++pi; }
        const int e = list1[c0 + pi];
        const DevSrc& s = src[eS[e]];
        if (ch >= s.gridding * s.gridding) continue;
        const DevDyn& d = dyn[eS[e]];
        ++nchild_sum;
        double* out = csum + (long long)c * nv;
        bool again = false;
        switch (s.kind) {
          case APB_SERSIC: again = pool_child<APB_SERSIC, GRAD, T>(s, d, mode, eX[e], eY[e], ch, out, nv); break;
          case APB_EXPONENTIAL: again = pool_child<APB_EXPONENTIAL, GRAD, T>(s, d, mode, eX[e], eY[e], ch, out, nv); break;
          case APB_GAUSSIAN: again = pool_child<APB_GAUSSIAN, GRAD, T>(s, d, mode, eX[e], eY[e], ch, out, nv); break;
          case APB_MOFFAT: again = pool_child<APB_MOFFAT, GRAD, T>(s, d, mode, eX[e], eY[e], ch, out, nv); break;
          case APB_SPLINE: again = pool_child_spline<GRAD, T>(s, d, mode, eX[e], eY[e], ch, out, nv); break;
          default: break;
        }
        if (again) list2[atomicAdd(&s_nf2, 1)] = (unsigned short)c;
      }
      if (nchild_sum) atomicAdd(&q.count[2], nchild_sum);
      __syncthreads();
      // ---- phase 3
      const int nf2 = s_nf2;
      bool pooled3 = false;
      if constexpr (!GRAD) {
        // Value-only passes with max_depth <= 3 (the default): a listed child's grandchildren are final, so they are
        // pooled like the children were -- thread c takes grandchild c % G^2 of listed child c / G^2, every lane on its
        // own cell (the warp-per-child form below keeps 25 of 32 lanes busy and one cell per lane in flight).
        pooled3 = pmax <= 3 && g2 <= POOL_C2;
        if (pooled3) {
          const int per3 = POOL_C2 / g2;             // listed children whose grandchildren fit in csum2
          const int st_k = POOL_B / g2, st_g = POOL_B - st_k * g2;
          for (int k0 = 0; k0 < nf2; k0 += per3) {
            const int nk = min(per3, nf2 - k0);
            int ngc = 0;
            int kk = tid / g2, gc = tid - kk * g2;
            for (int c = tid; c < nk * g2; c += POOL_B, kk += st_k, gc += st_g) {
              if (gc >= g2) { gc -= g2; ++kk; }
              const int cidx = list2[k0 + kk];
              const int pi = cidx / g2, ch = cidx - pi * g2;
              const int e = list1[c0 + pi];
              const DevSrc& s = src[eS[e]];
              double v = 0.0;
              if (gc < s.gridding * s.gridding) {
                const DevDyn& d = dyn[eS[e]];
                ++ngc;
                switch (s.kind) {
                  case APB_SERSIC: v = pool_grandchild<APB_SERSIC, T>(s, d, eX[e], eY[e], ch, gc); break;
                  case APB_EXPONENTIAL: v = pool_grandchild<APB_EXPONENTIAL, T>(s, d, eX[e], eY[e], ch, gc); break;
                  case APB_GAUSSIAN: v = pool_grandchild<APB_GAUSSIAN, T>(s, d, eX[e], eY[e], ch, gc); break;
                  case APB_MOFFAT: v = pool_grandchild<APB_MOFFAT, T>(s, d, eX[e], eY[e], ch, gc); break;
                  case APB_SPLINE: v = pool_grandchild_spline<T>(s, d, eX[e], eY[e], ch, gc); break;
                  default: break;
                }
              }
              csum2[c] = v;
            }
            if (ngc) atomicAdd(&q.count[3], ngc);
            __syncthreads();
            for (int k = tid; k < nk; k += POOL_B) {      // grandchildren added in order: deterministic
              double tot = 0.0;
              for (int g = 0; g < g2; ++g) tot += csum2[k * g2 + g];
              csum[(long long)list2[k0 + k] * nv] = tot;
            }
            __syncthreads();
          }
        }
      }
      for (int k = wid; k < nf2 && !pooled3; k += POOL_B / 32) {
        const int c = list2[k];
        const int pi = c / g2, ch = c - pi * g2;
        const int e = list1[c0 + pi];
        const DevSrc& s = src[eS[e]];
        const DevDyn& d = dyn[eS[e]];
        double* out = csum + (long long)c * nv;
        switch (s.kind) {
          case APB_SERSIC: pool_split_child<APB_SERSIC, GRAD, T>(s, d, mode, eX[e], eY[e], ch, out, nv, q.count); break;
          case APB_EXPONENTIAL: pool_split_child<APB_EXPONENTIAL, GRAD, T>(s, d, mode, eX[e], eY[e], ch, out, nv, q.count); break;
          case APB_GAUSSIAN: pool_split_child<APB_GAUSSIAN, GRAD, T>(s, d, mode, eX[e], eY[e], ch, out, nv, q.count); break;
          case APB_MOFFAT: pool_split_child<APB_MOFFAT, GRAD, T>(s, d, mode, eX[e], eY[e], ch, out, nv, q.count); break;
          case APB_SPLINE: pool_split_child_spline<GRAD, T>(s, d, mode, eX[e], eY[e], ch, out, nv, q.count); break;
          default: break;
        }
      }
      __syncthreads();
      // ---- phase 4
      const int nvals = GRAD ? nv : 1;
      for (int idx = tid; idx < nb * nvals; idx += POOL_B) {
        const int pi = idx / nvals, v = idx - pi * nvals;
        const int e = list1[c0 + pi];
        const DevSrc& s = src[eS[e]];
        const int nchild = s.gridding * s.gridding;
        const double* in = csum + (long long)pi * g2 * nv + v;
        double tot = 0.0;
        for (int ch = 0; ch < nchild; ++ch) tot += in[(long long)ch * nv];
        double* basep = stamp + s.stamp_off + eP[e];
        if (v == 0) {
          basep[0] = tot;
        } else if (v - 1 < s.n_elem) {
          const int p = s.plane[v - 1];
          if (p > 0) basep[(long long)p * s.plane_stride] = tot * dyn[eS[e]].chain[v - 1];
        }
      }
      __syncthreads();
    }
  }
}

// ----------------------------------------------------------------------------
// PSF models: divide by the sum over the evaluation region (psf_model_object.py:255-256),
// quotient rule for the derivative planes.  One CTA per source.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_normalize(const DevSrc* __restrict__ src, const int* __restrict__ list,
                                                   int mode, double* __restrict__ stamp, int grad) {
  __shared__ double sh[8];
  __shared__ double tot_s, dtot_s;
  const DevSrc& s = src[list[blockIdx.x]];
  const Geo& g = s.geo[mode];
  const int n = g.ew * g.eh;
  const long long off0 = (long long)(g.ey0 - g.my0) * g.mw + (g.ex0 - g.mx0);
  double* p0 = stamp + s.stamp_off + off0;
  double v = 0.0;
  for (int q = threadIdx.x; q < n; q += 256) v += p0[(long long)(q / g.ew) * g.mw + (q % g.ew)];
  double t = block_sum<256>(v, sh);
  if (threadIdx.x == 0) tot_s = t;
  __syncthreads();
  const double tot = tot_s;
  const int np = grad ? s.n_act : 0;
  const int pamp = s.amp_elem >= 0 ? s.plane[s.amp_elem] : 0;   // written by k_amp afterwards
  for (int p = 1; p <= np; ++p) {
    if (p == pamp) continue;
    double* pp = p0 + (long long)p * s.plane_stride;
    v = 0.0;
    for (int q = threadIdx.x; q < n; q += 256) v += pp[(long long)(q / g.ew) * g.mw + (q % g.ew)];
    t = block_sum<256>(v, sh);
    if (threadIdx.x == 0) dtot_s = t;
    __syncthreads();
    const double dtot = dtot_s;
    for (int q = threadIdx.x; q < n; q += 256) {
      const long long o = (long long)(q / g.ew) * g.mw + (q % g.ew);
      pp[o] = pp[o] / tot - p0[o] * (dtot / (tot * tot));
    }
    __syncthreads();
  }
  for (int q = threadIdx.x; q < n; q += 256) {
    const long long o = (long long)(q / g.ew) * g.mw + (q % g.ew);
    p0[o] = p0[o] / tot;
  }
}

// ----------------------------------------------------------------------------
// APB_F_AMP: point source drawn from a PSF *model* (point_source.py:122-140).  The sampled (and normalised) PSF-model
// stamp and its derivative planes are scaled by A = 10^flux; the flux plane is ln10 A value (times the chain factor
// of the flux parameter).  One CTA per source, over the evaluation region (= the output window: no PSF border).
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_amp(const DevSrc* __restrict__ src, const DevDyn* __restrict__ dyn,
                                             const int* __restrict__ list, int mode, double* __restrict__ stamp,
                                             int grad) {
  const int si = list[blockIdx.x];
  const DevSrc& s = src[si];
  const DevDyn& d = dyn[si];
  const Geo& g = s.geo[mode];
  const int n = g.ew * g.eh;
  const long long off0 = (long long)(g.ey0 - g.my0) * g.mw + (g.ex0 - g.mx0);
  double* p0 = stamp + s.stamp_off + off0;
  const double A = exp10(d.el[s.amp_elem]);
  const int pamp = s.plane[s.amp_elem];
  const double ca = APB_LN10 * d.chain[s.amp_elem];
  const int np = grad ? s.n_act : 0;
  for (int q = threadIdx.x; q < n; q += 256) {
    const long long o = (long long)(q / g.ew) * g.mw + (q % g.ew);
    const double v = A * p0[o];
    p0[o] = v;
    for (int p = 1; p <= np; ++p) {
      double* pp = p0 + (long long)p * s.plane_stride;
      pp[o] = (p == pamp) ? ca * v : A * pp[o];
    }
  }
}
