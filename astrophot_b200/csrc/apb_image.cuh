// PSF handling (sub-pixel shifted stamps, tiled direct convolution, point sources),
// image assembly / residuals, and the fused normal-equation accumulation.
#pragma once
#include "apb_internal.cuh"

// ----------------------------------------------------------------------------
// shifted, normalised PSF stamp per source + derivative wrt the centre
// (_model_methods.py:187-243, utils/interpolate.py:282-329).  One CTA per source.
// Output planes at psfst + s.psf_off: K0, Kcx, Kcy, each sph x spw; Kcx/Kcy already contain
// d shift / d centre = S^-1 and the chain factor, so conv(value, Kcx) is the J column.
// ----------------------------------------------------------------------------
// Lanczos taps (utils/interpolate.py:145-163): f(t) = sinc(t) sinc(t / k), sinc(u) = sin(pi u) / (pi u)
__device__ __forceinline__ double apb_sinc(double u) { return u == 0.0 ? 1.0 : sinpi(u) / (APB_PI * u); }
__device__ __forceinline__ double apb_dsinc(double u) {
  if (fabs(u) < 1e-2) {   // (cos(pi u) - sinc(u)) / u cancels catastrophically near 0: series
    const double p2 = APB_PI * APB_PI, u2 = u * u;
    return u * (-p2 / 3.0 + u2 * (p2 * p2 / 30.0 - u2 * (p2 * p2 * p2 / 840.0)));
  }
  return (cospi(u) - apb_sinc(u)) / u;
}
// value of the zero-padded raw stamp (or of a derivative plane of an auxiliary PSF model) cross-correlated with the
// separable Lanczos kernel at padded-frame position (yp, xp) (_model_methods.py:209-227: conv2d, padding "same")
__device__ __forceinline__ double lanczos_corr(const double* __restrict__ p, int stride, int ph, int pw, int lz, int yp, int xp,
                                               const double* __restrict__ ky, const double* __restrict__ kx) {
  double acc = 0.0;
  for (int j = -lz; j <= lz; ++j) {
    const int y = yp + j - lz;
    if (y < 0 || y >= ph) continue;
    double row = 0.0;
    for (int i = -lz; i <= lz; ++i) {
      const int x = xp + i - lz;
      if (x >= 0 && x < pw) row = fma(kx[i + lz], p[y * stride + x], row);
    }
    acc = fma(ky[j + lz], row, acc);
  }
  return acc;
}

__global__ void __launch_bounds__(256) k_psf_stamp(const DevSrc* __restrict__ src, const DevDyn* __restrict__ dyn,
                                                   const int* __restrict__ list, const apb_psf_t* __restrict__ psfs,
                                                   double* __restrict__ psfst, int grad, int mode,
                                                   const double* __restrict__ stamp, const double* __restrict__ outar) {
  __shared__ double sh[8];
  __shared__ double tot_s[3];
  const int si = list[blockIdx.x];
  const DevSrc& s = src[si];
  const DevDyn& d = dyn[si];
  const apb_psf_t P = psfs[s.psf];
  const int pw = P.w, ph = P.h;
  const int spw = s.spw, sph = s.sph, n = spw * sph;
  double* K0 = psfst + s.psf_off;
  double* K1 = K0 + n;
  double* K2 = K1 + n;
  const bool shifted = s.psf_shift != APB_SHIFT_NONE;
  // the raw stamp: a static PSF image, or plane 0 of the auxiliary PSF model sampled earlier in this pass
  const double* pdat = P.data;
  int pstride = pw;
  if (s.psf_src >= 0) {
    const PlaneView pv = out_plane(src[s.psf_src], mode, 0, stamp, outar);
    pdat = pv.p;
    pstride = pv.stride;
  }
  // padded image is (ph+2) x (pw+2) [lanczos:k: (ph+2k) x (pw+2k)]; stamps either keep the pad (galaxies) or crop it (points)
  const int lz = s.psf_shift >= 10 ? s.psf_shift - 10 : 0;      // Lanczos order, 0: bilinear / none
  const int crop = (s.kind == APB_POINT && shifted) ? (lz ? lz : 1) : 0;
  const int W2 = pw + 2, H2 = ph + 2;
  __shared__ double lzk[4][17];     // Lx, Ly (normalised), dLx/dsx, dLy/dsy (quotient rule applied)
  if (lz) {
    const int nt = 2 * lz + 1;
    if (threadIdx.x < 2 * nt) {
      const int axis = threadIdx.x / nt, i = threadIdx.x - axis * nt;
      const double t = (double)(i - lz) + (axis ? d.sy : d.sx);
      lzk[axis][i] = apb_sinc(t) * apb_sinc(t / lz);
      lzk[2 + axis][i] = apb_dsinc(t) * apb_sinc(t / lz) + apb_sinc(t) * apb_dsinc(t / lz) / lz;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      const int axis = threadIdx.x;
      double S = 0.0, dS = 0.0;
      for (int i = 0; i < nt; ++i) { S += lzk[axis][i]; dS += lzk[2 + axis][i]; }
      for (int i = 0; i < nt; ++i) {
        const double L = lzk[axis][i];
        lzk[2 + axis][i] = lzk[2 + axis][i] / S - L * (dS / (S * S));
        lzk[axis][i] = L / S;
      }
    }
    __syncthreads();
  }
  double v0 = 0, v1 = 0, v2 = 0;
  for (int q = threadIdx.x; q < n; q += 256) {
    const int a = q / spw, b = q % spw;
    double val, gx = 0, gy = 0;
    if (!shifted) {
      val = pdat[a * pstride + b];
    } else if (lz) {
      val = lanczos_corr(pdat, pstride, ph, pw, lz, a + crop, b + crop, lzk[1], lzk[0]);
      if (grad) {
        gx = lanczos_corr(pdat, pstride, ph, pw, lz, a + crop, b + crop, lzk[1], lzk[2]);
        gy = lanczos_corr(pdat, pstride, ph, pw, lz, a + crop, b + crop, lzk[3], lzk[0]);
      }
    } else {
      const int i = b + crop, j = a + crop;  // index in the padded image
      const double x = (double)i - d.sx, y = (double)j - d.sy;
      int x0 = (int)floor(x), y0 = (int)floor(y);
      int x1 = min(max(x0 + 1, 1), W2 - 1), y1 = min(max(y0 + 1, 1), H2 - 1);
      x0 = min(max(x0, 0), W2 - 2);
      y0 = min(max(y0, 0), H2 - 2);
      auto pad = [&](int yy, int xx) -> double {
        return (yy >= 1 && yy <= ph && xx >= 1 && xx <= pw) ? pdat[(yy - 1) * pstride + (xx - 1)] : 0.0;
      };
      const double fa = pad(y0, x0), fb = pad(y1, x0), fc = pad(y0, x1), fd = pad(y1, x1);
      const double wx0 = (double)x1 - x, wx1 = x - (double)x0, wy0 = (double)y1 - y, wy1 = y - (double)y0;
      val = fa * (wx0 * wy0) + fb * (wx0 * wy1) + fc * (wx1 * wy0) + fd * (wx1 * wy1);
      gx = fa * wy0 + fb * wy1 - fc * wy0 - fd * wy1;  // d/dsx
      gy = fa * wx0 - fb * wx0 + fc * wx1 - fd * wx1;  // d/dsy
    }
    K0[q] = val;
    v0 += val;
    if (grad && shifted) {
      K1[q] = gx;
      K2[q] = gy;
      v1 += gx;
      v2 += gy;
    }
  }
  double t = block_sum<256>(v0, sh);
  if (threadIdx.x == 0) tot_s[0] = t;
  t = block_sum<256>(v1, sh);
  if (threadIdx.x == 0) tot_s[1] = t;
  t = block_sum<256>(v2, sh);
  if (threadIdx.x == 0) tot_s[2] = t;
  __syncthreads();
  const double tot = tot_s[0], tx = tot_s[1], ty = tot_s[2];
  // auxiliary PSF model: d/d theta of the shifted, normalised stamp.  The shift is linear in the PSF, so with
  // D = shift(d psf / d theta):  dK = D / T - shift(psf) sum(D) / T^2  (planes 3.. of this source's stamps; the PSF
  // source's planes already carry d value / d representation)
  if (grad && s.psf_src >= 0) {
    const DevSrc& psrc = src[s.psf_src];
    for (int k = 0; k < s.n_pp; ++k) {
      const PlaneView dv = out_plane(psrc, mode, 1 + k, stamp, outar);
      double* Kk = K0 + (long long)(3 + k) * n;
      double vk = 0.0;
      for (int q = threadIdx.x; q < n; q += 256) {
        const int a = q / spw, b = q % spw;
        double val;
        if (!shifted) {
          val = dv.p[a * dv.stride + b];
        } else if (lz) {
          val = lanczos_corr(dv.p, dv.stride, ph, pw, lz, a + crop, b + crop, lzk[1], lzk[0]);
        } else {
          const int i = b + crop, j = a + crop;
          const double x = (double)i - d.sx, y = (double)j - d.sy;
          int x0 = (int)floor(x), y0 = (int)floor(y);
          int x1 = min(max(x0 + 1, 1), W2 - 1), y1 = min(max(y0 + 1, 1), H2 - 1);
          x0 = min(max(x0, 0), W2 - 2);
          y0 = min(max(y0, 0), H2 - 2);
          auto pad = [&](int yy, int xx) -> double {
            return (yy >= 1 && yy <= ph && xx >= 1 && xx <= pw) ? dv.p[(yy - 1) * dv.stride + (xx - 1)] : 0.0;
          };
          const double wx0 = (double)x1 - x, wx1 = x - (double)x0, wy0 = (double)y1 - y, wy1 = y - (double)y0;
          val = pad(y0, x0) * (wx0 * wy0) + pad(y1, x0) * (wx0 * wy1) + pad(y0, x1) * (wx1 * wy0) + pad(y1, x1) * (wx1 * wy1);
        }
        Kk[q] = val;
        vk += val;
      }
      const double tk = block_sum<256>(vk, sh);
      __shared__ double tk_s;
      if (threadIdx.x == 0) tk_s = tk;
      __syncthreads();
      const double tks = tk_s;
      for (int q = threadIdx.x; q < n; q += 256) Kk[q] = Kk[q] / tot - K0[q] * (tks / (tot * tot));
      __syncthreads();
    }
  }
  for (int q = threadIdx.x; q < n; q += 256) {
    const double st = K0[q];
    if (grad && shifted) {
      const double gx = K1[q] / tot - st * (tx / (tot * tot));
      const double gy = K2[q] / tot - st * (ty / (tot * tot));
      K1[q] = (s.Sinv[0] * gx + s.Sinv[2] * gy) * d.chain[0];
      K2[q] = (s.Sinv[1] * gx + s.Sinv[3] * gy) * d.chain[1];
    }
    K0[q] = st / tot;
  }
}

// ----------------------------------------------------------------------------
// point sources (point_source.py:145-175): PSF stamp x 10^flux placed at the rounded centre
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_point(const DevSrc* __restrict__ src, const DevDyn* __restrict__ dyn,
                                               const int* __restrict__ list, const double* __restrict__ psfst,
                                               double* __restrict__ outar, int grad) {
  const int si = list[blockIdx.x];
  const DevSrc& s = src[si];
  const DevDyn& d = dyn[si];
  // (the fine output window: the output window itself unless the PSF is super-sampled, DevSrc::up)
  const int n = s.fow * s.foh;
  const int spw = s.spw, sph = s.sph, ns = spw * sph;
  const double* K0 = psfst + s.psf_off;
  const double F = d.k[0];
  double* o = outar + s.fine_off;
  const int x_lo = d.rx - (spw - 1) / 2, y_lo = d.ry - (sph - 1) / 2;
  for (int q = threadIdx.x; q < n; q += 256) {
    const int x = s.fox + q % s.fow, y = s.foy + q / s.fow;
    const int b = x - x_lo, a = y - y_lo;
    const bool in = a >= 0 && a < sph && b >= 0 && b < spw;
    const int k = a * spw + b;
    o[q] = in ? F * K0[k] : 0.0;
    if (grad) {
      if (s.plane[0] > 0) o[(long long)s.plane[0] * n + q] = (in && s.psf_shift != APB_SHIFT_NONE) ? F * K0[ns + k] : 0.0;
      if (s.plane[1] > 0) o[(long long)s.plane[1] * n + q] = (in && s.psf_shift != APB_SHIFT_NONE) ? F * K0[2 * ns + k] : 0.0;
      if (s.plane[2] > 0) o[(long long)s.plane[2] * n + q] = in ? APB_LN10 * F * K0[k] * d.chain[2] : 0.0;
    }
  }
}

// ----------------------------------------------------------------------------
// tiled direct convolution, fp64.  One CTA = 32 rows x 64 columns of output; warp w owns an
// 8-column strip, lane = row; the input tile and the kernel are staged in shared memory; each
// thread slides an 8-wide register window along the kernel row (1 LDS per 8 DFMA).
// job: {src, in_plane, kernel index (0..2), out_plane};  tile: {job, tx0, ty0}
// ----------------------------------------------------------------------------
#define CONV_TW 64
#define CONV_TH 32
__global__ void __launch_bounds__(256) k_conv(const DevSrc* __restrict__ src, const int4* __restrict__ jobs,
                                              const int4* __restrict__ tiles, int mode,
                                              const double* __restrict__ stamp, const double* __restrict__ psfst,
                                              double* __restrict__ outar) {
  extern __shared__ double smem[];
  const int4 tl = tiles[blockIdx.x];
  const int4 jb = jobs[tl.x];
  const DevSrc& s = src[jb.x];
  const Geo& g = s.geo[mode];
  const int spw = s.spw, sph = s.sph;
  const int cw = (spw - 1) / 2, chh = (sph - 1) / 2;
  const int iw = CONV_TW + spw - 1, ih = CONV_TH + sph - 1;
  const int istr = iw | 1;  // odd row stride: conflict-free when lanes walk down rows
  double* sK = smem;                 // sph*spw, stored flipped so the inner loop walks forward
  double* sI = smem + sph * spw;     // ih * istr
  const double* K = psfst + s.psf_off + (long long)jb.z * spw * sph;
  for (int q = threadIdx.x; q < spw * sph; q += 256) sK[q] = K[spw * sph - 1 - q];
  // input tile origin in evaluation-region coordinates
  const int ex = tl.y + s.bx - cw, ey = tl.z + s.by - chh;
  const double* in = stamp + s.stamp_off + (long long)jb.y * s.plane_stride +
                     (long long)(g.ey0 - g.my0) * g.mw + (g.ex0 - g.mx0);
  // lanczos:k stamps are 2 (k - 1) pixels wider than the PSF border: the reference convolves circularly over the padded
  // image (utils/operations.py:9-36, img_prepadded), so their outer taps wrap around it
  const bool wrap = s.psf_shift >= 10;
  for (int q = threadIdx.x; q < iw * ih; q += 256) {
    const int r = q / iw, c = q % iw;
    int yy = ey + r, xx = ex + c;
    if (wrap) {
      yy = ((yy % g.eh) + g.eh) % g.eh;
      xx = ((xx % g.ew) + g.ew) % g.ew;
    }
    sI[r * istr + c] = (yy >= 0 && yy < g.eh && xx >= 0 && xx < g.ew) ? in[(long long)yy * g.mw + xx] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.0;
  // out[y][x] = sum_{a,b} Kflip[a][b] * tile[y + a][x + b]
  for (int a = 0; a < sph; ++a) {
    const double* row = sI + (lane + a) * istr + w * 8;
    const double* kr = sK + a * spw;
    double win[8];
#pragma unroll
    for (int k = 0; k < 7; ++k) win[k + 1] = row[k];
    for (int b = 0; b < spw; ++b) {
#pragma unroll
      for (int k = 0; k < 7; ++k) win[k] = win[k + 1];
      win[7] = row[b + 7];
      const double kv = kr[b];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fma(kv, win[k], acc[k]);
    }
  }
  const int oy = tl.z + lane;
  if (oy < s.foh) {
    double* o = outar + s.fine_off + (long long)jb.w * s.fow * s.foh + (long long)oy * s.fow;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ox = tl.y + w * 8 + k;
      if (ox < s.fow) o[ox] = acc[k];
    }
  }
}

// ----------------------------------------------------------------------------
// super-sampled PSFs (model_object.py:348-349, point_source.py:181: ``working_image.reduce(psf_upscale)``): the fine
// output window of a source (up x up pixels per image pixel) summed into its planes in image pixels.
// grid: (sources of the list, planes, row slices)
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reduce_up(const DevSrc* __restrict__ src, const int* __restrict__ list, int mode,
                                                   const double* __restrict__ stamp, double* __restrict__ outar, int grad) {
  const DevSrc& s = src[list[blockIdx.x]];
  const int pl = blockIdx.y;
  if (pl > (grad ? s.n_act : 0)) return;
  const int up = s.up, ow = s.ow;
  // the fine output window: planes of the out arena (convolved sources, point sources), or -- a source without
  // convolution, i.e. a point source drawn from a PSF model -- the output window inside its stamp
  const double* f;
  int fow;
  if (s.fine_off >= 0) {
    f = outar + s.fine_off + (long long)pl * s.fow * s.foh;
    fow = s.fow;
  } else {
    const Geo& g = s.geo[mode];
    f = stamp + s.stamp_off + (long long)pl * s.plane_stride + (long long)(s.foy - g.my0) * g.mw + (s.fox - g.mx0);
    fow = g.mw;
  }
  double* o = outar + s.out_off + (long long)pl * s.ow * s.oh;
  const int n = s.ow * s.oh;
  for (int q = blockIdx.z * 256 + threadIdx.x; q < n; q += 256 * gridDim.z) {
    const int y = q / ow, x = q - y * ow;
    const double* b = f + (long long)(y * up) * fow + x * up;
    double acc = 0.0;
    for (int i = 0; i < up; ++i)
      for (int j = 0; j < up; ++j) acc += b[(long long)i * fow + j];
    o[q] = acc;
  }
}

// ----------------------------------------------------------------------------
// assemble the model image from the per-source stamps (gather by 32x32 image tile, sources in
// model order => deterministic), fused with residual / chi^2 / finiteness.
// tile: {image, tx0, ty0, bin};  bins CSR: bin_ptr, bin_src
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_assemble(const DevSrc* __restrict__ src, const DevDyn* __restrict__ dyn,
                                                  const apb_image_t* __restrict__ imgs, const int4* __restrict__ tiles,
                                                  const int* __restrict__ bin_ptr, const int* __restrict__ bin_src,
                                                  int mode, const double* __restrict__ stamp,
                                                  const double* __restrict__ outar, double* const* __restrict__ model_out,
                                                  double* const* __restrict__ resid_out, double* __restrict__ chipart,
                                                  unsigned int* __restrict__ done, double* __restrict__ out2, int write_flag,
                                                  const int* __restrict__ overflow) {
  __shared__ double sh[8];
  __shared__ bool last_s;
  const int4 t = tiles[blockIdx.x];
  const apb_image_t im = imgs[t.x];
  const int b0 = bin_ptr[t.w], b1 = bin_ptr[t.w + 1];
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int x = t.y + lx;
  // the four rows of this thread side by side: their loads are independent and all in flight together
  double m[4] = {0.0, 0.0, 0.0, 0.0};
  for (int b = b0; b < b1; ++b) {
    const int si = bin_src[b];
    const DevSrc& s = src[si];
    if (x < s.ox || x >= s.ox + s.ow) continue;
    if (s.kind == APB_FLAT_SKY) {
      const double v = dyn[si].k[0];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int y = t.z + ly + 8 * r;
        if (y >= s.oy && y < s.oy + s.oh && !src_masked(s, x, y)) m[r] += v;
      }
    } else {
      const PlaneView v = out_plane(s, mode, 0, stamp, outar);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int y = t.z + ly + 8 * r;
        if (y >= s.oy && y < s.oy + s.oh && !src_masked(s, x, y)) m[r] += v.p[(long long)(y - s.oy) * v.stride + (x - s.ox)];
      }
    }
  }
  double chi = 0.0, bad = 0.0;
  double dat[4], wgt[4];
  bool keep[4], in[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int y = t.z + ly + 8 * r;
    in[r] = x < im.W && y < im.H;
    const long long p = (long long)y * im.W + x;
    keep[r] = in[r] && !(im.mask && im.mask[p]);
    dat[r] = (in[r] && im.data) ? im.data[p] : 0.0;
    wgt[r] = (in[r] && im.weight) ? im.weight[p] : 1.0;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (!in[r]) continue;
    const int y = t.z + ly + 8 * r;
    const long long p = (long long)y * im.W + x;
    if (model_out) model_out[t.x][p] = m[r];
    if (im.data) {
      const double df = m[r] - dat[r];
      const double rr = keep[r] ? wgt[r] * df : 0.0;   // r = W (Y0 - Y) on unmasked pixels (lm.py:381-385)
      if (resid_out) resid_out[t.x][p] = rr;
      if (keep[r]) {
        chi += wgt[r] * df * df;
        if (!isfinite(m[r])) bad = 1.0;
      }
    }
  }
  if (!chipart) return;
  const double c = block_sum<256>(chi, sh);
  const double bsum = block_sum<256>(bad, sh);
  // chi^2 record: out2[0] = chi^2; out2[1] = 1 finite, 0 non-finite pixels, -1 a refinement queue
  // overflowed (results invalid: apb_plan_reserve, then repeat the call).  The CTA that finishes
  // last adds the per-tile partials in tile order: one launch, fixed summation order.
  if (threadIdx.x == 0) {
    chipart[2 * blockIdx.x] = c;
    chipart[2 * blockIdx.x + 1] = bsum;
    __threadfence();
    last_s = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last_s) return;
  __threadfence();
  double cs = 0.0, bs = 0.0;
  for (int q = threadIdx.x; q < (int)gridDim.x; q += 256) {
    cs += __ldcg(&chipart[2 * q]);
    bs += __ldcg(&chipart[2 * q + 1]);
  }
  cs = block_sum<256>(cs, sh);
  bs = block_sum<256>(bs, sh);
  if (threadIdx.x == 0) {
    out2[0] = cs;
    if (write_flag) out2[1] = *overflow ? -1.0 : ((bs == 0.0 && isfinite(cs)) ? 1.0 : 0.0);
    *done = 0u;
  }
}

// dense Jacobian for small problems (seam 2): J[pix * P + slot] += plane
__global__ void __launch_bounds__(256) k_jac_dense(const DevSrc* __restrict__ src, const apb_image_t* __restrict__ imgs,
                                                   const int4* __restrict__ tiles, const int* __restrict__ bin_ptr,
                                                   const int* __restrict__ bin_src, const double* __restrict__ stamp,
                                                   const double* __restrict__ outar, const double* __restrict__ skyJ,
                                                   double* const* __restrict__ jac_out, int P) {
  const int4 t = tiles[blockIdx.x];
  const apb_image_t im = imgs[t.x];
  const int b0 = bin_ptr[t.w], b1 = bin_ptr[t.w + 1];
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  for (int r = 0; r < 4; ++r) {
    const int x = t.y + lx, y = t.z + ly + 8 * r;
    if (x >= im.W || y >= im.H) continue;
    double* J = jac_out[t.x] + ((long long)y * im.W + x) * P;
    for (int b = b0; b < b1; ++b) {
      const int si = bin_src[b];
      const DevSrc& s = src[si];
      if (x < s.ox || x >= s.ox + s.ow || y < s.oy || y >= s.oy + s.oh || src_masked(s, x, y)) continue;
      for (int e = 0; e < s.n_elem_all; ++e) {
        const int p = s.plane[e];
        if (p <= 0) continue;
        double v;
        if (s.kind == APB_FLAT_SKY) {
          v = skyJ[si];
        } else {
          const PlaneView pv = out_plane(s, 1, p, stamp, outar);
          v = pv.p[(long long)(y - s.oy) * pv.stride + (x - s.ox)];
        }
        J[s.slot[e]] += v;
      }
    }
  }
}

// ----------------------------------------------------------------------------
// normal equations.  work item: source a (planes pa0.., na), source b (planes pb0.., nb), a rectangle
// of image pixels (rows y0..y0+nrow of the overlap).  Each CTA accumulates the na x nb block of
// J^T W J (and, for diagonal items, J^T r) over its rectangle in registers, reduces in a fixed
// order and writes one partial; k_block_final sums the partials of a block in order and adds
// them into H / g.  The Jacobian exists only as the per-source stamp planes.
// ----------------------------------------------------------------------------
struct BlockItem {
  int a, b;            // sources
  int pa0, na, pb0, nb;  // 0-based offsets into the active-plane lists
  int x0, y0, w, h;    // rectangle in image pixels
  int diag;            // 1: a == b and pa0 == pb0 (symmetric; also accumulates J^T r)
  int block;           // block id
};

__device__ __forceinline__ double plane_at(const DevSrc& s, int plane, int x, int y, const double* stamp,
                                           const double* outar, const double* skyJ, int si) {
  if (s.kind == APB_FLAT_SKY) return skyJ[si];
  const PlaneView pv = out_plane(s, 1, plane, stamp, outar);
  return pv.p[(long long)(y - s.oy) * pv.stride + (x - s.ox)];
}

#define NB_MAX 8
#define BLK_VALS (NB_MAX * NB_MAX + NB_MAX)

// D(8x8) += A(8x4) . B(4x8) on the FP64 tensor pipe.  Fragments (PTX ISA, mma.m8n8k4.f64):
// lane l holds A[l>>2][l&3], B[l&3][l>>2] and D[l>>2][2*(l&3) + {0,1}].
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// The contraction over pixels is a tall-skinny GEMM (8 x Npix) . (Npix x 8): four pixels per DMMA.
// Lane l reads ONE Jacobian value per 4-pixel group -- plane (l>>2) at pixel (l&3) -- which is at the
// same time its A fragment (times the weight) and, for diagonal blocks, its B fragment.  Two
// accumulator registers per lane instead of 64: occupancy is no longer register-bound and the
// kernel streams the derivative planes at HBM speed.  J^T r rides along on the FP64 pipe.
__global__ void __launch_bounds__(256) k_blocks(const DevSrc* __restrict__ src, const apb_image_t* __restrict__ imgs,
                                                const BlockItem* __restrict__ items, const double* __restrict__ stamp,
                                                const double* __restrict__ outar, const double* __restrict__ skyJ,
                                                double* const* __restrict__ resid, double gsign, int vec_only,
                                                double* __restrict__ part) {
  __shared__ double sh[8][BLK_VALS];
  const BlockItem it = items[blockIdx.x];
  const DevSrc& A = src[it.a];
  const DevSrc& B = src[it.b];
  const apb_image_t im = imgs[A.image];
  const double* rimg = resid[A.image];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int pl = lane >> 2, pq = lane & 3;     // plane and pixel-in-group of this lane
  const bool has_a = pl < it.na, has_b = pl < it.nb;
  const bool need_r = it.diag || vec_only;
  // per-lane plane pointers (sky planes are a constant)
  const bool a_sky = A.kind == APB_FLAT_SKY, b_sky = B.kind == APB_FLAT_SKY;
  PlaneView va{nullptr, 0}, vb{nullptr, 0};
  if (has_a && !a_sky) va = out_plane(A, 1, it.pa0 + pl + 1, stamp, outar);
  if (has_b && !b_sky) vb = out_plane(B, 1, it.pb0 + pl + 1, stamp, outar);
  const double ca = a_sky ? skyJ[it.a] : 0.0, cb = b_sky ? skyJ[it.b] : 0.0;
  double d0 = 0.0, d1 = 0.0, gv = 0.0;
  const int n = it.w * it.h;
  const int ngroups = (n + 3) >> 2;
#pragma unroll 4
  for (int gq = wid; gq < ngroups; gq += 8) {
    const int q = 4 * gq + pq;
    double ja = 0.0, jb = 0.0, w = 0.0, r = 0.0;
    if (q < n) {
      const int yy = q / it.w, xx = q - yy * it.w;
      const int x = it.x0 + xx, y = it.y0 + yy;
      const long long p = (long long)y * im.W + x;
      const bool masked = im.mask && im.mask[p];
      if (!masked) {
        w = im.weight ? im.weight[p] : 1.0;
        if (need_r) r = rimg[p];
        if (has_a && !src_masked(A, x, y)) ja = a_sky ? ca : va.p[(long long)(y - A.oy) * va.stride + (x - A.ox)];
        if (it.diag) jb = ja;
        else if (has_b && !vec_only && !src_masked(B, x, y)) jb = b_sky ? cb : vb.p[(long long)(y - B.oy) * vb.stride + (x - B.ox)];
      }
    }
    gv = fma(r, ja, gv);
    if (!vec_only) dmma_8x8x4(d0, d1, w * ja, jb);
  }
  // J^T r: the four lanes of a plane hold partial sums
  gv += __shfl_xor_sync(0xffffffffu, gv, 1);
  gv += __shfl_xor_sync(0xffffffffu, gv, 2);
  sh[wid][pl * NB_MAX + 2 * pq] = d0;
  sh[wid][pl * NB_MAX + 2 * pq + 1] = d1;
  if (pq == 0) sh[wid][NB_MAX * NB_MAX + pl] = gv;
  __syncthreads();
  if (threadIdx.x < BLK_VALS) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += sh[k][threadIdx.x];
    part[(long long)blockIdx.x * BLK_VALS + threadIdx.x] = v;
  }
}

// block descriptor: its items are contiguous [item0, item0+nitem)
struct BlockDesc {
  int a, b, pa0, na, pb0, nb, diag, item0, nitem;
  // block of the owner-level sparse matrix this one adds into: offset in the packed block array (-1: none), its leading
  // dimension (parameters of its b side), 1: transposed
  long long coff;
  int cld, ctrans;
};

// grid (blocks, BLK_VALS/8): one warp per value of a block; its lanes stride over the block's partials
// (independent loads, 4 in flight per lane) and combine with a fixed shuffle tree, then lane 0 stores the value's
// total in btot[block][value].  Nothing is added into H or g here: an entry of the normal equations may collect
// several blocks (linked parameters of joint fits, the pieces of a model cut into tiles), and adding them with atomics
// made joint fits differ in the last bits from run to run.  k_block_gather adds them in a fixed order.
__global__ void __launch_bounds__(256) k_block_final(const BlockDesc* __restrict__ blocks, const double* __restrict__ part,
                                                     int vec_only, double* __restrict__ btot) {
  const int bi = blockIdx.x;
  const BlockDesc bd = blocks[bi];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int v = blockIdx.y * 8 + wid;
  const int i = v < NB_MAX * NB_MAX ? v / NB_MAX : v - NB_MAX * NB_MAX;
  const int j = v < NB_MAX * NB_MAX ? v % NB_MAX : -1;
  if (i >= bd.na) return;
  if (j >= 0 && (j >= bd.nb || vec_only)) return;
  if (j < 0 && !(bd.diag || vec_only)) return;
  const double* pp = part + (long long)bd.item0 * BLK_VALS + v;
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  int k = lane;
  for (; k + 96 < bd.nitem; k += 128) {
    t0 += pp[(long long)k * BLK_VALS];
    t1 += pp[(long long)(k + 32) * BLK_VALS];
    t2 += pp[(long long)(k + 64) * BLK_VALS];
    t3 += pp[(long long)(k + 96) * BLK_VALS];
  }
  for (; k < bd.nitem; k += 32) t0 += pp[(long long)k * BLK_VALS];
  double tot = (t0 + t1) + (t2 + t3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xffffffffu, tot, o);
  if (lane == 0) btot[(long long)bi * BLK_VALS + v] = tot;
}

// One thread per entry of the normal equations that any block contributes to: the contributions (indices into btot,
// listed by the host in block order) are added in that order and the sum is STORED -- every entry has exactly one
// writer, so J^T W J, J^T W r, the block-sparse copy for the PCG solver and its diagonal are bit-reproducible run to
// run whatever the sharing of parameters between models.  Targets: 0 vector (g, sign gsign), 1 block-sparse values,
// 2 diagonal of H, 3 dense H (listed last: the launch stops before them when the caller wants no dense matrix).
struct GatherDst {
  long long index;
  int kind, s0, s1, _pad;
};
__global__ void __launch_bounds__(256) k_block_gather(const GatherDst* __restrict__ dst, int ndst, const int* __restrict__ srcs,
                                                      const double* __restrict__ btot, double* __restrict__ g, double gsign,
                                                      double* __restrict__ bvals, double* __restrict__ diagH,
                                                      double* __restrict__ H) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= ndst) return;
  const GatherDst d = dst[t];
  double* out = d.kind == 0 ? g : d.kind == 1 ? bvals : d.kind == 2 ? diagH : H;
  if (!out) return;
  double v = 0.0;
  for (int k = d.s0; k < d.s1; ++k) v += btot[srcs[k]];
  out[d.index] = d.kind == 0 ? gsign * v : v;
}

// J h per pixel and v = (2/d) ((rh - r)/d - W J h)   (lm.py:401-406); gather by image tile
__global__ void __launch_bounds__(256) k_geo_v(const DevSrc* __restrict__ src, const apb_image_t* __restrict__ imgs,
                                               const int4* __restrict__ tiles, const int* __restrict__ bin_ptr,
                                               const int* __restrict__ bin_src, const double* __restrict__ stamp,
                                               const double* __restrict__ outar, const double* __restrict__ skyJ,
                                               const double* __restrict__ h, double dstep,
                                               double* const* __restrict__ r0, double* const* __restrict__ rh_inout) {
  const int4 t = tiles[blockIdx.x];
  const apb_image_t im = imgs[t.x];
  const int b0 = bin_ptr[t.w], b1 = bin_ptr[t.w + 1];
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  // a thread owns four pixels of one column (rows ly, ly+8, ly+16, ly+24 of the tile): per derivative plane their four
  // loads are independent and in flight together -- the kernel is a stream over the planes, bound by loads in flight
  const int x = t.y + lx;
  double jh[4] = {0.0, 0.0, 0.0, 0.0};
  if (x < im.W) {
    for (int b = b0; b < b1; ++b) {
      const int si = bin_src[b];
      const DevSrc& s = src[si];
      if (x < s.ox || x >= s.ox + s.ow) continue;
      bool in[4];
      long long off[4];
      bool any = false;
      const bool sky = s.kind == APB_FLAT_SKY;
      const PlaneView pv = sky ? PlaneView{nullptr, 0} : out_plane(s, 1, 0, stamp, outar);
      const long long pstride = s.out_off >= 0 ? (long long)s.ow * s.oh : s.plane_stride;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int y = t.z + ly + 8 * r;
        in[r] = y < im.H && y >= s.oy && y < s.oy + s.oh && !src_masked(s, x, y);
        off[r] = in[r] ? (long long)(y - s.oy) * pv.stride + (x - s.ox) : 0;
        any |= in[r];
      }
      if (!any) continue;
      for (int e = 0; e < s.n_elem_all; ++e) {
        const int pl = s.plane[e];
        if (pl <= 0) continue;
        const double hv = h[s.slot[e]];
        if (sky) {
          const double c = skyJ[si] * hv;
#pragma unroll
          for (int r = 0; r < 4; ++r) jh[r] += in[r] ? c : 0.0;
        } else {
          const double* base = pv.p + (long long)pl * pstride;
          double v[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) v[r] = in[r] ? base[off[r]] : 0.0;
#pragma unroll
          for (int r = 0; r < 4; ++r) jh[r] = fma(v[r], hv, jh[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int y = t.z + ly + 8 * r;
      if (y >= im.H) continue;
      const long long p = (long long)y * im.W + x;
      const bool keep = !(im.mask && im.mask[p]);
      const double w = im.weight ? im.weight[p] : 1.0;
      const double v = keep ? (2.0 / dstep) * ((rh_inout[t.x][p] - r0[t.x][p]) / dstep - w * jh[r]) : 0.0;
      rh_inout[t.x][p] = v;
    }
  }
}

// ----------------------------------------------------------------------------
// damped solve for small P: one CTA, Gaussian elimination with partial pivoting in shared memory
// (lm.py:359-371), with the vector algebra of one lambda-trial (lm.py:274-290) as epilogues so that
// a trial is kernels only, no host-side tensor arithmetic:
//   mode 1: h = solve(H, g);            xout = x + d h                      (point of the geodesic probe)
//   mode 2: s = solve(H, rhs = rpp);    a = L > 1e-4 ? -s/2 : 0;  ha = hin + acc a;  xout = x + ha;
//           rec[2] = |a|, rec[3] = |hin|                                     (curvature test rho = |a|/|h|)
// ----------------------------------------------------------------------------
struct LmEpi {
  int mode;
  const double* x;
  const double* hin;
  double d, acc;
  double* xout;
  double* ha;     // mode 1: x + h (or NULL);  mode 2: ha
  double* rec;
  const double* tail;  // mode 2: {chi^2, non-finite count, overflow count} to fold into rec, or NULL
};

// joins the concurrent chi^2 pass of a trial into the record the host reads
// (ovf_c: the donor plan of a trial that reads another plan's stamp Jacobian -- a queue overflow during THAT plan's
//  Jacobian pass left J^T W J and J^T W r without some sub-pixel refinements; it must reach the record the host reads,
//  or the grow-and-retry of LM.step never fires.  c2 may alias rec.)
__global__ void k_trial_join(const double* c2, const int* __restrict__ ovf_a, const int* __restrict__ ovf_b,
                             const int* __restrict__ ovf_c, double* rec) {
  const double chi = c2[0], flag = c2[1];
  rec[0] = chi;
  rec[1] = (*ovf_a || (ovf_b && *ovf_b) || (ovf_c && *ovf_c)) ? -1.0 : flag;
}
// same, as summable counts for a cross-rank all-reduce: tail = {chi^2, ranks with non-finite pixels,
// ranks whose refinement queues overflowed}
__global__ void k_trial_tail(const double* __restrict__ c2, const int* __restrict__ ovf_a, const int* __restrict__ ovf_b,
                             double* __restrict__ tail) {
  const bool ovf = *ovf_a || (ovf_b && *ovf_b) || c2[1] < 0.0;
  tail[0] = c2[0];
  tail[1] = (!ovf && c2[1] == 0.0) ? 1.0 : 0.0;
  tail[2] = ovf ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(256) k_lm_solve_small(const double* __restrict__ H, const double* __restrict__ g,
                                                        double L, int P, double* __restrict__ h, int* __restrict__ info,
                                                        LmEpi epi) {
  extern __shared__ double sm[];
  double* A = sm;          // P x (P+1) augmented
  __shared__ int piv_s;
  const int ld = P + 1;
  for (int q = threadIdx.x; q < P * P; q += blockDim.x) {
    const int i = q / P, j = q % P;
    const double hij = H[q];
    A[i * ld + j] = (i == j) ? hij + L * (1.0 + hij) : hij / (1.0 + L);
  }
  for (int q = threadIdx.x; q < P; q += blockDim.x) A[q * ld + P] = g[q];
  __syncthreads();
  int bad = 0;
  for (int k = 0; k < P; ++k) {
    if (threadIdx.x == 0) {
      int pv = k;
      double best = fabs(A[k * ld + k]);
      for (int i = k + 1; i < P; ++i) {
        const double v = fabs(A[i * ld + k]);
        if (v > best) { best = v; pv = i; }
      }
      piv_s = pv;
      if (!(best > 0.0)) bad = 1;
    }
    __syncthreads();
    const int pv = piv_s;
    if (pv != k)
      for (int j = threadIdx.x; j <= P; j += blockDim.x) {
        const double tmp = A[k * ld + j];
        A[k * ld + j] = A[pv * ld + j];
        A[pv * ld + j] = tmp;
      }
    __syncthreads();
    const double inv = 1.0 / A[k * ld + k];
    for (int q = threadIdx.x; q < (P - k - 1) * (P - k); q += blockDim.x) {
      const int i = k + 1 + q / (P - k), j = k + 1 + q % (P - k);
      A[i * ld + j] -= A[i * ld + k] * inv * A[k * ld + j];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int i = P - 1; i >= 0; --i) {
      double v = A[i * ld + P];
      for (int j = i + 1; j < P; ++j) v -= A[i * ld + j] * h[j];
      h[i] = v / A[i * ld + i];
    }
    if (info) *info = bad;
    if (epi.mode == 1) {
      for (int i = 0; i < P; ++i) {
        epi.xout[i] = epi.x[i] + epi.d * h[i];
        if (epi.ha) epi.ha[i] = epi.x[i] + h[i];
      }
    } else if (epi.mode == 2) {
      double na = 0.0, nh = 0.0;
      for (int i = 0; i < P; ++i) {
        const double a = L > 1e-4 ? -h[i] / 2 : 0.0;
        const double hh = epi.hin[i];
        const double ha = hh + a * epi.acc;
        na += a * a;
        nh += hh * hh;
        epi.ha[i] = ha;
        epi.xout[i] = epi.x[i] + ha;
      }
      epi.rec[2] = sqrt(na);
      epi.rec[3] = sqrt(nh);
      if (epi.tail) {   // (summed) tail of apb_lm_trial_begin -> chi^2 and status flag
        epi.rec[0] = epi.tail[0];
        epi.rec[1] = epi.tail[2] > 0.0 ? -1.0 : (epi.tail[1] > 0.0 ? 0.0 : 1.0);
      }
    }
  }
}
