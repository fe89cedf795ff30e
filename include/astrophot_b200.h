/*
 * astrophot_b200 — C ABI of the B200-native forward-model-and-fit hot path.
 *
 * The reference (Autostronomy/AstroPhot) has no FFI seam: its hot path is the
 * Python method surface `model.sample()` / `model.jacobian()` / `fit.LM.step()`.
 * This header is the seam a maintainer binds instead (ctypes stub in
 * INTEGRATION.md).  Every entry point names the reference code it replaces.
 *
 * Conventions
 *  - plain C structs, pointers and sizes; no torch types.
 *  - every `double*` / `uint8_t*` that is documented "device" is a CUDA device
 *    pointer owned by the caller (torch allocates it); the library never frees
 *    caller memory.  Internal workspace is owned by the plan.
 *  - every function returns 0 on success, <0 on error; apb_last_error() gives
 *    the message of the last failure on the calling thread.
 *  - `stream` is a CUstream / cudaStream_t passed as void* (NULL = default).
 *    Calls are asynchronous with respect to the host unless stated.
 *  - a plan is thread-compatible (one thread at a time), like the reference's
 *    model objects (core_model.py:484-502 mutates model.parameters).
 *  - image data are row-major [y][x] (astrophot: data[j, i]); pixel (i, j) has
 *    plane coordinates  S . ((i, j) - rij) + rxy   (image/wcs.py:561-584).
 */
#ifndef ASTROPHOT_B200_H
#define ASTROPHOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APB_MAX_ELEM 24 /* elements (scalar parameters) per source            */
#define APB_MAX_PROF 20 /* spline nodes                                       */
#define APB_MAX_DEPTH 4 /* integrate_max_depth                                */
#define APB_MAX_QUAD 9  /* Gauss-Legendre order per axis                      */

/* source kinds; element order = column order of the reference's Jacobian
 *  */
enum {
  APB_SERSIC = 0,      /* cx cy q PA n Re Ie      models/sersic_model.py:41       */
  APB_EXPONENTIAL = 1, /* cx cy q PA Re Ie        models/exponential_model.py:40  */
  APB_GAUSSIAN = 2,    /* cx cy q PA sigma flux   models/gaussian_model.py:37     */
  APB_MOFFAT = 3,      /* cx cy q PA n Rd I0      models/moffat_model.py:23       */
  APB_SPLINE = 4,      /* cx cy q PA v[0..K-1]    models/spline_model.py:27       */
  APB_POINT = 5,       /* cx cy flux              models/point_source.py:17       */
  APB_FLAT_SKY = 6,    /* cx cy F                 models/flatsky_model.py:13      */
  APB_PLANE_SKY = 7    /* cx cy F dx dy           models/planesky_model.py:13 (set APB_F_RADIAL) */
};
enum { APB_F_RADIAL = 1, APB_F_NORMALIZE = 2, /* psf_model_object.py:36-57,255 */
       APB_F_AMP = 4 };  /* the LAST element is a log10 amplitude applied after sampling / normalisation and never seen
                            by the profile: a point source drawn from a PSF *model* (point_source.py:122-140) is that
                            model's profile centred on the point source, times 10^flux.  No PSF on such a source. */
enum { APB_TR_NONE = 0, APB_TR_LOWER, APB_TR_UPPER, APB_TR_BOTH, APB_TR_CYCLIC }; /* utils/conversions/optimization.py:6-54 */
enum { APB_SAMPLE_MIDPOINT = 0, APB_SAMPLE_SIMPSONS, APB_SAMPLE_QUAD, APB_SAMPLE_TRAPEZOID }; /* _model_methods.py:82-148 */
enum { APB_INTEGRATE_NONE = 0, APB_INTEGRATE_THRESHOLD };                                  /* _model_methods.py:155-184 */
enum { APB_REF_MEAN = 0, APB_REF_SERSIC_FLUX }; /* _model_methods.py:151-152, sersic_model.py:87-89 */
enum { APB_SHIFT_NONE = 0, APB_SHIFT_BILINEAR = 1, APB_SHIFT_LANCZOS = 10 }; /* _model_methods.py:187-230; "lanczos:k" = 10 + k, k = 1..8 */
/* "fft" in the reference = APB_CONV_AUTO here (tiled direct convolution for small stamps, FFT for
 * large ones: same valid region, utils/operations.py:9-36); "direct" = APB_CONV_DIRECT */
enum { APB_CONV_AUTO = 0, APB_CONV_DIRECT = 1, APB_CONV_FFT = 2 };

typedef struct apb_plan apb_plan_t;

/* one free parameter of the flat vector x (param/parameter.py:330-420) */
typedef struct {
  int32_t transform; /* APB_TR_*  */
  int32_t _pad;
  double lo, hi;     /* limits (ignored where absent) */
} apb_param_t;

/* one target image region: target[fit_window] (fit/lm.py:191-222) */
typedef struct {
  int32_t H, W;
  double S[4];           /* pixelscale, row-major 2x2                       */
  double rij[2], rxy[2]; /* reference pixel / plane position                */
  const double *data;    /* device, H*W; may be NULL when only sampling     */
  const double *weight;  /* device, H*W; NULL = ones (target_image.py:182)  */
  const uint8_t *mask;   /* device, H*W; 1 = ignore pixel; NULL = none      */
  int32_t flags;         /* APB_IMG_AUX: grid of an auxiliary PSF model     */
  int32_t _pad;
} apb_image_t;
/* APB_IMG_AUX: the image is the PSF_Image grid an auxiliary PSF model is sampled on
 * (model_object.py:133-147,307-310): its sources are evaluated on every pass, but it is no output of
 * apb_sample / apb_jacobian (its model_out / jac_out entry is ignored) and no term of chi^2. */
enum { APB_IMG_AUX = 1 };

/* a PSF stamp (image/psf_image.py:17-93): odd h, w; un-normalised is fine */
typedef struct {
  int32_t h, w;
  const double *data; /* device, h*w; NULL when the stamp comes from `source`                       */
  int32_t source;     /* >= 0: index of the PSF-model source (APB_F_NORMALIZE, alone on an
                         APB_IMG_AUX image of h x w pixels) that produces the stamp on every pass; its
                         free parameters become columns of the Jacobian of every source using this
                         PSF (model_object.py:133-147: set_aux_psf links the parameters).  -1: data */
  int32_t _pad;
} apb_psf_t;

/* one component model, lowered (models/model_object.py:64-95 for the knobs) */
typedef struct {
  int32_t kind, flags, image;
  int32_t out[4]; /* x0 y0 w h : where the source adds flux (image pixels)            */
  int32_t fwd[4]; /* working window when sampled in the forward model
                     (group_model_object.py:211-227 hands sub-models the group window) */
  int32_t jac[4]; /* working window when differentiated (_model_methods.py:294-299)   */
  int32_t n_elem;
  int32_t slot[APB_MAX_ELEM]; /* index into x, or -1 = locked                         */
  double cval[APB_MAX_ELEM];  /* natural value of locked elements                     */
  int32_t n_prof;
  double prof[APB_MAX_PROF];  /* spline node radii                                    */
  int32_t sampling_mode, quad_init, integrate_mode, quad_level, gridding, max_depth;
  int32_t ref_mode, psf, psf_shift;
  int32_t conv_mode; /* APB_CONV_*: psf_convolve_mode (_model_methods.py:245-255) */
  int32_t owner;     /* index into apb_opts_t.owners of the model this source is a piece of (an image
                        cut into tiles hands every tile its clipped copy of the model); ignored when no
                        owner table is given */
  int32_t upscale;   /* super-sampled PSF (models/model_object.py:312-315,348-349; point_source.py:147-149,181):
                        the PSF's pixels are 1 / upscale of the image's; the source is sampled, integrated and
                        convolved (a point source: its PSF dropped) on that finer grid and block-summed back.  `out`,
                        `fwd`, `jac` stay in image pixels.  Needs a PSF image (apb_psf_t.data); 0 or 1: none */
  double tolerance, softening;
  /* the model's own mask (models/model_object.py:370-371, point_source.py:184-185): device, mask_rect[2] x mask_rect[3]
   * bytes, row-major, element (0, 0) = image pixel (mask_rect[0], mask_rect[1]); where it is non-zero the source
   * contributes nothing -- value and derivatives (the reference differentiates through the product).  NULL = none. */
  const uint8_t *mask;
  int32_t mask_rect[4];
} apb_source_t;

/* One component model of the WHOLE fit, for fits whose pixels are split into tiles and / or over
 * ranks: the pieces of a model (apb_source_t.owner) share its parameter slots, and
 * the block-sparse form of J^T W J (apb_lm_solve_sparse) is laid out on the owners -- the same
 * layout on every rank, so that the ranks' blocks add up with one sum all-reduce.  `image` and
 * `out` are in the coordinates of the uncut images and only serve to find which owners overlap. */
typedef struct {
  int32_t image;
  int32_t out[4];             /* x0 y0 w h of the model's window on the uncut image       */
  int32_t n_slot;
  int32_t slot[APB_MAX_ELEM]; /* its free parameters (indices into x), in element order   */
} apb_owner_t;

typedef struct {
  int64_t queue_capacity; /* entries per refinement level; 0 = automatic (grown by apb_plan_reserve) */
  int32_t flags;          /* bits 0-1: APB_CONV_* override for every source; bit 2: per-depth
                             refinement launches instead of the fused k_integrate; bit 3: pooled
                             (throughput) integration kernel whatever the queue length; bit 4: the profile
                             kernels (first pass, mean reference, sub-pixel integration) compute in fp32 --
                             the reference's AP_config.ap_dtype = torch.float32 (AP_config.py:7); pixel
                             offsets from the source centre are still formed in fp64, planes stay fp64 */
  int32_t n_owners;       /* 0: every source is its own owner */
  const apb_owner_t *owners;
} apb_opts_t;

/* counters of the last call, for benchmarks (SURVEY.md §8d "SPE") */
typedef struct {
  int64_t first_pass_evals;           /* profile evaluations of the first pass       */
  int64_t queued[APB_MAX_DEPTH + 1];  /* entries processed per refinement depth 1..  */
  int64_t launches;                   /* kernels launched by the last call           */
  int64_t overflow;                   /* !=0: a queue overflowed, results invalid    */
  /* totals since the plan was created; [0]: value-only sampling passes, [1]: value + derivative passes.
     profile evaluations of a pass = first-pass evaluations + quad_level^2 x the entries of every depth */
  int64_t cum_passes[2];
  int64_t cum_first_pass_evals[2];
  int64_t cum_queued[2][APB_MAX_DEPTH + 1];
} apb_stats_t;

/* Build the device tables and workspace for a lowered model tree.
 * Replaces the per-call Python walk of Group_Model.sample / .jacobian
 * (group_model_object.py:183-283).  Synchronous. */
int apb_plan_create(const apb_source_t *src, int n_src, const apb_image_t *img, int n_img,
                    const apb_psf_t *psf, int n_psf, const apb_param_t *par, int n_par,
                    const apb_opts_t *opts, apb_plan_t **out);
int apb_plan_destroy(apb_plan_t *plan);

/* Point image `image` of the plan at other data / weight / mask buffers of the same shape (device pointers, same
 * meaning as in apb_image_t; weight / mask may be NULL).  Takes effect for work enqueued on `stream` after this call:
 * a fit loop can upload the next exposure into a second set of buffers while the current one is being fitted, and
 * one plan serves every image that shares the model geometry (the reference rebuilds Y, W, mask per LM object,
 * fit/lm.py:191-222). */
int apb_plan_set_image_data(apb_plan_t *plan, int image, const double *data, const double *weight,
                            const uint8_t *mask, void *stream);

/* seam 1 — model(parameters=x, as_representation=as_rep) -> model image(s)
 * (core_model.py:484-502, model_object.py:258-375, group_model_object.py:183-231).
 * x: device, n_par doubles.  model_out[i]: device, H_i*W_i doubles, OVERWRITTEN. */
int apb_sample(apb_plan_t *plan, const double *x, int as_rep, double *const *model_out, void *stream);

/* seam 2 — model.jacobian(parameters=x, as_representation=as_rep)
 * (_model_methods.py:260-347, group_model_object.py:233-283).
 * jac_out[i]: device, H_i*W_i*n_par doubles (pixel-major, parameter fastest), OVERWRITTEN.
 * For small problems and tests only; LM never materialises this. */
int apb_jacobian(apb_plan_t *plan, const double *x, int as_rep, double *const *jac_out, void *stream);

/* seam 3 — the inside of LM.step (fit/lm.py:256-260):  Y0 = forward(x); J = jacobian(x);
 * JtWJ = J^T W J (n_par x n_par, row-major), JtWr = J^T W (Y - Y0), chi2 = sum W (Y - Y0)^2,
 * all over unmasked pixels.  The per-source stamp Jacobian stays cached in the plan for
 * apb_geodesic.  All outputs device pointers. */
int apb_normal_eq(apb_plan_t *plan, const double *x_rep, int as_rep, double *JtWJ, double *JtWr,
                  double *chi2 /* 2 doubles: chi^2, status flag as in apb_chi2 */, void *stream);

/* fit/lm.py:277-281,401-406:  rpp = J^T [ (2/d) ( (W (Y(x + d h) - Y) - r)/d - W (J h) ) ]
 * with J, r from the last apb_normal_eq.  xdh = x + d*h (device, n_par), h device. */
int apb_geodesic(apb_plan_t *plan, const double *xdh_rep, const double *h, double d, double *rpp,
                 void *stream);

/* fit/lm.py:289-293,373-378: out[0] = sum W (Y - model(x))^2 over unmasked pixels,
 * out[1] = 1.0 if every model pixel is finite, 0.0 if not, -1.0 if a sub-pixel refinement queue
 * overflowed in this or an earlier call (sticky; results invalid: call apb_plan_reserve and repeat).
 * (Caller divides by ndf.) */
int apb_chi2(apb_plan_t *plan, const double *x_rep, double *out2, void *stream);

/* The adaptive integration (utils/operations.py:150-247) queues re-gridded sub-pixels per depth;
 * depth 1 is bounded by the pixel count, deeper levels are sized heuristically.  After an
 * overflow, grow the queues: caps[d], d = 1..APB_MAX_DEPTH, entries wanted at depth d (0 = keep),
 * or NULL to size from the counts of the last call.  Clears the overflow flag.  Synchronous. */
int apb_plan_reserve(apb_plan_t *plan, const int64_t *caps);

/* fit/lm.py:359-371:  solve (H o (I + (1-I)/(1+L)) + L I (1 + diag H)) h = g.
 * H: device P*P (not modified), g, h: device P.  info: device int, 0 ok. */
int apb_lm_solve(const double *H, const double *g, double L, int P, double *h, int *info, void *stream);

/* The same damped system, dense, beyond the single-CTA solver (fit/lm.py:359-371 with a few hundred to a few thousand
 * parameters some of which are shared between sources: joint multi-band fits, auxiliary PSF models).  The matrix is
 * symmetric positive definite for L > 0: apb_chol_factor builds it from H and factors it (blocked Cholesky, one persistent
 * cooperative kernel), apb_chol_solve serves any number of right-hand sides -- the two solves of a lambda-trial
 * (lm.py:274,283) share one factor.  W: device, P*P + 2 doubles.  info: device int, 0 ok, 1 = a pivot was not positive
 * (non-finite H): solve another way.  rhs and x may alias. */
int apb_chol_factor(const double *H, double L, int P, double *W, int *info, void *stream);
int apb_chol_solve(const double *W, const double *rhs, int P, double *x, void *stream);

/* The same damped system for large parameter counts (crowded fields, fit/lm.py:359-371 with P ~ 1e4):
 * J^T W J of the last apb_normal_eq is kept inside the plan as its list of <= 8x8 source-pair blocks, and the
 * system is solved by block-Jacobi preconditioned conjugate gradients in one persistent cooperative kernel.
 * g, h: device, n_par.  x0: device, n_par, starting point of the iteration, or NULL for zero (the solution of the previous
 * lambda-trial of the same LM iteration saves iterations; x0 must not alias h).  info: device, 2 doubles
 * {iterations, final |r|/|b|}.  tol <= 0: 1e-14; max_iter <= 0: 2000.
 * Returns 1 (no error set) if the plan cannot use it (a parameter shared between sources, as in joint fits):
 * use apb_lm_solve or a dense library solver on the JtWJ of apb_normal_eq instead. */
int apb_lm_solve_sparse(apb_plan_t *plan, const double *g, double L, const double *x0, double *h, double *info,
                        double tol, int max_iter, void *stream);

/* The block-sparse J^T W J of the plan as one flat array: the blocks, tightly packed (n_a x n_b doubles each,
 * row-major), followed by diag(J^T W J) (n_par doubles).  apb_plan_block_doubles returns its length (0: the plan has no
 * block-sparse form, see apb_lm_solve_sparse).  apb_plan_bind_blocks makes apb_normal_eq write it into a
 * caller-owned device buffer of that length instead of the plan's own, which is how a fit sharded by
 * image tile merges the ranks' normal equations: sum all-reduce of that buffer and of JtWr, then every
 * rank solves the same system (fit/lm.py:256-260 on the pixels of all ranks).  A plan with a block-sparse
 * form accepts JtWJ == NULL in apb_normal_eq (no dense copy). */
long long apb_plan_block_doubles(apb_plan_t *plan);
int apb_plan_bind_blocks(apb_plan_t *plan, double *buf);

/* fit/lm.py:268-293, one pass of the lambda search with every tensor operation on the device:
 *   h = solve(L, g);  rpp = geodesic(x + d h, h, d);  a = -solve(L, rpp)/2 (zeros when L <= 1e-4);
 *   ha = h + acceleration a;  rec = { chi2(x + ha), status flag (see apb_chi2), |a|, |h| }.
 * H, g from the last apb_normal_eq; h_out, ha_out: device, n_par doubles; rec: device, 4 doubles,
 * the one record the host reads back per trial.  n_par <= 159 (single-CTA solver).
 * plan2: NULL, or a second plan created from the same scene: when acceleration == 0 (the reference's
 * default, lm.py:187) chi2(x + h) does not depend on the geodesic term and is evaluated on plan2,
 * concurrently with the geodesic pass on plan. */
int apb_lm_trial(apb_plan_t *plan, apb_plan_t *plan2, const double *H, const double *g, double L,
                 const double *x_rep, double d, double acceleration, double *h_out, double *ha_out, double *rec,
                 void *stream);

/* The same trial, run SPECULATIVELY beside another one: `donor` is the plan whose last apb_normal_eq left the stamp
 * Jacobian and the residual at x (read only), `plan` / `plan2` are forward-only plans of the same scene with their own
 * workspace.  The lambda search of fit/lm.py:268-357 is sequential, but its next damping is one of two values
 * (L / Ldn after an improvement, L * Lup after a failure): evaluating the likely one on a second stream while the
 * current trial runs halves the latency of the search without changing any result. */
int apb_lm_trial_spec(apb_plan_t *plan, apb_plan_t *plan2, apb_plan_t *donor, const double *H, const double *g, double L,
                      const double *x_rep, double d, double acceleration, double *h_out, double *ha_out, double *rec,
                      void *stream);

/* The same trial in two halves for fits sharded over several GPUs (acceleration == 0 only):
 *   begin: h = solve(L, g); buf = { local rpp[n_par], local chi2(x + h), #non-finite, #overflow }
 *   -- the caller sums buf over the ranks (n_par + 3 doubles, one all-reduce per trial) --
 *   end:   a = -solve(L, rpp)/2;  rec = { chi2, status flag, |a|, |h| };  ha = h. */
int apb_lm_trial_begin(apb_plan_t *plan, apb_plan_t *plan2, const double *H, const double *g, double L,
                       const double *x_rep, double d, double *h_out, double *buf, void *stream);
int apb_lm_trial_end(apb_plan_t *plan, const double *H, double L, const double *x_rep, const double *h,
                     const double *buf, double *ha_out, double *rec, void *stream);

int apb_plan_stats(apb_plan_t *plan, apb_stats_t *out); /* synchronises the plan's last stream */

/* ---- multi-GPU exchange (SURVEY.md 8b `apb_allreduce`, 8e) ------------------------------------------------------
 * A fit whose pixels are sharded over the GPUs of one node sums J^T W J | J^T W r | chi^2 (fit/lm.py:256-260) and the
 * per-trial record over the ranks.  These buffers are small (<= a few MB), so the exchange is ONE kernel over NVLink
 * peer memory, ordered on the caller's stream behind the kernel that produced the buffer -- no host round trip, no NCCL
 * launch: every rank publishes its buffer in a slot of cudaIpc-shared device memory, flags every peer, and adds all
 * slots in rank order (deterministic, bit-identical on all ranks).
 *   apb_comm_alloc   allocates this rank's exchange buffer (slots of max_doubles) and returns its 64-byte IPC handle;
 *   (the caller gathers the handles of all ranks, e.g. with torch.distributed.all_gather)
 *   apb_comm_create  maps the peers' buffers; `handles`: world x 64 bytes in rank order; takes ownership of `local`;
 *   apb_allreduce    buf (device, n <= max_doubles doubles) <- sum over ranks, in place, asynchronous on `stream`.
 *                    Every rank must make the same sequence of calls. */
typedef struct apb_comm apb_comm_t;
int apb_comm_alloc(size_t max_doubles, void **local_out, void *handle_out);
int apb_comm_create(int rank, int world, void *local, const void *handles, size_t max_doubles, apb_comm_t **out);
int apb_allreduce(apb_comm_t *comm, double *buf, size_t n, void *stream);
int apb_comm_destroy(apb_comm_t *comm);

/* ---- measurement (bench.py; no reference counterpart) ---------------------------------- */
typedef struct {
  char name[32];
  int64_t launches;
  double total_ms; /* sum of CUDA-event durations on the launching stream */
} apb_kernel_time_t;
int apb_profile(apb_plan_t *plan, int enable); /* bracket every kernel launch with CUDA events */
int apb_profile_read(apb_plan_t *plan, apb_kernel_time_t *out, int max_out, int *n_out, int reset);
long long apb_launch_count(void);               /* kernels launched by this library since load */
int apb_bench_peaks(double *dfma_tflops, double *copy_gbs); /* DFMA-stream and fp64 copy ceilings, synchronous */
int apb_fft_length(int n); /* transform length the FFT convolution uses for a padded stamp of n pixels */
const char *apb_last_error(void);
int apb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ASTROPHOT_B200_H */
