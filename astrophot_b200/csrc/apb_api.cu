// libastrophot_b200: host side of the C ABI (include/astrophot_b200.h).
// Builds the device tables once per plan, then every call is a fixed sequence of
// kernel launches on the caller's stream with no host synchronisation.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <unordered_map>

#include "apb_internal.cuh"
#include "apb_sample.cuh"
#include "apb_image.cuh"
#include "apb_fft.cuh"
#include "apb_solve.cuh"
#include "apb_comm.cuh"
#include "apb_chol.cuh"


static thread_local std::string g_err;
#define APB_FAIL(msg)                      \
  do {                                     \
    g_err = std::string(msg);              \
    return -1;                             \
  } while (0)
#define CU(call)                                                                               \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      g_err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +     \
              std::to_string(__LINE__) + ")";                                                  \
      return -2;                                                                               \
    }                                                                                          \
  } while (0)

template <typename T>
static int upload(const std::vector<T>& v, T** out) {
  *out = nullptr;
  const size_t n = std::max<size_t>(v.size(), 1);
  CU(cudaMalloc((void**)out, n * sizeof(T)));
  if (!v.empty()) CU(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// ---- optional per-kernel timing (CUDA events on the launching stream) ------------------------
enum KId { K_PREP, K_PSF, K_FIRST, K_MEAN, K_SELECT, K_REFINE, K_REDUCE, K_SCATTER, K_NORM, K_POINT, K_CONV,
           K_ASSEMBLE, K_CHI, K_JAC, K_BLOCKS, K_BLOCKFIN, K_GEOV, K_FFTROWS, K_FFTCOLS, K_FFTINV, K_INTEGRATE, K_INTEGRATE_G, K_FIRST_G, K_POOL, K_POOL_G, K_PCG, K_AMP, K_UPSUM, K_COUNT };
static const char* const kKNames[K_COUNT] = {"k_prep", "k_psf_stamp", "k_first", "k_mean", "k_select", "k_refine",
                                             "k_reduce_level", "k_scatter", "k_normalize", "k_point", "k_conv",
                                             "k_assemble", "k_chi_final", "k_jac_dense", "k_blocks", "k_block_final",
                                             "k_geo_v", "k_fft_rows", "k_fft_cols", "k_fft_rows_inv", "k_integrate", "k_integrate_grad", "k_first_grad", "k_integrate_pool", "k_integrate_pool_grad", "k_pcg", "k_amp", "k_reduce_up"};
static long long g_launches = 0;
struct ProfRec { int id; cudaEvent_t a, b; };

struct ModeTables {
  int4* tiles = nullptr;       // first-pass tiles {src, tx, ty, 0}
  int n_tiles = 0;
  int4* chunks = nullptr;      // mean-reference chunks {src, first, n, slot}
  int n_chunks = 0;
  int* mean_list = nullptr;    // sources with REF_MEAN + threshold
  int n_mean = 0;
  int4* conv_jobs[2] = {nullptr, nullptr};   // [grad]
  int4* conv_tiles[2] = {nullptr, nullptr};
  int n_conv_tiles[2] = {0, 0};
  size_t conv_smem = 0;
};

// FFT-convolution work lists; independent of the sampling mode (the evaluation region is)
struct FftTables {
  int4* rows = nullptr; int n_rows = 0;          // image rows {src, plane, row0, nrows}
  int4* rows_psf = nullptr; int n_rows_psf = 0;  // PSF rows {src, -1-k, row0, nrows}
  int4* jobs = nullptr;                          // {src, in_plane | -1-k, kernel, out_plane}
  int4* cols_psf = nullptr; int n_cols_psf = 0;  // {job, kx0, ncols, 0}
  int4* cols_img = nullptr; int n_cols_img = 0;
  int4* rows_inv = nullptr; int n_rows_inv = 0;  // {src, out_plane, row0, nrows}
};

struct apb_plan {
  int n_src = 0, n_img = 0, n_par = 0, n_psf = 0;
  std::vector<DevSrc> h_src;
  std::vector<apb_image_t> h_img;
  DevSrc* d_src = nullptr;
  DevDyn* d_dyn = nullptr;
  apb_image_t* d_img = nullptr;
  apb_param_t* d_par = nullptr;
  apb_psf_t* d_psf = nullptr;
  ModeTables mt[2];
  FftTables ft[2];               // [grad]
  FftDesc* d_fftdesc = nullptr;
  cpx *d_twid = nullptr, *d_spec = nullptr;
  size_t fft_smem_rows = 0, fft_smem_cols = 0;
  int fft_nt_rows = 256, fft_nt_cols = 256;   // threads per CTA of the row / column kernels
  int n_fft_src = 0;
  cudaStream_t side = nullptr;        // PSF branch (stamp + spectrum) runs here, concurrent with the profile sampling
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int* psf_list = nullptr; int n_psf_list = 0;      // sources needing a shifted PSF stamp
  int* point_list = nullptr; int n_point = 0;
  int* up_list = nullptr; int n_up = 0;             // sources on a super-sampled grid (k_reduce_up)
  int* norm_list = nullptr; int n_norm = 0;
  int* amp_list = nullptr; int n_amp = 0;     // APB_F_AMP sources
  bool any_threshold = false, all_same_geo = true;
  int max_depth = 1;
  int NVp_grad = 1;
  bool fp32 = false;               // profile kernels in single precision (opts.flags bit 4; AP_config.ap_dtype = float32)
  bool use_coop = false;           // fused integration kernel (k_integrate) instead of per-depth launches
  int refine_lanes = 16;           // lanes sharing one queue entry
  int integrate_grid[2] = {148 * 4, 148 * 3};   // [grad] persistent CTAs of k_integrate (one resident wave)
  int pool_grid[2] = {148 * 4, 148 * 3};        // [grad] persistent CTAs of k_integrate_pool
  int pool_min = 150000;                        // depth-1 queue length from which the pooled kernel takes over
  int pool_g2 = 25, pool_nv = 8;                // largest gridding^2 / values per child of the plan
  bool pool_ok = false;
  int* h_qlast = nullptr;   // pinned: depth-1 queue length of the last pass per mode (-1: unknown), read back asynchronously
  // arenas
  double *d_stamp = nullptr, *d_out = nullptr, *d_psfst = nullptr, *d_meanpart = nullptr, *d_skyJ = nullptr;
  // queues
  Queues q{};
  std::vector<void*> owned;
  // image tiles + bins
  int4* img_tiles = nullptr; int n_img_tiles = 0;
  int *bin_ptr = nullptr, *bin_src = nullptr;
  double* d_chipart = nullptr;
  unsigned int* d_done = nullptr;   // CTA completion counter of k_assemble (last CTA sums the partials)
  // per-image buffers
  std::vector<double*> h_model, h_resid, h_resid2;
  double **d_model = nullptr, **d_resid = nullptr, **d_resid2 = nullptr, **d_userptr = nullptr;
  // normal equations
  BlockItem* d_items = nullptr; int n_items = 0;
  BlockDesc* d_blocks = nullptr; int n_blocks = 0;
  BlockItem* d_vitems = nullptr; int n_vitems = 0;   // diagonal items only (J^T v)
  BlockDesc* d_vblocks = nullptr; int n_vblocks = 0;
  // fixed-order gather of the block totals into the normal equations (k_block_gather): [0] full build, [1] vector only
  GatherDst* d_gdst[2] = {nullptr, nullptr}; int* d_gsrc[2] = {nullptr, nullptr}; int n_gdst[2] = {0, 0}, n_gdst_noH[2] = {0, 0};
  double* d_btot = nullptr; size_t btot_cap = 0;
  int *d_act_slot = nullptr, *d_act_off = nullptr;
  double* d_part = nullptr;
  size_t part_cap = 0;     // work items d_part has room for
  // block-sparse PCG solver (apb_solve.cuh): usable when no parameter is shared between sources
  bool sparse_ok = false;
  bool any_aux_psf = false;   // PSF stamps depend on a PSF-model source sampled in the same pass
  PcgRow* d_prows = nullptr; int n_prows = 0;
  PcgPass* d_ppasses = nullptr; int n_ppass = 0;
  int *d_pack_src = nullptr, *d_pslots = nullptr;
  double* d_packed = nullptr; long long n_ppacked = 0;
  PcgItem* d_pitems = nullptr; int n_pitems = 0;
  double *d_bvals = nullptr, *d_diagH = nullptr, *d_pfac = nullptr, *d_pvec = nullptr;
  double* d_bvals_own = nullptr;   // the plan's own [blocks | diag H] array (d_bvals may point at a caller's, apb_plan_bind_blocks)
  long long n_cblocks = 0, n_bvals = 0;   // owner blocks; doubles they occupy (packed n_a x n_b each)
  int* d_multi_rows = nullptr; int n_multi = 0;
  double *d_qpart = nullptr, *d_pcg_part = nullptr;
  unsigned int* d_pcg_bar = nullptr;
  int *d_own_slot = nullptr, *d_own_off = nullptr;   // free-parameter lists of the owners (PCG rows are owner rows)
  int pcg_grid = 0;
  double *d_xtmp = nullptr, *d_xtmp2 = nullptr, *d_rpp = nullptr, *d_atmp = nullptr, *d_atmp2 = nullptr, *d_rec2 = nullptr;
  cudaStream_t trial_stream = nullptr;   // concurrent chi^2 pass of apb_lm_trial (on the second plan)
  cudaEvent_t ev_trial_fork = nullptr, ev_trial_join = nullptr;
  apb_stats_t stats{};
  long long launches = 0;
  cudaStream_t last_stream = nullptr;
  bool profiling = false;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  double prof_ms[K_COUNT] = {0};
  long long prof_n[K_COUNT] = {0};
  cudaEvent_t get_event() {
    if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void pbegin(int id, cudaStream_t st) {
    if (!profiling) return;
    ProfRec r{id, get_event(), get_event()};
    cudaEventRecord(r.a, st);
    prof.push_back(r);
  }
  void pend(cudaStream_t st) {
    if (!profiling) return;
    cudaEventRecord(prof.back().b, st);
  }
  int first_evals[2] = {0, 0};
  long long cum_passes[2] = {0, 0}, cum_first[2] = {0, 0};
};

static int ceil_div(int a, int b) { return (a + b - 1) / b; }


extern "C" const char* apb_last_error(void) { return g_err.c_str(); }
static int fft_len(int n);
extern "C" int apb_fft_length(int n) { return n > 0 ? fft_len(n) : 0; }
extern "C" int apb_version(void) { return 103; }

static void gauss_legendre(int n, double* x, double* w) {
  // Newton iteration on P_n (same nodes as scipy.special.roots_legendre to rounding)
  for (int i = 0; i < n; ++i) {
    double z = cos(APB_PI * (i + 0.75) / (n + 0.5));
    double pp = 0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) {
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1);
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double z1 = z;
      z = z1 - p1 / pp;
      if (fabs(z - z1) < 1e-16) break;
    }
    x[n - 1 - i] = z;
    w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
  if (n % 2 == 1) x[n / 2] = 0.0;
}

static void set_geo(Geo& g, const int* out, const int* work, int bx, int by, bool ring) {
  g.ex0 = out[0] - bx; g.ey0 = out[1] - by; g.ew = out[2] + 2 * bx; g.eh = out[3] + 2 * by;
  g.rx0 = work[0] - bx; g.ry0 = work[1] - by; g.rw = work[2] + 2 * bx; g.rh = work[3] + 2 * by;
  const int r = ring ? 1 : 0;
  const int x0 = std::max(g.ex0 - r, g.rx0), y0 = std::max(g.ey0 - r, g.ry0);
  const int x1 = std::min(g.ex0 + g.ew + r, g.rx0 + g.rw), y1 = std::min(g.ey0 + g.eh + r, g.ry0 + g.rh);
  g.mx0 = x0; g.my0 = y0; g.mw = x1 - x0; g.mh = y1 - y0;
  g.tile0 = g.ntile = g.chunk0 = g.nchunk = 0;
}

// Stage radices of a length-N transform: 16s first (their index arithmetic is shifts and masks),
// then the remaining power of two, then 9/3/5.
static FftDesc fft_desc(int N) {
  FftDesc D;
  memset(&D, 0, sizeof(D));
  D.N = N;
  int n = N;
  while (n % 16 == 0) { D.radix[D.nstage++] = 16; n /= 16; }
  if (n % 8 == 0) { D.radix[D.nstage++] = 8; n /= 8; }
  if (n % 4 == 0) { D.radix[D.nstage++] = 4; n /= 4; }
  if (n % 2 == 0) { D.radix[D.nstage++] = 2; n /= 2; }
  while (n % 9 == 0) { D.radix[D.nstage++] = 9; n /= 9; }
  while (n % 3 == 0) { D.radix[D.nstage++] = 3; n /= 3; }
  while (n % 5 == 0) { D.radix[D.nstage++] = 5; n /= 5; }
  if (n != 1) D.nstage = 0;   // not 2-3-5 smooth
  return D;
}

// transform length for a padded stamp of n pixels: the 5-smooth m in [n, 1.3 n] with the lowest
// m * (stages + 1.5) (every stage is one shared-memory exchange of the whole sequence)
static int fft_len(int n) {
  int best = 0;
  double best_cost = 0.0;
  for (int m = std::max(n, 2); m <= std::max(n, 2) * 13 / 10 + 16; ++m) {
    const FftDesc D = fft_desc(m);
    if (D.nstage == 0 || D.nstage > APB_FFT_MAX_STAGE) continue;
    const double cost = (double)m * (D.nstage + 1.5);
    if (!best || cost < best_cost) { best = m; best_cost = cost; }
  }
  return best;
}

extern "C" int apb_plan_destroy(apb_plan_t* p) {
  if (!p) return 0;
  for (void* q : p->owned) cudaFree(q);
  if (p->h_qlast) cudaFreeHost(p->h_qlast);
  if (p->side) cudaStreamDestroy(p->side);
  if (p->trial_stream) cudaStreamDestroy(p->trial_stream);
  if (p->ev_trial_fork) cudaEventDestroy(p->ev_trial_fork);
  if (p->ev_trial_join) cudaEventDestroy(p->ev_trial_join);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  for (int d = 1; d <= APB_MAX_DEPTH; ++d) {
    Level& L = p->q.lv[d];
    void* olds[6] = {L.src, L.x, L.y, L.parent, L.child, L.res};
    for (void* o : olds)
      if (o) cudaFree(o);
  }
  delete p;
  return 0;
}

extern "C" int apb_plan_set_image_data(apb_plan_t* p, int image, const double* data, const double* weight,
                                       const uint8_t* mask, void* stream) {
  if (!p) APB_FAIL("plan is NULL");
  if (image < 0 || image >= p->n_img) APB_FAIL("apb_plan_set_image_data: image index out of range");
  apb_image_t& im = p->h_img[image];
  im.data = data; im.weight = weight; im.mask = mask;
  // pageable source: the copy is staged before the call returns, and ordered on `stream` like a kernel
  CU(cudaMemcpyAsync(p->d_img + image, &im, sizeof(apb_image_t), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}

template <typename T>
static int own_upload(apb_plan* p, const std::vector<T>& v, T** out) {
  int rc = upload(v, out);
  if (rc == 0) p->owned.push_back(*out);
  return rc;
}
static int own_alloc(apb_plan* p, void** out, size_t bytes) {
  CU(cudaMalloc(out, std::max<size_t>(bytes, 8)));
  p->owned.push_back(*out);
  return 0;
}

// (re)allocate the refinement queues with caps[d] entries at depth d (0 = keep)
static int alloc_queues(apb_plan* p, const long long* caps) {
  for (int d = 1; d <= p->max_depth; ++d) {
    const long long cap = std::min<long long>(caps[d], 1LL << 28);
    if (cap <= 0 || cap == p->q.cap[d]) continue;
    Level& L = p->q.lv[d];
    void* olds[6] = {L.src, L.x, L.y, L.parent, L.child, L.res};
    for (void* o : olds)
      if (o) cudaFree(o);
    memset(&L, 0, sizeof(L));
    p->q.cap[d] = 0;
    CU(cudaMalloc((void**)&L.src, sizeof(int) * cap));
    CU(cudaMalloc((void**)&L.x, sizeof(double) * cap));
    CU(cudaMalloc((void**)&L.y, sizeof(double) * cap));
    CU(cudaMalloc((void**)&L.parent, sizeof(int) * cap));
    CU(cudaMalloc((void**)&L.child, sizeof(int) * cap));
    CU(cudaMalloc((void**)&L.res, sizeof(double) * cap * p->NVp_grad));
    p->q.cap[d] = (int)cap;
  }
  return 0;
}

// Grow the refinement queues after an overflow.  caps[d] (d = 1..APB_MAX_DEPTH) = entries wanted
// at depth d; NULL = size every level from the counts of the last call (x1.5).  Clears the
// overflow flag.  Synchronous.
extern "C" int apb_plan_reserve(apb_plan_t* p, const int64_t* caps) {
  if (!p) APB_FAIL("apb_plan_reserve: NULL plan");
  CU(cudaDeviceSynchronize());
  long long want[APB_MAX_DEPTH + 1] = {0};
  if (caps) {
    for (int d = 1; d <= APB_MAX_DEPTH; ++d) want[d] = caps[d];
  } else {
    int cnt[APB_MAX_DEPTH + 2];
    CU(cudaMemcpy(cnt, p->q.count, sizeof(cnt), cudaMemcpyDeviceToHost));
    for (int d = 1; d <= p->max_depth; ++d) {
      // a level fed by a truncated parent has not seen all its entries yet: leave it head-room
      const long long need = (long long)cnt[d] + cnt[d] / 2;
      want[d] = std::max<long long>(p->q.cap[d], need);
      if (d > 1 && cnt[d - 1] > p->q.cap[d - 1]) want[d] = std::max<long long>(want[d], 2LL * p->q.cap[d]);
    }
  }
  if (p->any_threshold) {
    int rc = alloc_queues(p, want);
    if (rc) return rc;
  }
  CU(cudaMemset(p->q.overflow, 0, sizeof(int)));
  return 0;
}

extern "C" int apb_plan_create(const apb_source_t* src, int n_src, const apb_image_t* img, int n_img,
                               const apb_psf_t* psf, int n_psf, const apb_param_t* par, int n_par,
                               const apb_opts_t* opts, apb_plan_t** out) {
  if (!out) APB_FAIL("apb_plan_create: out is NULL");
  *out = nullptr;
  if (n_src < 0 || n_img <= 0 || n_par < 0) APB_FAIL("apb_plan_create: bad counts");
  int dev_count = 0;
  CU(cudaGetDeviceCount(&dev_count));
  apb_plan* p = new apb_plan();
  p->n_src = n_src; p->n_img = n_img; p->n_par = n_par; p->n_psf = n_psf;
  p->h_img.assign(img, img + n_img);
#define PFAIL(msg) do { apb_plan_destroy(p); APB_FAIL(msg); } while (0)
#define PCU(call) do { int rc_ = [&]() -> int { CU(call); return 0; }(); if (rc_) { apb_plan_destroy(p); return rc_; } } while (0)
#define PRC(call) do { int rc_ = (call); if (rc_) { apb_plan_destroy(p); return rc_; } } while (0)

  // quadrature tables
  {
    QuadTab qt;
    memset(&qt, 0, sizeof(qt));
    for (int n = 1; n <= APB_MAX_QUAD; ++n) {
      double x[APB_MAX_QUAD], w[APB_MAX_QUAD];
      gauss_legendre(n, x, w);
      for (int i = 0; i < n; ++i) { qt.a[n][i] = x[i] / 2.0; qt.w[n][i] = w[i] / 2.0; }
    }
    PCU(cudaMemcpyToSymbol(c_quad, &qt, sizeof(qt)));
    static ApbMathTab mt;     // exp / log tables of the profile kernels (apb_math.cuh)
    apb_math_fill(&mt);
    PCU(cudaMemcpyToSymbol(g_mathtab, &mt, sizeof(mt)));
  }

  // ---- sources
  std::vector<DevSrc>& S = p->h_src;
  S.resize(n_src);
  long long stamp_total = 0, out_total = 0, psfst_total = 0, spec_total = 0;
  std::vector<int> psf_list, point_list, norm_list, amp_list, up_list, act_slot, act_off(n_src + 1, 0);
  int max_nact = 0;
  std::vector<FftDesc> fdescs;
  std::vector<cpx> twid;
  auto get_desc = [&](int N) -> int {
    for (size_t k = 0; k < fdescs.size(); ++k)
      if (fdescs[k].N == N) return (int)k;
    FftDesc D = fft_desc(N);
    D.tw_off = (long long)twid.size();
    for (int k = 0; k < N; ++k) {
      const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)N;
      twid.push_back(cpx{(double)cosl(a), (double)sinl(a)});
    }
    fdescs.push_back(D);
    return (int)fdescs.size() - 1;
  };
  apb_opts_t opts_env{};   // APB_PLAN_FLAGS ORs bits into opts.flags (experiments: kernel-family switches without a rebuild)
  if (const char* e = getenv("APB_PLAN_FLAGS")) {
    if (opts) opts_env = *opts;
    opts_env.flags |= atoi(e);
    opts = &opts_env;
  }
  const int conv_force = opts ? (opts->flags & 3) : 0;   // 1: direct everywhere, 2: FFT everywhere
  p->fp32 = opts && (opts->flags & 16);
  if (p->fp32 && (opts->flags & 4)) PFAIL("single-precision profile kernels need the fused integration (flags bit 2 unset)");
  auto is_aux_img = [&](int ii) { return (img[ii].flags & APB_IMG_AUX) != 0; };
  for (int k = 0; k < n_psf; ++k)
    if (psf[k].source >= 0) {
      if (psf[k].source >= n_src) PFAIL("psf source index out of range");
      const apb_source_t& ps = src[psf[k].source];
      if (ps.image < 0 || ps.image >= n_img || !is_aux_img(ps.image)) PFAIL("a PSF-model source must sit on an APB_IMG_AUX image");
      if (ps.psf >= 0) PFAIL("a PSF-model source cannot itself be PSF-convolved");
      if (img[ps.image].W != psf[k].w || img[ps.image].H != psf[k].h || ps.out[0] != 0 || ps.out[1] != 0 ||
          ps.out[2] != psf[k].w || ps.out[3] != psf[k].h)
        PFAIL("a PSF-model source must cover its aux image, which must have the PSF's shape");
      p->any_aux_psf = true;
    }
  for (int i = 0; i < n_src; ++i) {
    const apb_source_t& a = src[i];
    DevSrc& s = S[i];
    memset(&s, 0, sizeof(s));
    if (a.kind < 0 || a.kind > APB_PLANE_SKY) PFAIL("unknown source kind");
    if (a.kind == APB_PLANE_SKY && (a.n_elem != 5 || a.psf >= 0 || a.integrate_mode != APB_INTEGRATE_NONE))
      PFAIL("plane sky: 5 elements, no PSF, no sub-pixel integration");
    if (a.image < 0 || a.image >= n_img) PFAIL("source image index out of range");
    if (a.n_elem < 3 || a.n_elem > APB_MAX_ELEM) PFAIL("bad n_elem");
    // APB_F_AMP: the last element is the amplitude; the profile sees the elements before it
    const bool has_amp = (a.flags & APB_F_AMP) != 0;
    const int n_el = a.n_elem - (has_amp ? 1 : 0);
    if (has_amp && (a.kind == APB_POINT || a.kind == APB_FLAT_SKY || a.kind == APB_PLANE_SKY || a.psf >= 0 || n_el < 4))
      PFAIL("APB_F_AMP: a profile source without PSF whose last element is the amplitude");
    if (a.sampling_mode < APB_SAMPLE_MIDPOINT || a.sampling_mode > APB_SAMPLE_TRAPEZOID) PFAIL("unknown sampling_mode");
    if (a.max_depth < 1 || a.max_depth > APB_MAX_DEPTH) PFAIL("integrate_max_depth out of range (1..4)");
    if (a.quad_level < 1 || a.quad_level > APB_MAX_QUAD || a.quad_init < 1 || a.quad_init > APB_MAX_QUAD)
      PFAIL("quadrature level out of range (1..9)");
    if (a.gridding < 1 || a.gridding > 16) PFAIL("integrate_gridding out of range (1..16)");
    const apb_image_t& im = img[a.image];
    s.kind = a.kind; s.flags = a.flags; s.image = a.image; s.n_elem = n_el;
    s.ox = a.out[0]; s.oy = a.out[1]; s.ow = a.out[2]; s.oh = a.out[3];
    if (s.ow <= 0 || s.oh <= 0 || s.ox < 0 || s.oy < 0 || s.ox + s.ow > im.W || s.oy + s.oh > im.H)
      PFAIL("source output window outside its image");
    s.n_act = 0;
    for (int e = 0; e < a.n_elem; ++e) {
      s.slot[e] = a.slot[e]; s.cval[e] = a.cval[e];
      if (a.slot[e] >= n_par) PFAIL("parameter slot out of range");
      if (a.slot[e] >= 0) { s.plane[e] = ++s.n_act; act_slot.push_back(a.slot[e]); } else s.plane[e] = 0;
    }
    // (with APB_F_AMP the loop above has already filed the amplitude as pseudo-element n_elem: its plane is the last)
    s.n_elem_all = a.n_elem; s.psf_src = -1; s.n_pp = 0;
    s.amp_elem = has_amp ? n_el : -1;
    if (has_amp) amp_list.push_back(i);
    if (a.psf >= 0 && a.psf < n_psf && psf[a.psf].source >= 0 && a.kind != APB_FLAT_SKY) {
      // auxiliary PSF model: its free parameters become pseudo-elements (and derivative planes) of this source
      if (a.kind == APB_POINT) PFAIL("point sources with a PSF model are not supported");
      s.psf_src = psf[a.psf].source;
      const apb_source_t& ps = src[s.psf_src];
      for (int e = 0; e < ps.n_elem; ++e)
        if (ps.slot[e] >= 0) {
          if (s.n_elem_all >= APB_MAX_ELEM) PFAIL("too many parameters for one source (own + auxiliary PSF model)");
          s.slot[s.n_elem_all] = ps.slot[e];
          s.plane[s.n_elem_all] = ++s.n_act;
          act_slot.push_back(ps.slot[e]);
          ++s.n_elem_all; ++s.n_pp;
        }
    }
    act_off[i + 1] = (int)act_slot.size();
    max_nact = std::max(max_nact, s.n_act);
    s.n_prof = a.n_prof;
    if (a.kind == APB_SPLINE && (a.n_prof < 2 || a.n_prof > APB_MAX_PROF || n_el != 4 + a.n_prof))
      PFAIL("spline source needs 2..20 nodes and n_elem = 4 + n_prof");
    for (int k = 0; k < a.n_prof && k < APB_MAX_PROF; ++k) s.prof[k] = a.prof[k];
    s.sampling_mode = a.sampling_mode; s.quad_init = a.quad_init; s.integrate_mode = a.integrate_mode;
    s.quad_level = a.quad_level; s.gridding = a.gridding; s.max_depth = a.max_depth; s.ref_mode = a.ref_mode;
    s.tol = a.tolerance; s.soft2 = a.softening * a.softening;
    {
      const int G = a.gridding;
      s.gsc[0] = s.gasc[0] = s.gsc[1] = s.gasc[1] = 1.0;
      for (int dd = 2; dd <= APB_MAX_DEPTH + 1; ++dd) {
        s.gsc[dd] = s.gsc[dd - 1] / (double)G;
        s.gasc[dd] = s.gasc[dd - 1] / (double)(G * G);
      }
      for (int k = 0; k < 16; ++k) s.goff[k] = -(G - 1) / (2.0 * G) + (double)k / G;
      s.g2d = (double)(G * G);
      s.gmagic = (65536 + G - 1) / G;
    }
    s.psf = a.psf; s.psf_shift = a.psf_shift;
    s.mask = a.mask; s.mask_x0 = a.mask_rect[0]; s.mask_y0 = a.mask_rect[1]; s.mask_w = a.mask_rect[2]; s.mask_h = a.mask_rect[3];
    if (a.mask && (a.mask_rect[2] <= 0 || a.mask_rect[3] <= 0)) PFAIL("source mask with an empty shape");
    // super-sampled PSF: the source lives on pixels 1 / up of the image's (window_object.py:233-239: pixelscale / up,
    // reference_imageij -> (rij + 0.5) up - 0.5); its windows are the image-pixel windows times up
    const int up = (a.upscale > 1 && a.kind != APB_FLAT_SKY) ? a.upscale : 1;
    if (up > 16) PFAIL("upscale out of range (1..16)");
    if (up > 1 && ((a.psf < 0 && !has_amp) || a.kind == APB_PLANE_SKY))
      PFAIL("upscale > 1 needs a PSF-convolved source, a point source, or a point source drawn from a PSF model (APB_F_AMP)");
    int fout[4], ffwd[4], fjac[4];
    for (int k = 0; k < 4; ++k) { fout[k] = a.out[k] * up; ffwd[k] = a.fwd[k] * up; fjac[k] = a.jac[k] * up; }
    s.up = up; s.fox = fout[0]; s.foy = fout[1]; s.fow = fout[2]; s.foh = fout[3];
    for (int k = 0; k < 4; ++k) s.S[k] = up > 1 ? im.S[k] / up : im.S[k];
    const double det = s.S[0] * s.S[3] - s.S[1] * s.S[2];
    if (det == 0.0) PFAIL("singular pixelscale");
    s.Sinv[0] = s.S[3] / det; s.Sinv[1] = -s.S[1] / det; s.Sinv[2] = -s.S[2] / det; s.Sinv[3] = s.S[0] / det;
    s.area = fabs(det);
    for (int k = 0; k < 2; ++k) s.rij[k] = up > 1 ? (im.rij[k] + 0.5) * up - 0.5 : im.rij[k];
    s.rxy[0] = im.rxy[0]; s.rxy[1] = im.rxy[1];
    s.bx = s.by = 0; s.out_off = -1; s.fine_off = -1; s.psf_off = -1;
    if (a.kind == APB_FLAT_SKY || a.kind == APB_POINT) s.integrate_mode = APB_INTEGRATE_NONE;
    if (a.kind == APB_POINT && a.psf < 0) PFAIL("point source without a PSF");
    if (a.kind == APB_FLAT_SKY) s.psf = -1;
    if (s.psf >= 0) {
      if (s.psf >= n_psf) PFAIL("psf index out of range");
      const int lanczos = a.psf_shift >= APB_SHIFT_LANCZOS ? a.psf_shift - APB_SHIFT_LANCZOS : 0;
      if (a.psf_shift != APB_SHIFT_NONE && a.psf_shift != APB_SHIFT_BILINEAR && (lanczos < 1 || lanczos > 8))
        PFAIL("unsupported psf_subpixel_shift");
      s.pw = psf[s.psf].w; s.ph = psf[s.psf].h;
      if (s.pw % 2 != 1 || s.ph % 2 != 1) PFAIL("psf must have odd shape");
      // the shifted stamp keeps the pad of the shift kernel (1 pixel bilinear, k pixels lanczos:k) except for point sources
      const int pad = (a.psf_shift == APB_SHIFT_NONE || a.kind == APB_POINT) ? 0 : (lanczos ? lanczos : 1);
      s.spw = s.pw + 2 * pad; s.sph = s.ph + 2 * pad;
      if (lanczos > 1 && a.kind != APB_POINT && (memcmp(a.out, a.fwd, sizeof(a.out)) || memcmp(a.out, a.jac, sizeof(a.out))))
        PFAIL("lanczos:k shifts of a PSF-convolved model need the model's window to be the window it is sampled on "
              "(the wider stamp wraps around the padded working image in the reference)");
      if (a.kind != APB_POINT) { s.bx = (s.pw + 2) / 2; s.by = (s.ph + 2) / 2; }  // ceil((1+P)/2), psf_image.py:71-93
      s.psf_off = psfst_total; psfst_total += (3LL + s.n_pp) * s.spw * s.sph;
      s.out_off = out_total; out_total += (long long)(1 + s.n_act) * s.ow * s.oh;
      s.fine_off = s.out_off;
      if (up > 1) { s.fine_off = out_total; out_total += (long long)(1 + s.n_act) * s.fow * s.foh; up_list.push_back(i); }
      psf_list.push_back(i);
      if (a.kind == APB_POINT) point_list.push_back(i);
    }
    if (up > 1 && s.psf < 0) {
      // a point source drawn from a super-sampled PSF model (point_source.py:122-143,181): sampled on the fine grid into
      // its stamp, summed into image-pixel planes of the out arena (fine_off stays -1: k_reduce_up reads the stamp)
      s.out_off = out_total; out_total += (long long)(1 + s.n_act) * s.ow * s.oh;
      up_list.push_back(i);
    }
    const bool ring = (a.kind != APB_FLAT_SKY && a.kind != APB_POINT &&
                       (s.sampling_mode == APB_SAMPLE_MIDPOINT || s.sampling_mode == APB_SAMPLE_TRAPEZOID) &&
                       s.integrate_mode == APB_INTEGRATE_THRESHOLD);
    set_geo(s.geo[0], fout, ffwd, s.bx, s.by, ring);
    set_geo(s.geo[1], fout, fjac, s.bx, s.by, ring);
    for (int m = 0; m < 2; ++m) {
      const Geo& g = s.geo[m];
      if (g.ex0 < g.rx0 || g.ey0 < g.ry0 || g.ex0 + g.ew > g.rx0 + g.rw || g.ey0 + g.eh > g.ry0 + g.rh)
        PFAIL("working window must contain the output window");
    }
    // ---- convolution method.  The reference convolves by FFT unless psf_convolve_mode="direct"
    //      (_model_methods.py:245-255); both give the same valid region, so "fft" here means
    //      "whichever is faster": direct tiles for small stamps, FFT above ~17x17.
    if (s.psf >= 0 && a.kind != APB_POINT) {
      int want = conv_force ? conv_force : a.conv_mode;
      if (want == 0) want = (s.spw * s.sph > 17 * 17) ? 2 : 1;
      if (s.psf_shift >= APB_SHIFT_LANCZOS) want = 1;   // circular over the padded image: the direct tile kernel wraps its loads
      if (want == 2) {
        const Geo& g = s.geo[0];
        const int Nx = fft_len(g.ew), Ny = fft_len(g.eh);
        // shared-memory sequences are skewed (FPAD): i -> i + i/16
        const int ldx = Nx + Nx / 16 + 1, ldy0 = Ny + Ny / 16 + 1;
        // columns per CTA: as many as keep two CTAs per SM resident (<= 112 KB each: landing / ping buffer, pong buffer,
        // PSF-spectrum tile), at least one
        // (APB_FFT_COL_KB / APB_FFT_NF_MAX / APB_FFT_NT_COLS / APB_FFT_NT_ROWS: tuning knobs for experiments)
        const char* ekb = getenv("APB_FFT_COL_KB");
        const size_t col_kb = ekb ? (size_t)atoi(ekb) : 112;
        auto col_need = [&](int n) { return (size_t)(2 * n * ldy0 + n * Ny) * sizeof(cpx); };
        int nc = 8;
        while (nc > 1 && col_need(nc) > col_kb * 1024) nc /= 2;
        // row transforms per CTA (each carries two real rows)
        const char* enf = getenv("APB_FFT_NF_MAX");
        int nf = std::max(1, std::min(enf ? atoi(enf) : 16, 4096 / Nx));
        while (nf > 1 && (size_t)(2 * nf * ldx) * sizeof(cpx) > 110 * 1024) --nf;
        if (const char* e = getenv("APB_FFT_NT_COLS")) p->fft_nt_cols = atoi(e);
        if (const char* e = getenv("APB_FFT_NT_ROWS")) p->fft_nt_rows = atoi(e);
        const int pad = 0;     // (the tile arrives by bulk copy: no transposing load whose banks would need a pad)
        const size_t col_bytes = col_need(nc);
        const size_t row_bytes = (size_t)(2 * nf * ldx) * sizeof(cpx);
        const bool fits = col_bytes <= 227 * 1024 && row_bytes <= 227 * 1024;
        if (fits) {
          s.conv_fft = 1;
          s.fftx = get_desc(Nx); s.ffty = get_desc(Ny);
          s.fft_nx = Nx; s.nxh = Nx / 2 + 1; s.nxp = (s.nxh + 3) & ~3;
          s.fft_nf = nf; s.fft_nc = nc; s.fft_ld = ldy0 + pad;
          // column-major spectra (apb_fft.cuh): a column is eh / oh / sph / Ny consecutive complex numbers
          s.specA_off = spec_total; spec_total += (long long)(1 + s.n_act) * s.nxh * g.eh;
          s.specB_off = spec_total; spec_total += (long long)(1 + s.n_act) * s.nxh * s.foh;
          s.specK_off = spec_total; spec_total += (3LL + s.n_pp) * s.nxh * s.sph;
          s.specKT_off = spec_total; spec_total += (3LL + s.n_pp) * s.nxh * Ny;
          p->fft_smem_rows = std::max(p->fft_smem_rows, row_bytes);
          p->fft_smem_cols = std::max(p->fft_smem_cols, col_bytes);
          p->n_fft_src++;
        } else if (conv_force == 2 || a.conv_mode == 2) {
          PFAIL("stamp too large for the shared-memory FFT convolution (transform length > ~7000)");
        }
      }
    }
    s.same_geo = memcmp(&s.geo[0], &s.geo[1], sizeof(Geo)) == 0;
    if (!s.same_geo && a.kind != APB_FLAT_SKY && a.kind != APB_POINT) p->all_same_geo = false;
    if (a.kind != APB_FLAT_SKY && a.kind != APB_POINT) {
      s.plane_stride = std::max((long long)s.geo[0].mw * s.geo[0].mh, (long long)s.geo[1].mw * s.geo[1].mh);
      s.stamp_off = stamp_total;
      stamp_total += s.plane_stride * (s.n_act + 2);   // value + derivatives + first-pass error
      if (s.integrate_mode == APB_INTEGRATE_THRESHOLD) { p->any_threshold = true; p->max_depth = std::max(p->max_depth, s.max_depth); }
      if (s.flags & APB_F_NORMALIZE) norm_list.push_back(i);
    }
  }
  p->NVp_grad = 1 + max_nact;

  // ---- per-mode tile / chunk / conv lists
  for (int m = 0; m < 2; ++m) {
    ModeTables& T = p->mt[m];
    std::vector<int4> tiles, chunks;
    std::vector<int> mean_list;
    for (int i = 0; i < n_src; ++i) {
      DevSrc& s = S[i];
      if (s.kind == APB_FLAT_SKY || s.kind == APB_POINT) continue;
      Geo& g = s.geo[m];
      g.tile0 = (int)tiles.size();
      for (int ty = 0; ty < g.mh; ty += 32)     // 32 x 32 pixels: a thread of k_first / k_select owns four rows
        for (int tx = 0; tx < g.mw; tx += 32) tiles.push_back(make_int4(i, tx, ty, 0));
      g.ntile = (int)tiles.size() - g.tile0;
      p->first_evals[m] += g.mw * g.mh;
      if (s.integrate_mode == APB_INTEGRATE_THRESHOLD && s.ref_mode == APB_REF_MEAN) {
        mean_list.push_back(i);
        g.chunk0 = (int)chunks.size();
        // row bands of the working region: >= 8192 pixels each, at most 128 per source ({src, row0, nrows, slot})
        const int rows = std::max(ceil_div(g.rh, 128), ceil_div(8192, std::max(g.rw, 1)));
        for (int r0 = 0; r0 < g.rh; r0 += rows)
          chunks.push_back(make_int4(i, r0, std::min(rows, g.rh - r0), (int)chunks.size()));
        g.nchunk = (int)chunks.size() - g.chunk0;
      }
    }
    T.n_tiles = (int)tiles.size(); T.n_chunks = (int)chunks.size(); T.n_mean = (int)mean_list.size();
    PRC(own_upload(p, tiles, &T.tiles));
    PRC(own_upload(p, chunks, &T.chunks));
    PRC(own_upload(p, mean_list, &T.mean_list));
    for (int gr = 0; gr < 2; ++gr) {
      std::vector<int4> jobs, ctiles;
      size_t smem = 0;
      for (int i = 0; i < n_src; ++i) {
        const DevSrc& s = S[i];
        if (s.psf < 0 || s.kind == APB_POINT || s.conv_fft) continue;
        const int cw = (s.spw - 1) / 2, chh = (s.sph - 1) / 2;
        if ((cw > s.bx || chh > s.by) && s.psf_shift < APB_SHIFT_LANCZOS) PFAIL("internal: psf stamp wider than the border");
        auto add_job = [&](int in_plane, int kern, int out_plane) {
          const int j = (int)jobs.size();
          jobs.push_back(make_int4(i, in_plane, kern, out_plane));
          for (int ty = 0; ty < s.foh; ty += CONV_TH)
            for (int tx = 0; tx < s.fow; tx += CONV_TW) ctiles.push_back(make_int4(j, tx, ty, 0));
        };
        add_job(0, 0, 0);
        if (gr) {
          const bool shifted = s.psf_shift != APB_SHIFT_NONE;
          for (int e = 0; e < s.n_elem; ++e) {
            if (s.plane[e] <= 0) continue;
            if (e < 2 && shifted) add_job(0, 1 + e, s.plane[e]);   // centre: through the PSF shift
            else add_job(s.plane[e], 0, s.plane[e]);
          }
          for (int k = 0; k < s.n_pp; ++k) add_job(0, 3 + k, s.plane[s.n_elem + k]);   // auxiliary PSF parameters
        }
        const size_t need = ((size_t)s.spw * s.sph + (size_t)(CONV_TH + s.sph - 1) * ((CONV_TW + s.spw - 1) | 1)) * sizeof(double);
        smem = std::max(smem, need);
      }
      if (smem > 227 * 1024) PFAIL("PSF too large for the direct convolution tile (max ~91x91)");
      T.conv_smem = std::max(T.conv_smem, smem);
      T.n_conv_tiles[gr] = (int)ctiles.size();
      PRC(own_upload(p, jobs, &T.conv_jobs[gr]));
      PRC(own_upload(p, ctiles, &T.conv_tiles[gr]));
    }
  }
  // The dynamic shared-memory ceiling is an attribute of the KERNEL, not of a launch: several plans with different
  // tile sizes live in one process, so every kernel is simply opted in to the device maximum (227 KB on sm_100).
  {
    int dev = 0, optin = 227 * 1024;
    PCU(cudaGetDevice(&dev));
    PCU(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    auto opt_in = [&](const void* fn) -> cudaError_t {
      cudaFuncAttributes fa;
      cudaError_t e = cudaFuncGetAttributes(&fa, fn);
      if (e != cudaSuccess) return e;
      return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
    };
    PCU(opt_in((const void*)k_conv));
    PCU(opt_in((const void*)k_fft_rows));
    PCU(opt_in((const void*)k_fft_rows_inv));
    PCU(opt_in((const void*)k_fft_cols));
    PCU(opt_in((const void*)k_lm_solve_small));
  }

  // ---- FFT convolution work lists
  for (int gr = 0; gr < 2 && p->n_fft_src; ++gr) {
    FftTables& F = p->ft[gr];
    std::vector<int4> rows, rpsf, jobs, cpsf, cimg, rinv;
    for (int i = 0; i < n_src; ++i) {
      const DevSrc& s = S[i];
      if (!s.conv_fft) continue;
      const Geo& g = s.geo[0];
      const bool shifted = s.psf_shift != APB_SHIFT_NONE;
      auto add_rows = [&](std::vector<int4>& v, int code, int nrows) {
        for (int r = 0; r < nrows; r += 2 * s.fft_nf) v.push_back(make_int4(i, code, r, std::min(2 * s.fft_nf, nrows - r)));
      };
      auto add_cols = [&](std::vector<int4>& v, int job) {
        for (int k = 0; k < s.nxh; k += s.fft_nc) v.push_back(make_int4(job, k, std::min(s.fft_nc, s.nxh - k), 0));
      };
      bool need_k[3 + APB_MAX_ELEM] = {true};
      std::vector<int> in_planes(1, 0);
      std::vector<int4> conv;   // {in_plane, kernel, out_plane}
      conv.push_back(make_int4(0, 0, 0, 0));
      if (gr)
        for (int e = 0; e < s.n_elem; ++e) {
          if (s.plane[e] <= 0) continue;
          if (e < 2 && shifted) { need_k[1 + e] = true; conv.push_back(make_int4(0, 1 + e, s.plane[e], 0)); }
          else { in_planes.push_back(s.plane[e]); conv.push_back(make_int4(s.plane[e], 0, s.plane[e], 0)); }
        }
      if (gr)
        for (int k = 0; k < s.n_pp; ++k) { need_k[3 + k] = true; conv.push_back(make_int4(0, 3 + k, s.plane[s.n_elem + k], 0)); }
      for (int k = 0; k < 3 + s.n_pp; ++k)
        if (need_k[k]) {
          add_rows(rpsf, -1 - k, s.sph);
          jobs.push_back(make_int4(i, -1 - k, 0, 0));
          add_cols(cpsf, (int)jobs.size() - 1);
        }
      for (int pl : in_planes) add_rows(rows, pl, g.eh);
      for (const int4& c : conv) {
        jobs.push_back(make_int4(i, c.x, c.y, c.z));
        add_cols(cimg, (int)jobs.size() - 1);
        add_rows(rinv, c.z, s.foh);
      }
    }
    F.n_rows = (int)rows.size(); F.n_rows_psf = (int)rpsf.size(); F.n_cols_psf = (int)cpsf.size(); F.n_cols_img = (int)cimg.size(); F.n_rows_inv = (int)rinv.size();
    PRC(own_upload(p, rows, &F.rows));
    PRC(own_upload(p, rpsf, &F.rows_psf));
    PRC(own_upload(p, jobs, &F.jobs));
    PRC(own_upload(p, cpsf, &F.cols_psf));
    PRC(own_upload(p, cimg, &F.cols_img));
    PRC(own_upload(p, rinv, &F.rows_inv));
  }
  if (p->n_fft_src) {
    PRC(own_upload(p, fdescs, &p->d_fftdesc));
    PRC(own_upload(p, twid, &p->d_twid));
    PRC(own_alloc(p, (void**)&p->d_spec, sizeof(cpx) * (size_t)spec_total));
  }

  // ---- image tiles and source bins (32x32 pixels)
  {
    std::vector<int4> itiles;
    std::vector<int> bptr(1, 0), bsrc;
    for (int ii = 0; ii < n_img; ++ii) {
      if (is_aux_img(ii)) continue;   // PSF-model grids are no output and no chi^2 term
      const int ntx = ceil_div(img[ii].W, 32), nty = ceil_div(img[ii].H, 32);
      std::vector<std::vector<int>> bins((size_t)ntx * nty);
      for (int i = 0; i < n_src; ++i) {
        const DevSrc& s = S[i];
        if (s.image != ii) continue;
        for (int ty = s.oy / 32; ty <= (s.oy + s.oh - 1) / 32; ++ty)
          for (int tx = s.ox / 32; tx <= (s.ox + s.ow - 1) / 32; ++tx) bins[(size_t)ty * ntx + tx].push_back(i);
      }
      for (int ty = 0; ty < nty; ++ty)
        for (int tx = 0; tx < ntx; ++tx) {
          itiles.push_back(make_int4(ii, tx * 32, ty * 32, (int)bptr.size() - 1));
          const auto& b = bins[(size_t)ty * ntx + tx];
          bsrc.insert(bsrc.end(), b.begin(), b.end());
          bptr.push_back((int)bsrc.size());
        }
    }
    p->n_img_tiles = (int)itiles.size();
    PRC(own_upload(p, itiles, &p->img_tiles));
    PRC(own_upload(p, bptr, &p->bin_ptr));
    PRC(own_upload(p, bsrc, &p->bin_src));
    PRC(own_alloc(p, (void**)&p->d_chipart, sizeof(double) * 2 * itiles.size()));
    PRC(own_alloc(p, (void**)&p->d_done, sizeof(unsigned int)));
    PCU(cudaMemset(p->d_done, 0, sizeof(unsigned int)));
  }

  // ---- normal-equation work lists: diagonal blocks + overlapping pairs, split in <=8-plane
  //      sub-blocks and <=2048-pixel rectangles (one CTA each)
  {
    std::vector<BlockItem> items, vitems;
    std::vector<BlockDesc> blocks, vblocks;
    auto add_block = [&](int a, int b, int pa0, int na, int pb0, int nb, int diag, int x0, int y0, int w, int h,
                         std::vector<BlockItem>& it, std::vector<BlockDesc>& bl) {
      BlockDesc bd{a, b, pa0, na, pb0, nb, diag, (int)it.size(), 0, -1, 0, 0};
      const int rows = std::max(1, 2048 / std::max(w, 1));
      for (int r = 0; r < h; r += rows) {
        BlockItem bi{a, b, pa0, na, pb0, nb, x0, y0 + r, w, std::min(rows, h - r), diag, (int)bl.size()};
        it.push_back(bi);
      }
      bd.nitem = (int)it.size() - bd.item0;
      bl.push_back(bd);
    };
    // spatial hash of sources per image to find overlapping pairs
    for (int ii = 0; ii < n_img; ++ii) {
      if (is_aux_img(ii)) continue;
      std::vector<int> ids;
      for (int i = 0; i < n_src; ++i)
        if (S[i].image == ii && S[i].n_act > 0) ids.push_back(i);
      for (int a : ids) {
        const DevSrc& A = S[a];
        for (int pa0 = 0; pa0 < A.n_act; pa0 += NB_MAX) {
          const int na = std::min(NB_MAX, A.n_act - pa0);
          add_block(a, a, pa0, na, pa0, na, 1, A.ox, A.oy, A.ow, A.oh, vitems, vblocks);
          for (int pb0 = pa0; pb0 < A.n_act; pb0 += NB_MAX) {
            const int nb = std::min(NB_MAX, A.n_act - pb0);
            add_block(a, a, pa0, na, pb0, nb, pa0 == pb0, A.ox, A.oy, A.ow, A.oh, items, blocks);
          }
        }
      }
      // pairs via a coarse grid
      const int CELL = 64;
      const int ncx = ceil_div(img[ii].W, CELL), ncy = ceil_div(img[ii].H, CELL);
      std::vector<std::vector<int>> cells((size_t)ncx * ncy);
      std::vector<int> big;   // sources covering many cells (sky): pair with everything directly
      for (int a : ids) {
        const DevSrc& A = S[a];
        const int c0x = A.ox / CELL, c1x = (A.ox + A.ow - 1) / CELL, c0y = A.oy / CELL, c1y = (A.oy + A.oh - 1) / CELL;
        if ((long long)(c1x - c0x + 1) * (c1y - c0y + 1) > 64) { big.push_back(a); continue; }
        for (int cy = c0y; cy <= c1y; ++cy)
          for (int cx = c0x; cx <= c1x; ++cx) cells[(size_t)cy * ncx + cx].push_back(a);
      }
      std::vector<std::pair<int, int>> pairs;
      for (auto& c : cells)
        for (size_t u = 0; u < c.size(); ++u)
          for (size_t v = u + 1; v < c.size(); ++v) pairs.emplace_back(std::min(c[u], c[v]), std::max(c[u], c[v]));
      for (int a : big)
        for (int b : ids)
          if (b != a && !(std::find(big.begin(), big.end(), b) != big.end() && b < a))
            pairs.emplace_back(std::min(a, b), std::max(a, b));
      std::sort(pairs.begin(), pairs.end());
      pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
      for (auto& pr : pairs) {
        const DevSrc& A = S[pr.first];
        const DevSrc& B = S[pr.second];
        const int x0 = std::max(A.ox, B.ox), y0 = std::max(A.oy, B.oy);
        const int x1 = std::min(A.ox + A.ow, B.ox + B.ow), y1 = std::min(A.oy + A.oh, B.oy + B.oh);
        if (x1 <= x0 || y1 <= y0) continue;
        for (int pa0 = 0; pa0 < A.n_act; pa0 += NB_MAX)
          for (int pb0 = 0; pb0 < B.n_act; pb0 += NB_MAX)
            add_block(pr.first, pr.second, pa0, std::min(NB_MAX, A.n_act - pa0), pb0, std::min(NB_MAX, B.n_act - pb0), 0,
                      x0, y0, x1 - x0, y1 - y0, items, blocks);
      }
    }
    // ---- block-sparse structure of J^T W J for the PCG solver, laid out on the OWNERS (the models of the whole
    //      fit; without an owner table every source is its own owner): row blocks = (owner, plane chunk), one
    //      block per (owner pair, chunk pair) that overlaps anywhere -- the same list on every rank of a fit
    //      sharded by image tile.  Every local block adds into its owner block (BlockDesc.cblock).
    {
      const int RPS = APB_MAX_ELEM / NB_MAX + 1;
      const bool tabled = opts && opts->owners && opts->n_owners > 0;
      const int n_own = tabled ? opts->n_owners : n_src;
      std::vector<int> own_off(n_own + 1, 0), own_slot, owner_of(n_src, -1);
      struct Rect { int image, x0, y0, w, h; };
      std::vector<Rect> orect(n_own);
      bool ok = n_par > 0;
      if (tabled) {
        for (int o = 0; o < n_own; ++o) {
          const apb_owner_t& ow = opts->owners[o];
          if (ow.n_slot < 0 || ow.n_slot > APB_MAX_ELEM) PFAIL("owner table: bad n_slot");
          for (int k = 0; k < ow.n_slot; ++k) {
            if (ow.slot[k] < 0 || ow.slot[k] >= n_par) PFAIL("owner table: parameter slot out of range");
            own_slot.push_back(ow.slot[k]);
          }
          own_off[o + 1] = (int)own_slot.size();
          orect[o] = Rect{ow.image, ow.out[0], ow.out[1], ow.out[2], ow.out[3]};
        }
        for (int i = 0; i < n_src; ++i) {
          const int o = src[i].owner;
          if (o < 0 || o >= n_own) PFAIL("source owner index out of range");
          if (own_off[o + 1] - own_off[o] != S[i].n_act) PFAIL("a source and its owner disagree on the free parameters");
          for (int k = 0; k < S[i].n_act; ++k)
            if (own_slot[own_off[o] + k] != act_slot[act_off[i] + k]) PFAIL("a source and its owner disagree on the free parameters");
          owner_of[i] = o;
        }
      } else {
        own_slot = act_slot; own_off = act_off;
        for (int i = 0; i < n_src; ++i) { owner_of[i] = i; orect[i] = Rect{S[i].image, S[i].ox, S[i].oy, S[i].ow, S[i].oh}; }
      }
      {   // the PCG needs every free parameter in exactly one owner
        std::vector<int> uses(std::max(n_par, 1), 0);
        for (int sl : own_slot)
          if (++uses[sl] > 1) ok = false;
        for (int k = 0; k < n_par; ++k)
          if (uses[k] == 0) ok = false;   // a free parameter no model depends on: singular block, leave it to the dense path
      }
      if (ok) {
        struct CB { int oa, ob, pa0, na, pb0, nb, diag; long long off; };
        std::vector<CB> cbs;
        std::unordered_map<unsigned long long, int> cb_of;
        const unsigned long long KR = (unsigned long long)n_own * RPS;
        auto key = [&](int oa, int ca, int ob, int cb) { return ((unsigned long long)oa * RPS + ca) * KR + ((unsigned long long)ob * RPS + cb); };
        auto nact = [&](int o) { return own_off[o + 1] - own_off[o]; };
        auto add_cb = [&](int oa, int pa0, int ob, int pb0, int diag) {
          cb_of[key(oa, pa0 / NB_MAX, ob, pb0 / NB_MAX)] = (int)cbs.size();
          CB c{oa, ob, pa0, std::min(NB_MAX, nact(oa) - pa0), pb0, std::min(NB_MAX, nact(ob) - pb0), diag, p->n_bvals};
          p->n_bvals += (long long)c.na * c.nb;
          cbs.push_back(c);
        };
        for (int o = 0; o < n_own; ++o)
          for (int pa0 = 0; pa0 < nact(o); pa0 += NB_MAX)
            for (int pb0 = pa0; pb0 < nact(o); pb0 += NB_MAX) add_cb(o, pa0, o, pb0, pa0 == pb0);
        // overlapping owner pairs per (uncut) image, via a coarse grid
        std::vector<int> order(n_own);
        for (int o = 0; o < n_own; ++o) order[o] = o;
        std::stable_sort(order.begin(), order.end(), [&](int u, int v) { return orect[u].image < orect[v].image; });
        for (size_t g0 = 0; g0 < order.size();) {
          size_t g1 = g0;
          while (g1 < order.size() && orect[order[g1]].image == orect[order[g0]].image) ++g1;
          std::vector<int> ids;
          int maxx = 1, maxy = 1;
          for (size_t k = g0; k < g1; ++k) {
            const int o = order[k];
            if (nact(o) == 0 || orect[o].w <= 0 || orect[o].h <= 0) continue;
            ids.push_back(o);
            maxx = std::max(maxx, orect[o].x0 + orect[o].w); maxy = std::max(maxy, orect[o].y0 + orect[o].h);
          }
          g0 = g1;
          const int CELL = 64;
          const int ncx = ceil_div(maxx, CELL), ncy = ceil_div(maxy, CELL);
          std::vector<std::vector<int>> cells((size_t)ncx * ncy);
          std::vector<int> big;   // owners covering many cells (sky): pair with everything directly
          for (int a : ids) {
            const Rect& A = orect[a];
            const int c0x = std::max(A.x0, 0) / CELL, c1x = (A.x0 + A.w - 1) / CELL, c0y = std::max(A.y0, 0) / CELL, c1y = (A.y0 + A.h - 1) / CELL;
            if ((long long)(c1x - c0x + 1) * (c1y - c0y + 1) > 64) { big.push_back(a); continue; }
            for (int cy = c0y; cy <= c1y; ++cy)
              for (int cx = c0x; cx <= c1x; ++cx) cells[(size_t)cy * ncx + cx].push_back(a);
          }
          std::vector<std::pair<int, int>> pairs;
          for (auto& c : cells)
            for (size_t u = 0; u < c.size(); ++u)
              for (size_t v = u + 1; v < c.size(); ++v) pairs.emplace_back(std::min(c[u], c[v]), std::max(c[u], c[v]));
          for (int a : big)
            for (int b : ids)
              if (b != a) pairs.emplace_back(std::min(a, b), std::max(a, b));
          std::sort(pairs.begin(), pairs.end());
          pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
          for (auto& pr : pairs) {
            const Rect& A = orect[pr.first];
            const Rect& B = orect[pr.second];
            if (std::min(A.x0 + A.w, B.x0 + B.w) <= std::max(A.x0, B.x0) || std::min(A.y0 + A.h, B.y0 + B.h) <= std::max(A.y0, B.y0)) continue;
            for (int pa0 = 0; pa0 < nact(pr.first); pa0 += NB_MAX)
              for (int pb0 = 0; pb0 < nact(pr.second); pb0 += NB_MAX) add_cb(pr.first, pa0, pr.second, pb0, 0);
          }
        }
        // local blocks -> owner blocks
        for (auto& bd : blocks) {
          int oa = owner_of[bd.a], ob = owner_of[bd.b], pa0 = bd.pa0, pb0 = bd.pb0, tr = 0;
          if (oa == ob && bd.a != bd.b) PFAIL("two pieces of one model overlap on one image");
          if (oa > ob) { std::swap(oa, ob); std::swap(pa0, pb0); tr = 1; }
          auto itb = cb_of.find(key(oa, pa0 / NB_MAX, ob, pb0 / NB_MAX));
          if (itb == cb_of.end()) PFAIL("owner table: two sources overlap but their owners' windows do not");
          bd.coff = cbs[itb->second].off; bd.cld = cbs[itb->second].nb; bd.ctrans = tr;
        }
        std::vector<PcgRow> prows;
        std::vector<std::vector<PcgEntry>> ents;
        std::vector<int> row_of((size_t)n_own * RPS, -1);
        for (size_t k = 0; k < cbs.size(); ++k) {
          const CB& cb = cbs[k];
          if (!cb.diag) continue;
          row_of[(size_t)cb.oa * RPS + cb.pa0 / NB_MAX] = (int)prows.size();
          prows.push_back(PcgRow{cb.oa, cb.pa0, cb.na, 0, 0, cb.off, own_off[cb.oa] + cb.pa0, 0});
          ents.emplace_back();
        }
        for (size_t k = 0; k < cbs.size(); ++k) {
          const CB& cb = cbs[k];
          const int ra = row_of[(size_t)cb.oa * RPS + cb.pa0 / NB_MAX], rb = row_of[(size_t)cb.ob * RPS + cb.pb0 / NB_MAX];
          PcgEntry ea{cb.off, cb.nb, 0, cb.nb, {0}}, eb{cb.off, cb.nb, 1, cb.na, {0}};
          for (int j = 0; j < cb.nb; ++j) ea.sl[j] = own_slot[own_off[cb.ob] + cb.pb0 + j];
          for (int j = 0; j < cb.na; ++j) eb.sl[j] = own_slot[own_off[cb.oa] + cb.pa0 + j];
          ents[ra].push_back(ea);
          if (!cb.diag) ents[rb].push_back(eb);
        }
        std::vector<PcgEntry> flat;
        std::vector<PcgItem> pitems;
        std::vector<int> multi_rows;
        // a row is one work item unless it has very many blocks (the sky row couples to every model): then it is
        // split into chunks whose partial rows the solver adds in order.  A row's widest blocks go first (the lanes of
        // one sweep then hold blocks of similar width: less padding in the packed copy).
        const int SPLIT = 64, CH = 8;
        for (size_t r = 0; r < prows.size(); ++r) {
          std::stable_sort(ents[r].begin(), ents[r].end(), [](const PcgEntry& x, const PcgEntry& y) { return x.n > y.n; });
          const int e0 = (int)flat.size();
          flat.insert(flat.end(), ents[r].begin(), ents[r].end());
          const int e1 = (int)flat.size();
          const bool multi = e1 - e0 > SPLIT;
          if (!multi) pitems.push_back(PcgItem{(int)r, e0, e1, 0, prows[r].n, prows[r].slot0});
          else for (int e = e0; e < e1; e += CH) pitems.push_back(PcgItem{(int)r, e, std::min(e + CH, e1), e == e0 ? 1 : 2, prows[r].n, prows[r].slot0});
          if (multi) multi_rows.push_back((int)r);
        }
        // lanes per item: enough for one sweep (8 / 16 / 32; NB_MAX at least: lane i of the group finishes row element i)
        auto lanes_of = [](const PcgItem& w) { const int e = w.e1 - w.e0; return e <= 8 ? 8 : e <= 16 ? 16 : 32; };
        // work items in the order the solver deals them: whole rows, widest groups and most blocks first (then tallest
        // first), then the chunks of the split rows (kept together: their partial rows are added in item order)
        {
          std::vector<int> ord(pitems.size());
          for (size_t k = 0; k < ord.size(); ++k) ord[k] = (int)k;
          std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) {
            const PcgItem &wx = pitems[x], &wy = pitems[y];
            if ((wx.multi != 0) != (wy.multi != 0)) return wx.multi == 0;
            if (wx.multi) return false;
            if (lanes_of(wx) != lanes_of(wy)) return lanes_of(wx) > lanes_of(wy);
            const int sx = (wx.e1 - wx.e0 + 31) / 32, sy = (wy.e1 - wy.e0 + 31) / 32;
            if (sx != sy) return sx > sy;
            return wx.n > wy.n;
          });
          std::vector<PcgItem> sorted(pitems.size());
          for (size_t k = 0; k < ord.size(); ++k) sorted[k] = pitems[ord[k]];
          pitems.swap(sorted);
          for (auto& r : prows) { r.item0 = -1; r.nitem = 0; }
          for (size_t k = 0; k < pitems.size(); ++k) {
            PcgRow& r = prows[pitems[k].rb];
            if (r.item0 < 0) r.item0 = (int)k;
            r.nitem++;
          }
        }
        // packed tables (PcgPass): consecutive items with the same group width share a warp
        std::vector<PcgPass> passes;
        std::vector<int> pack_src, pslots;
        for (size_t k0 = 0; k0 < pitems.size();) {
          const int G = lanes_of(pitems[k0]);
          int nit = 0, nsweep = 0, ni = 0, nj = 0;
          while (nit < 32 / G && k0 + nit < pitems.size() && lanes_of(pitems[k0 + nit]) == G) {
            const PcgItem& w = pitems[k0 + nit];
            nsweep = std::max(nsweep, (w.e1 - w.e0 + G - 1) / G);
            ni = std::max(ni, w.n);
            for (int e = w.e0; e < w.e1; ++e) nj = std::max(nj, flat[e].n);
            ++nit;
          }
          if ((long long)pack_src.size() + (long long)nsweep * ni * nj * 32 > 0x7fffffffLL) PFAIL("block-sparse solver: packed matrix too large");
          passes.push_back(PcgPass{(long long)pack_src.size(), (int)pslots.size(), ni, nj, nsweep, (int)k0, nit, G, 0});
          for (int sw = 0; sw < nsweep; ++sw) {
            const PcgEntry* le[32];
            int ln[32];
            for (int l = 0; l < 32; ++l) {
              le[l] = nullptr; ln[l] = 0;
              if (l / G >= nit) continue;
              const PcgItem& w = pitems[k0 + l / G];
              const int e = w.e0 + sw * G + l % G;
              if (e >= w.e1) continue;
              le[l] = &flat[e]; ln[l] = w.n;
            }
            for (int i = 0; i < ni; ++i)
              for (int j = 0; j < nj; ++j)
                for (int l = 0; l < 32; ++l) {
                  long long o = -1;
                  if (le[l] && i < ln[l] && j < le[l]->n) o = le[l]->off + (le[l]->transposed ? (long long)j * le[l]->ld + i : (long long)i * le[l]->ld + j);
                  if (o > 0x7fffffffLL) PFAIL("block-sparse solver: block values beyond the packed index range");
                  pack_src.push_back((int)o);
                }
            for (int j = 0; j < nj; ++j)
              for (int l = 0; l < 32; ++l) pslots.push_back(le[l] && j < le[l]->n ? le[l]->sl[j] : -1);
          }
          k0 += nit;
        }
        p->n_ppass = (int)passes.size(); p->n_ppacked = (long long)pack_src.size();
        if (getenv("APB_PCG_DEBUG")) fprintf(stderr, "pcg: %zu rows, %zu items, %zu passes, %zu packed doubles (%lld tight), %zu slots\n", prows.size(), pitems.size(), passes.size(), pack_src.size(), (long long)p->n_bvals, pslots.size());
        PRC(own_upload(p, passes, &p->d_ppasses));
        PRC(own_upload(p, pack_src, &p->d_pack_src));
        PRC(own_upload(p, pslots, &p->d_pslots));
        PRC(own_alloc(p, (void**)&p->d_packed, sizeof(double) * std::max<size_t>(pack_src.size(), 1)));
        p->n_prows = (int)prows.size(); p->n_pitems = (int)pitems.size();
        p->n_cblocks = (long long)cbs.size();
        PRC(own_upload(p, prows, &p->d_prows));
        PRC(own_upload(p, pitems, &p->d_pitems));
        PRC(own_upload(p, own_slot, &p->d_own_slot));
        PRC(own_upload(p, own_off, &p->d_own_off));
        PRC(own_upload(p, multi_rows, &p->d_multi_rows)); p->n_multi = (int)multi_rows.size();
        PRC(own_alloc(p, (void**)&p->d_qpart, sizeof(double) * 8 * std::max<size_t>(pitems.size(), 1)));
        PRC(own_alloc(p, (void**)&p->d_bvals_own, sizeof(double) * ((size_t)p->n_bvals + (size_t)n_par)));
        p->d_bvals = p->d_bvals_own; p->d_diagH = p->d_bvals + p->n_bvals;
        PRC(own_alloc(p, (void**)&p->d_pfac, sizeof(double) * 64 * std::max<size_t>(prows.size(), 1)));
        PRC(own_alloc(p, (void**)&p->d_pvec, sizeof(double) * 6 * (size_t)n_par));
        int dev = 0, sms = 148, per_sm = 1;
        PCU(cudaGetDevice(&dev));
        PCU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        PCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg, PCG_NT, 0));
        // enough warps for the work items, never more CTAs than can be co-resident (grid barrier)
        const int cap_sm = PCG_MINB;
        const int want = std::max(1, std::min(sms * std::min(per_sm, cap_sm), ceil_div(std::max(p->n_ppass, p->n_prows / 32 + 1), PCG_NT / 32)));
        p->pcg_grid = want;
        PRC(own_alloc(p, (void**)&p->d_pcg_part, sizeof(double) * 4 * (size_t)want));
        PRC(own_alloc(p, (void**)&p->d_pcg_bar, sizeof(unsigned int)));
        p->sparse_ok = true;
      } else if (tabled) {
        PFAIL("owner table: a free parameter belongs to no owner or to several");
      }
    }
    p->n_items = (int)items.size(); p->n_blocks = (int)blocks.size();
    p->n_vitems = (int)vitems.size(); p->n_vblocks = (int)vblocks.size();
    // where every value of every block goes (k_block_gather): contributions in block order per target entry
    for (int which = 0; which < 2; ++which) {
      const std::vector<BlockDesc>& bl = which ? vblocks : blocks;
      struct Con { int kind; long long index; int src; };
      std::vector<Con> cons;
      for (size_t bi = 0; bi < bl.size(); ++bi) {
        const BlockDesc& bd = bl[bi];
        const int* sa = act_slot.data() + act_off[bd.a] + bd.pa0;
        const int* sb = act_slot.data() + act_off[bd.b] + bd.pb0;
        for (int i = 0; i < bd.na; ++i) {
          const int vsrc = (int)(bi * BLK_VALS) + NB_MAX * NB_MAX + i;
          if (which || bd.diag) cons.push_back(Con{0, sa[i], vsrc});
          if (which) continue;
          for (int j = 0; j < bd.nb; ++j) {
            const int msrc = (int)(bi * BLK_VALS) + i * NB_MAX + j;
            cons.push_back(Con{3, (long long)sa[i] * n_par + sb[j], msrc});
            if (!bd.diag) cons.push_back(Con{3, (long long)sb[j] * n_par + sa[i], msrc});
            if (p->sparse_ok && bd.coff >= 0) {
              cons.push_back(Con{1, bd.coff + (bd.ctrans ? (long long)j * bd.cld + i : (long long)i * bd.cld + j), msrc});
              if (bd.diag && i == j) cons.push_back(Con{2, sa[i], msrc});
            }
          }
        }
      }
      std::stable_sort(cons.begin(), cons.end(), [](const Con& x, const Con& y) { return x.kind != y.kind ? x.kind < y.kind : x.index < y.index; });
      std::vector<GatherDst> gd;
      std::vector<int> gs(cons.size());
      int noH = 0;
      for (size_t k = 0; k < cons.size(); ++k) {
        gs[k] = cons[k].src;
        if (gd.empty() || gd.back().kind != cons[k].kind || gd.back().index != cons[k].index) {
          gd.push_back(GatherDst{cons[k].index, cons[k].kind, (int)k, (int)k, 0});
        }
        gd.back().s1 = (int)k + 1;
      }
      for (const GatherDst& d : gd) noH += d.kind != 3;
      p->n_gdst[which] = (int)gd.size(); p->n_gdst_noH[which] = noH;
      PRC(own_upload(p, gd, &p->d_gdst[which]));
      PRC(own_upload(p, gs, &p->d_gsrc[which]));
    }
    p->btot_cap = std::max(blocks.size(), vblocks.size());
    PRC(own_alloc(p, (void**)&p->d_btot, sizeof(double) * BLK_VALS * std::max<size_t>(p->btot_cap, 1)));
    PRC(own_upload(p, items, &p->d_items));
    PRC(own_upload(p, blocks, &p->d_blocks));
    PRC(own_upload(p, vitems, &p->d_vitems));
    PRC(own_upload(p, vblocks, &p->d_vblocks));
    PRC(own_upload(p, act_slot, &p->d_act_slot));
    PRC(own_upload(p, act_off, &p->d_act_off));
    p->part_cap = std::max(items.size(), vitems.size());
    PRC(own_alloc(p, (void**)&p->d_part, sizeof(double) * BLK_VALS * p->part_cap));
  }

  // ---- tables and arenas
  PRC(own_upload(p, S, &p->d_src));
  PRC(own_alloc(p, (void**)&p->d_dyn, sizeof(DevDyn) * (size_t)std::max(n_src, 1)));
  PRC(own_upload(p, p->h_img, &p->d_img));
  { std::vector<apb_param_t> pv(par, par + n_par); PRC(own_upload(p, pv, &p->d_par)); }
  { std::vector<apb_psf_t> pv(psf, psf + n_psf); PRC(own_upload(p, pv, &p->d_psf)); }
  PRC(own_upload(p, psf_list, &p->psf_list)); p->n_psf_list = (int)psf_list.size();
  PRC(own_upload(p, point_list, &p->point_list)); p->n_point = (int)point_list.size();
  PRC(own_upload(p, up_list, &p->up_list)); p->n_up = (int)up_list.size();
  PRC(own_upload(p, norm_list, &p->norm_list)); p->n_norm = (int)norm_list.size();
  PRC(own_upload(p, amp_list, &p->amp_list)); p->n_amp = (int)amp_list.size();
  PRC(own_alloc(p, (void**)&p->d_stamp, sizeof(double) * (size_t)stamp_total));
  PRC(own_alloc(p, (void**)&p->d_out, sizeof(double) * (size_t)out_total));
  PRC(own_alloc(p, (void**)&p->d_psfst, sizeof(double) * (size_t)psfst_total));
  PRC(own_alloc(p, (void**)&p->d_meanpart, sizeof(double) * (size_t)std::max(p->mt[0].n_chunks, p->mt[1].n_chunks)));
  PRC(own_alloc(p, (void**)&p->d_skyJ, sizeof(double) * (size_t)std::max(n_src, 1)));
  PRC(own_alloc(p, (void**)&p->d_xtmp, sizeof(double) * (size_t)std::max(n_par, 1)));
  PRC(own_alloc(p, (void**)&p->d_xtmp2, sizeof(double) * (size_t)std::max(n_par, 1)));
  PRC(own_alloc(p, (void**)&p->d_rpp, sizeof(double) * (size_t)std::max(n_par, 1)));
  PRC(own_alloc(p, (void**)&p->d_atmp, sizeof(double) * (size_t)std::max(n_par, 1)));
  PRC(own_alloc(p, (void**)&p->d_atmp2, sizeof(double) * (size_t)std::max(n_par, 1)));
  PRC(own_alloc(p, (void**)&p->d_rec2, sizeof(double) * 4));
  PCU(cudaMemset(p->d_stamp, 0, sizeof(double) * (size_t)std::max<long long>(stamp_total, 1)));

  // ---- queues: depth 1 can never hold more than the first-pass pixels; deeper levels start at a
  //      heuristic size and grow on demand (apb_plan_reserve) when the sticky overflow flag is raised
  {
    long long cap = opts ? opts->queue_capacity : 0;
    const long long px = std::max(p->first_evals[0], p->first_evals[1]);
    PRC(own_alloc(p, (void**)&p->q.count, sizeof(int) * (APB_MAX_DEPTH + 2)));
    PRC(own_alloc(p, (void**)&p->q.overflow, sizeof(int)));
    PCU(cudaMemset(p->q.overflow, 0, sizeof(int)));
    PCU(cudaMemset(p->q.count, 0, sizeof(int) * (APB_MAX_DEPTH + 2)));
    PRC(own_alloc(p, (void**)&p->q.cum, sizeof(unsigned long long) * 2 * (APB_MAX_DEPTH + 2)));
    PRC(own_alloc(p, (void**)&p->q.last_kind, sizeof(int)));
    PCU(cudaMemset(p->q.cum, 0, sizeof(unsigned long long) * 2 * (APB_MAX_DEPTH + 2)));
    PCU(cudaMemset(p->q.last_kind, 0, sizeof(int)));
    p->q.NVp = 1;
    if (p->any_threshold) {
      long long caps[APB_MAX_DEPTH + 1] = {0};
      for (int d = 1; d <= p->max_depth; ++d)
        caps[d] = cap > 0 ? cap : (d == 1 ? std::max<long long>(px, 1024) : std::max<long long>(px, 1 << 18));
      PRC(alloc_queues(p, caps));
    }
  }

  // ---- fused integration kernel (k_integrate): lanes sharing one depth-1 queue entry
  if (p->any_threshold && !(opts && (opts->flags & 4))) {
    p->use_coop = true;
    int qmax = 1;
    for (int i = 0; i < n_src; ++i)
      if (S[i].integrate_mode == APB_INTEGRATE_THRESHOLD) qmax = std::max(qmax, S[i].quad_level);
    int L = 1;
    while (L < qmax * qmax && L < 32) L <<= 1;
    p->refine_lanes = L;
    int dev = 0, sms = 148, b0 = 4, b1 = 4;
    PCU(cudaGetDevice(&dev));
    PCU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (p->fp32) {
      PCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_integrate<false, float>, 128, 0));
      PCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_integrate<true, float>, 128, 0));
    } else {
      PCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_integrate<false, double>, 128, 0));
      PCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_integrate<true, double>, 128, 0));
    }
    p->integrate_grid[0] = sms * std::max(1, b0);
    p->integrate_grid[1] = sms * std::max(1, b1);
    // throughput form (k_integrate_pool) for long queues
    int g2 = 1, nv = 1;
    for (int i = 0; i < n_src; ++i)
      if (S[i].integrate_mode == APB_INTEGRATE_THRESHOLD) {
        g2 = std::max(g2, S[i].gridding * S[i].gridding);
        nv = std::max(nv, 1 + S[i].n_elem);
      }
    p->pool_g2 = g2; p->pool_nv = nv;
    p->pool_ok = g2 * nv <= POOL_CSUM;
    if (const char* e = getenv("APB_POOL_MIN")) p->pool_min = atoi(e);
    if (opts && (opts->flags & 8)) p->pool_min = 0;
    if (p->fp32) {
      PCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_integrate_pool<false, float>, POOL_B, 0));
      PCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_integrate_pool<true, float>, POOL_B, 0));
    } else {
      PCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_integrate_pool<false, double>, POOL_B, 0));
      PCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_integrate_pool<true, double>, POOL_B, 0));
    }
    p->pool_grid[0] = sms * std::max(1, b0);
    p->pool_grid[1] = sms * std::max(1, b1);
  }

  PCU(cudaMallocHost((void**)&p->h_qlast, 2 * sizeof(int)));
  p->h_qlast[0] = p->h_qlast[1] = -1;
  PCU(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
  PCU(cudaStreamCreateWithFlags(&p->trial_stream, cudaStreamNonBlocking));
  PCU(cudaEventCreateWithFlags(&p->ev_trial_fork, cudaEventDisableTiming));
  PCU(cudaEventCreateWithFlags(&p->ev_trial_join, cudaEventDisableTiming));
  PCU(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
  PCU(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));

  // ---- per-image internal buffers
  p->h_model.resize(n_img); p->h_resid.resize(n_img); p->h_resid2.resize(n_img);
  for (int ii = 0; ii < n_img; ++ii) {
    const size_t n = (size_t)img[ii].H * img[ii].W;
    PRC(own_alloc(p, (void**)&p->h_model[ii], sizeof(double) * n));
    PRC(own_alloc(p, (void**)&p->h_resid[ii], sizeof(double) * n));
    PRC(own_alloc(p, (void**)&p->h_resid2[ii], sizeof(double) * n));
  }
  PRC(own_upload(p, p->h_model, &p->d_model));
  PRC(own_upload(p, p->h_resid, &p->d_resid));
  PRC(own_upload(p, p->h_resid2, &p->d_resid2));
  PRC(own_alloc(p, (void**)&p->d_userptr, sizeof(double*) * n_img));
  *out = p;
  return 0;
}

// ----------------------------------------------------------------------------
// one sampling pass: prep -> psf stamps -> first pass -> reference -> select -> refine -> scatter ->
// normalise -> point sources -> convolution.  Leaves the out-planes of every source ready.
// ----------------------------------------------------------------------------
#define PB(id) p->pbegin(id, st)
#define LAUNCH_CHECK() do { p->pend(st); p->launches++; g_launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { g_err = std::string("kernel launch: ") + cudaGetErrorString(e_) + " line " + std::to_string(__LINE__); return -3; } } while (0)

static int sample_pass(apb_plan* p, const double* x, int as_rep, int mode, int grad, cudaStream_t st) {
  const int n_src = p->n_src;
  if (n_src == 0) return 0;
  ModeTables& T = p->mt[mode];
  PB(K_PREP);
  k_prep<<<ceil_div(n_src, 4), 128, 0, st>>>(p->d_src, p->d_dyn, n_src, p->d_par, x, as_rep, p->q.count, p->d_skyJ, grad,
                                             p->q.cum, p->q.last_kind);
  LAUNCH_CHECK();
  p->cum_passes[grad ? 1 : 0]++;
  p->cum_first[grad ? 1 : 0] += p->first_evals[mode];
  // PSF branch on the side stream: shifted stamps and (FFT sources) their spectra depend only on
  // k_prep, so they overlap the first pass and the adaptive integration of the profiles
  const FftTables& F = p->ft[grad];
  // (while per-kernel timing is on, everything stays on one stream: an event pair around a side-stream
  //  kernel would also measure its wait for SMs held by the main stream's kernels)
  // With an auxiliary PSF model the stamps depend on a source sampled in this very pass: the PSF branch then runs
  // on the main stream after the sampling kernels.
  const bool fork = p->n_psf_list && !p->profiling && !p->any_aux_psf;
  auto psf_branch = [&]() -> int {
    cudaStream_t main_st = st;
    if (fork) {
      if (cudaEventRecord(p->ev_fork, main_st) != cudaSuccess || cudaStreamWaitEvent(p->side, p->ev_fork, 0) != cudaSuccess)
        APB_FAIL("fork to the PSF stream failed");
      st = p->side;
    }
    PB(K_PSF);
    k_psf_stamp<<<p->n_psf_list, 256, 0, st>>>(p->d_src, p->d_dyn, p->psf_list, p->d_psf, p->d_psfst, grad, mode, p->d_stamp,
                                               p->d_out);
    LAUNCH_CHECK();
    if (p->n_fft_src) {
      PB(K_FFTROWS);
      k_fft_rows<<<F.n_rows_psf, p->fft_nt_rows, p->fft_smem_rows, st>>>(p->d_src, p->d_fftdesc, p->d_twid, F.rows_psf, mode,
                                                                          p->d_stamp, p->d_psfst, p->d_spec);
      LAUNCH_CHECK();
      PB(K_FFTCOLS);
      k_fft_cols<<<F.n_cols_psf, p->fft_nt_cols, p->fft_smem_cols, st>>>(p->d_src, p->d_fftdesc, p->d_twid, F.jobs, F.cols_psf,
                                                                          mode, p->d_spec);
      LAUNCH_CHECK();
    }
    if (fork && cudaEventRecord(p->ev_join, st) != cudaSuccess) APB_FAIL("PSF stream event failed");
    st = main_st;
    return 0;
  };
  if (p->n_psf_list && !p->any_aux_psf) {
    const int rc = psf_branch();
    if (rc) return rc;
  }
  if (T.n_tiles) {
    PB(grad ? K_FIRST_G : K_FIRST);
    // (PROFILE_LAUNCH: the profile kernels exist in fp64 and in fp32 arithmetic, AP_config.ap_dtype)
#define PROFILE_LAUNCH(K, GRID, BLOCK, ...)                                                        \
    do {                                                                                           \
      if (p->fp32) { if (grad) K<true, float><<<GRID, BLOCK, 0, st>>>(__VA_ARGS__); else K<false, float><<<GRID, BLOCK, 0, st>>>(__VA_ARGS__); } \
      else { if (grad) K<true, double><<<GRID, BLOCK, 0, st>>>(__VA_ARGS__); else K<false, double><<<GRID, BLOCK, 0, st>>>(__VA_ARGS__); }      \
    } while (0)
    PROFILE_LAUNCH(k_first, T.n_tiles, 256, p->d_src, p->d_dyn, T.tiles, mode, p->d_stamp, 0);
    LAUNCH_CHECK();
    if (p->any_threshold) {
      if (T.n_mean) {
        PB(K_MEAN);
        if (p->fp32) k_mean_partial<float><<<T.n_chunks, 256, 0, st>>>(p->d_src, p->d_dyn, T.chunks, mode, p->d_stamp, p->d_meanpart);
        else k_mean_partial<double><<<T.n_chunks, 256, 0, st>>>(p->d_src, p->d_dyn, T.chunks, mode, p->d_stamp, p->d_meanpart);
        LAUNCH_CHECK();
        PB(K_MEAN);
        k_mean_final<<<ceil_div(T.n_mean, 128), 128, 0, st>>>(p->d_src, p->d_dyn, T.mean_list, T.n_mean, mode, p->d_meanpart);
        LAUNCH_CHECK();
      }
      Queues q = p->q;
      q.NVp = grad ? p->NVp_grad : 1;
      if (p->use_coop) {
        PB(K_SELECT);
        k_select<<<T.n_tiles, 256, 0, st>>>(p->d_src, p->d_dyn, T.tiles, mode, p->d_stamp, q);
        LAUNCH_CHECK();
        // the queue length is known only on the device: short queues are integrated by k_integrate (lanes share an
        // entry: lowest latency), long ones by k_integrate_pool (a lane per cell: full lanes); each kernel returns at
        // once when the queue is not its kind.  A queue cannot be longer than the first pass has pixels.
        // Which kernel the queue belongs to is a guess from the length the previous pass of this mode reported
        // (copied back asynchronously, never waited for); only while that is unknown are both launched.  Either
        // kernel integrates any queue correctly, so a stale guess costs time, not parity.
        const bool may_pool = p->pool_ok && p->first_evals[mode] >= p->pool_min;
        const int last = may_pool ? *(volatile int*)&p->h_qlast[mode] : -1;
        const bool run_lane = !may_pool || last < p->pool_min;     // unknown (-1) counts as short
        const bool run_pool = may_pool && (last < 0 || last >= p->pool_min);
        const int n_max = run_pool ? p->pool_min : INT_MAX;        // alone, k_integrate takes any length
        const int n_min = run_lane ? p->pool_min : 0;              // alone, k_integrate_pool takes any length
        if (may_pool) CU(cudaMemcpyAsync(&p->h_qlast[mode], p->q.count + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        if (run_lane && n_max > 0) {
          PB(grad ? K_INTEGRATE_G : K_INTEGRATE);
          PROFILE_LAUNCH(k_integrate, p->integrate_grid[grad ? 1 : 0], 128, p->d_src, p->d_dyn, mode, p->d_stamp, q, p->refine_lanes, n_max);
          LAUNCH_CHECK();
        }
        if (run_pool) {
          PB(grad ? K_POOL_G : K_POOL);
          PROFILE_LAUNCH(k_integrate_pool, p->pool_grid[grad ? 1 : 0], POOL_B, p->d_src, p->d_dyn, mode, p->d_stamp, q, n_min, p->pool_g2, grad ? p->pool_nv : 1, p->max_depth);
          LAUNCH_CHECK();
        }
      } else {
      PB(K_SELECT);
      k_select<<<T.n_tiles, 256, 0, st>>>(p->d_src, p->d_dyn, T.tiles, mode, p->d_stamp, q);
      LAUNCH_CHECK();
      const int grid = 148 * 8;
      for (int d = 1; d <= p->max_depth; ++d) {
        PB(K_REFINE);
        if (grad) k_refine<true><<<grid, 128, 0, st>>>(p->d_src, p->d_dyn, mode, d, q);
        else k_refine<false><<<grid, 128, 0, st>>>(p->d_src, p->d_dyn, mode, d, q);
        LAUNCH_CHECK();
      }
      for (int d = p->max_depth - 1; d >= 1; --d) {
        PB(K_REDUCE);
        k_reduce_level<<<grid, 128, 0, st>>>(p->d_src, d, q);
        LAUNCH_CHECK();
      }
      PB(K_SCATTER);
      k_scatter<<<grid, 128, 0, st>>>(p->d_src, q, p->d_stamp, grad);
      LAUNCH_CHECK();
      }
    }
    if (p->n_norm) {
      PB(K_NORM);
      k_normalize<<<p->n_norm, 256, 0, st>>>(p->d_src, p->norm_list, mode, p->d_stamp, grad);
      LAUNCH_CHECK();
    }
    if (p->n_amp) {
      PB(K_AMP);
      k_amp<<<p->n_amp, 256, 0, st>>>(p->d_src, p->d_dyn, p->amp_list, mode, p->d_stamp, grad);
      LAUNCH_CHECK();
    }
  }
  if (p->n_psf_list && p->any_aux_psf) {
    const int rc = psf_branch();
    if (rc) return rc;
  }
  if (fork && cudaStreamWaitEvent(st, p->ev_join, 0) != cudaSuccess) APB_FAIL("join of the PSF stream failed");
  if (p->n_point) {
    PB(K_POINT);
    k_point<<<p->n_point, 256, 0, st>>>(p->d_src, p->d_dyn, p->point_list, p->d_psfst, p->d_out, grad);
    LAUNCH_CHECK();
  }
  if (T.n_conv_tiles[grad]) {
    PB(K_CONV);
    k_conv<<<T.n_conv_tiles[grad], 256, T.conv_smem, st>>>(p->d_src, T.conv_jobs[grad], T.conv_tiles[grad], mode,
                                                           p->d_stamp, p->d_psfst, p->d_out);
    LAUNCH_CHECK();
  }
  if (p->n_fft_src) {
    PB(K_FFTROWS);
    k_fft_rows<<<F.n_rows, p->fft_nt_rows, p->fft_smem_rows, st>>>(p->d_src, p->d_fftdesc, p->d_twid, F.rows, mode, p->d_stamp,
                                                                    p->d_psfst, p->d_spec);
    LAUNCH_CHECK();
    PB(K_FFTCOLS);
    k_fft_cols<<<F.n_cols_img, p->fft_nt_cols, p->fft_smem_cols, st>>>(p->d_src, p->d_fftdesc, p->d_twid, F.jobs, F.cols_img, mode,
                                                            p->d_spec);
    LAUNCH_CHECK();
    PB(K_FFTINV);
    k_fft_rows_inv<<<F.n_rows_inv, p->fft_nt_rows, p->fft_smem_rows, st>>>(p->d_src, p->d_fftdesc, p->d_twid, F.rows_inv, p->d_spec,
                                                                p->d_out);
    LAUNCH_CHECK();
  }
  if (p->n_up) {
    // super-sampled sources: fine output window -> image pixels (image_object.py:331-376 ``reduce``)
    PB(K_UPSUM);
    k_reduce_up<<<dim3(p->n_up, grad ? p->NVp_grad : 1, 4), 256, 0, st>>>(p->d_src, p->up_list, mode, p->d_stamp, p->d_out, grad);
    LAUNCH_CHECK();
  }
  return 0;
}

static int assemble(apb_plan* p, int mode, double** model_out_dev, double** resid_out_dev, double* chi_out2,
                    int write_flag, cudaStream_t st) {
  PB(K_ASSEMBLE);
  k_assemble<<<p->n_img_tiles, 256, 0, st>>>(p->d_src, p->d_dyn, p->d_img, p->img_tiles, p->bin_ptr, p->bin_src, mode,
                                             p->d_stamp, p->d_out, model_out_dev, resid_out_dev,
                                             chi_out2 ? p->d_chipart : nullptr, p->d_done, chi_out2, write_flag,
                                             p->q.overflow);
  LAUNCH_CHECK();
  return 0;
}

static int begin_call(apb_plan* p, cudaStream_t st) {
  if (!p) APB_FAIL("plan is NULL");
  p->launches = 0;
  p->last_stream = st;
  return 0;
}

extern "C" int apb_sample(apb_plan_t* p, const double* x, int as_rep, double* const* model_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (begin_call(p, st)) return -1;
  if (!model_out) APB_FAIL("apb_sample: model_out is NULL");
  if (p->n_par > 0 && !x) APB_FAIL("apb_sample: x is NULL");
  CU(cudaMemcpyAsync(p->d_userptr, model_out, sizeof(double*) * p->n_img, cudaMemcpyHostToDevice, st));
  int rc = sample_pass(p, x, as_rep, 0, 0, st);
  if (rc) return rc;
  rc = assemble(p, 0, p->d_userptr, nullptr, nullptr, 0, st);
  p->stats.launches = p->launches;
  return rc;
}

extern "C" int apb_jacobian(apb_plan_t* p, const double* x, int as_rep, double* const* jac_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (begin_call(p, st)) return -1;
  if (!jac_out) APB_FAIL("apb_jacobian: jac_out is NULL");
  for (int ii = 0; ii < p->n_img; ++ii)
    if (!(p->h_img[ii].flags & APB_IMG_AUX))
      CU(cudaMemsetAsync(jac_out[ii], 0, sizeof(double) * (size_t)p->h_img[ii].H * p->h_img[ii].W * p->n_par, st));
  if (p->n_par == 0) return 0;
  CU(cudaMemcpyAsync(p->d_userptr, jac_out, sizeof(double*) * p->n_img, cudaMemcpyHostToDevice, st));
  int rc = sample_pass(p, x, as_rep, 1, 1, st);
  if (rc) return rc;
  PB(K_JAC);
  k_jac_dense<<<p->n_img_tiles, 256, 0, st>>>(p->d_src, p->d_img, p->img_tiles, p->bin_ptr, p->bin_src, p->d_stamp,
                                              p->d_out, p->d_skyJ, p->d_userptr, p->n_par);
  LAUNCH_CHECK();
  p->stats.launches = p->launches;
  return 0;
}

static int chi2_core(apb_plan* p, const double* x_rep, double* out2, cudaStream_t st) {
  for (int ii = 0; ii < p->n_img; ++ii)
    if (!p->h_img[ii].data && !(p->h_img[ii].flags & APB_IMG_AUX)) APB_FAIL("apb_chi2: image without data");
  int rc = sample_pass(p, x_rep, 1, 0, 0, st);
  if (rc) return rc;
  return assemble(p, 0, nullptr, nullptr, out2, 1, st);
}

extern "C" int apb_chi2(apb_plan_t* p, const double* x_rep, double* out2, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (begin_call(p, st)) return -1;
  int rc = chi2_core(p, x_rep, out2, st);
  p->stats.launches = p->launches;
  return rc;
}

extern "C" int apb_normal_eq(apb_plan_t* p, const double* x, int as_rep, double* JtWJ, double* JtWr, double* chi2,
                             void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (begin_call(p, st)) return -1;
  for (int ii = 0; ii < p->n_img; ++ii)
    if (!p->h_img[ii].data && !(p->h_img[ii].flags & APB_IMG_AUX)) APB_FAIL("apb_normal_eq: image without data");
  const int P = p->n_par;
  int rc;
  if (!p->all_same_geo) {
    // forward model in group geometry for the residual, then derivatives in own-window geometry
    if ((rc = sample_pass(p, x, as_rep, 0, 0, st))) return rc;
    if ((rc = assemble(p, 0, nullptr, p->d_resid, chi2, 1, st))) return rc;
    if ((rc = sample_pass(p, x, as_rep, 1, 1, st))) return rc;
  } else {
    if ((rc = sample_pass(p, x, as_rep, 1, 1, st))) return rc;
    if ((rc = assemble(p, 1, nullptr, p->d_resid, chi2, 1, st))) return rc;
  }
  if (!JtWJ && !p->sparse_ok) APB_FAIL("apb_normal_eq: JtWJ may only be NULL when the plan has a block-sparse form");
  if (JtWJ) CU(cudaMemsetAsync(JtWJ, 0, sizeof(double) * (size_t)P * P, st));
  CU(cudaMemsetAsync(JtWr, 0, sizeof(double) * (size_t)P, st));
  if (p->sparse_ok) CU(cudaMemsetAsync(p->d_bvals, 0, sizeof(double) * ((size_t)p->n_bvals + (size_t)P), st));
  if (p->n_items) {
    PB(K_BLOCKS);
    k_blocks<<<p->n_items, 256, 0, st>>>(p->d_src, p->d_img, p->d_items, p->d_stamp, p->d_out, p->d_skyJ, p->d_resid,
                                         -1.0, 0, p->d_part);
    LAUNCH_CHECK();
    PB(K_BLOCKFIN);
    k_block_final<<<dim3(p->n_blocks, BLK_VALS / 8), 256, 0, st>>>(p->d_blocks, p->d_part, 0, p->d_btot);
    LAUNCH_CHECK();
    const int nd = JtWJ ? p->n_gdst[0] : p->n_gdst_noH[0];
    if (nd) {
      PB(K_BLOCKFIN);
      k_block_gather<<<ceil_div(nd, 256), 256, 0, st>>>(p->d_gdst[0], nd, p->d_gsrc[0], p->d_btot, JtWr, -1.0,
                                                        p->sparse_ok ? p->d_bvals : nullptr, p->d_diagH, JtWJ);
      LAUNCH_CHECK();
    }
  }
  p->stats.launches = p->launches;
  return 0;
}

static int geodesic_core(apb_plan* p, apb_plan* pj, const double* xdh, const double* h, double d, double* rpp, cudaStream_t st);
extern "C" int apb_geodesic(apb_plan_t* p, const double* xdh, const double* h, double d, double* rpp, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (begin_call(p, st)) return -1;
  int rc = geodesic_core(p, p, xdh, h, d, rpp, st);
  p->stats.launches = p->launches;
  return rc;
}

// pj: the plan whose last apb_normal_eq left the stamp Jacobian (derivative planes) and the residual r at x; p itself,
// or -- for a speculative lambda-trial running beside the main one -- another plan of the same scene (only read here:
// a forward pass never writes derivative planes).
static int geodesic_core(apb_plan* p, apb_plan* pj, const double* xdh, const double* h, double d, double* rpp, cudaStream_t st) {
  const int P = p->n_par;
  int rc;
  // rh = W (Y(x + d h) - Y): forward pass only touches plane 0, the cached derivative planes stay valid
  if ((rc = sample_pass(p, xdh, 1, 0, 0, st))) return rc;
  if ((rc = assemble(p, 0, nullptr, p->d_resid2, nullptr, 0, st))) return rc;
  // everything that reads the stamp Jacobian runs on the DONOR's tables (its sources may be cut differently from
  // this plan's: a model whose window exceeds image_chunksize is one piece here and one piece per chunk there)
  PB(K_GEOV);
  k_geo_v<<<pj->n_img_tiles, 256, 0, st>>>(pj->d_src, pj->d_img, pj->img_tiles, pj->bin_ptr, pj->bin_src, pj->d_stamp, pj->d_out,
                                           pj->d_skyJ, h, d, pj->d_resid, p->d_resid2);
  LAUNCH_CHECK();
  CU(cudaMemsetAsync(rpp, 0, sizeof(double) * (size_t)P, st));
  if (pj->n_vitems) {
    // partial sums go to this plan's scratch when it is large enough (two trials may share one donor)
    double* part = p->part_cap >= (size_t)pj->n_vitems ? p->d_part : pj->d_part;
    PB(K_BLOCKS);
    k_blocks<<<pj->n_vitems, 256, 0, st>>>(pj->d_src, pj->d_img, pj->d_vitems, pj->d_stamp, pj->d_out, pj->d_skyJ, p->d_resid2,
                                           1.0, 1, part);
    LAUNCH_CHECK();
    PB(K_BLOCKFIN);
    double* btot = p->btot_cap >= (size_t)pj->n_vblocks ? p->d_btot : pj->d_btot;
    k_block_final<<<dim3(pj->n_vblocks, BLK_VALS / 8), 256, 0, st>>>(pj->d_vblocks, part, 1, btot);
    LAUNCH_CHECK();
    PB(K_BLOCKFIN);
    k_block_gather<<<ceil_div(pj->n_gdst[1], 256), 256, 0, st>>>(pj->d_gdst[1], pj->n_gdst[1], pj->d_gsrc[1], btot, rpp, 1.0,
                                                                 nullptr, nullptr, nullptr);
    LAUNCH_CHECK();
  }
  return 0;
}

static int lm_solve_launch(const double* H, const double* g, double L, int P, double* h, int* info, const LmEpi& epi,
                           cudaStream_t st) {
  const size_t smem = sizeof(double) * (size_t)P * (P + 1);
  if (smem > 200 * 1024) APB_FAIL("apb_lm_solve: P too large for the single-CTA solver (max 159); use a library solver");
  if (smem > 48 * 1024) {   // (plans opt the kernel in at creation; a bare apb_lm_solve call may come first)
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, (const void*)k_lm_solve_small));
    CU(cudaFuncSetAttribute(k_lm_solve_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - (int)fa.sharedSizeBytes));
  }
  k_lm_solve_small<<<1, 256, smem, st>>>(H, g, L, P, h, info, epi);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) APB_FAIL(std::string("apb_lm_solve launch: ") + cudaGetErrorString(e));
  g_launches++;
  return 0;
}

// One lambda-trial of LM.step (fit/lm.py:274-293) without leaving the device:
//   h = solve(L, g);  rpp = geodesic(x + d h, h);  a = -solve(L, rpp)/2 (0 if L <= 1e-4);
//   ha = h + acceleration a;  rec = [chi2(x + ha), status flag, |a|, |h|]
// H, g: the normal equations of the last apb_normal_eq.  h_out, ha_out: device, P doubles.
// rec: device, 4 doubles -- the single record the host reads back per trial.  P <= 159.
// With acceleration == 0 (the reference's default) x + ha = x + h does not depend on the geodesic
// term, so when a second, forward-only plan of the same scene is supplied the chi^2 pass runs on
// it concurrently (own stream, own workspace) with the geodesic pass: both are chains of small
// latency-bound launches, and side by side they take about the time of one.
extern "C" int apb_lm_trial_spec(apb_plan_t* p, apb_plan_t* p2, apb_plan_t* donor, const double* H, const double* g,
                                 double L, const double* x_rep, double d, double acceleration, double* h_out,
                                 double* ha_out, double* rec, void* stream);
extern "C" int apb_lm_trial(apb_plan_t* p, apb_plan_t* p2, const double* H, const double* g, double L,
                            const double* x_rep, double d, double acceleration, double* h_out, double* ha_out,
                            double* rec, void* stream) {
  return apb_lm_trial_spec(p, p2, p, H, g, L, x_rep, d, acceleration, h_out, ha_out, rec, stream);
}

extern "C" int apb_lm_trial_spec(apb_plan_t* p, apb_plan_t* p2, apb_plan_t* donor, const double* H, const double* g,
                                 double L, const double* x_rep, double d, double acceleration, double* h_out,
                                 double* ha_out, double* rec, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (begin_call(p, st)) return -1;
  if (!donor) donor = p;
  if (donor->n_par != p->n_par || donor->n_img != p->n_img)
    APB_FAIL("apb_lm_trial_spec: the donor plan does not describe the same fit (parameters, images)");
  const int P = p->n_par;
  if (P <= 0) APB_FAIL("apb_lm_trial: no parameters");
  const bool overlap = p2 != nullptr && acceleration == 0.0;
  if (overlap && (p2->n_par != P || p2->n_img != p->n_img || p2->n_src != p->n_src))
    APB_FAIL("apb_lm_trial: the second plan does not describe the same scene");
  int rc;
  LmEpi e1{1, x_rep, nullptr, d, 0.0, p->d_xtmp, overlap ? p->d_xtmp2 : nullptr, nullptr, nullptr};
  if ((rc = lm_solve_launch(H, g, L, P, h_out, nullptr, e1, st))) return rc;
  if (overlap) {
    CU(cudaEventRecord(p->ev_trial_fork, st));
    CU(cudaStreamWaitEvent(p->trial_stream, p->ev_trial_fork, 0));
    p2->launches = 0;
    p2->last_stream = p->trial_stream;
    if ((rc = chi2_core(p2, p->d_xtmp2, p->d_rec2, p->trial_stream))) return rc;
    CU(cudaEventRecord(p->ev_trial_join, p->trial_stream));
  }
  if ((rc = geodesic_core(p, donor, p->d_xtmp, h_out, d, p->d_rpp, st))) return rc;
  LmEpi e2{2, x_rep, h_out, d, acceleration, overlap ? p->d_atmp : p->d_xtmp2, ha_out, rec, nullptr};
  if ((rc = lm_solve_launch(H, p->d_rpp, L, P, p->d_atmp2, nullptr, e2, st))) return rc;
  if (overlap) {
    CU(cudaStreamWaitEvent(st, p->ev_trial_join, 0));
    k_trial_join<<<1, 1, 0, st>>>(p->d_rec2, p->q.overflow, p2->q.overflow, donor != p ? donor->q.overflow : nullptr, rec);
    p->launches += p2->launches + 1;
    g_launches++;
  } else {
    if ((rc = chi2_core(p, p->d_xtmp2, rec, st))) return rc;
    if (donor != p) {
      k_trial_join<<<1, 1, 0, st>>>(rec, p->q.overflow, nullptr, donor->q.overflow, rec);
      p->launches++;
      g_launches++;
    }
  }
  p->stats.launches = p->launches + 2;
  return 0;
}

// The same trial in two halves, for fits whose pixels are spread over several GPUs (acceleration == 0):
//   begin: h = solve(L, g); local rpp and local chi2(x + h) -> buf = { rpp[P], chi2, #non-finite, #overflow }
//   (caller: sum all-reduce of buf over the ranks)
//   end:   a = -solve(L, rpp)/2, rec = { chi2, status flag, |a|, |h| }, ha = h
// On one GPU the two are simply called back to back.
extern "C" int apb_lm_trial_begin(apb_plan_t* p, apb_plan_t* p2, const double* H, const double* g, double L,
                                  const double* x_rep, double d, double* h_out, double* buf, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (begin_call(p, st)) return -1;
  const int P = p->n_par;
  if (P <= 0) APB_FAIL("apb_lm_trial_begin: no parameters");
  if (p2 && (p2->n_par != P || p2->n_img != p->n_img || p2->n_src != p->n_src))
    APB_FAIL("apb_lm_trial_begin: the second plan does not describe the same scene");
  int rc;
  LmEpi e1{1, x_rep, nullptr, d, 0.0, p->d_xtmp, p->d_xtmp2, nullptr, nullptr};
  if ((rc = lm_solve_launch(H, g, L, P, h_out, nullptr, e1, st))) return rc;
  if (p2) {
    CU(cudaEventRecord(p->ev_trial_fork, st));
    CU(cudaStreamWaitEvent(p->trial_stream, p->ev_trial_fork, 0));
    p2->launches = 0;
    p2->last_stream = p->trial_stream;
    if ((rc = chi2_core(p2, p->d_xtmp2, p->d_rec2, p->trial_stream))) return rc;
    CU(cudaEventRecord(p->ev_trial_join, p->trial_stream));
  }
  if ((rc = geodesic_core(p, p, p->d_xtmp, h_out, d, buf, st))) return rc;
  if (p2) {
    CU(cudaStreamWaitEvent(st, p->ev_trial_join, 0));
    p->launches += p2->launches;
  } else {
    if ((rc = chi2_core(p, p->d_xtmp2, p->d_rec2, st))) return rc;
  }
  k_trial_tail<<<1, 1, 0, st>>>(p->d_rec2, p->q.overflow, p2 ? p2->q.overflow : nullptr, buf + P);
  g_launches++;
  p->stats.launches = p->launches + 2;
  return 0;
}

extern "C" int apb_lm_trial_end(apb_plan_t* p, const double* H, double L, const double* x_rep, const double* h,
                                const double* buf, double* ha_out, double* rec, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!p) APB_FAIL("plan is NULL");
  const int P = p->n_par;
  LmEpi e2{2, x_rep, h, 0.0, 0.0, p->d_atmp, ha_out, rec, buf + P};
  return lm_solve_launch(H, buf, L, P, p->d_atmp2, nullptr, e2, st);
}

extern "C" int apb_lm_solve(const double* H, const double* g, double L, int P, double* h, int* info, void* stream) {
  if (P <= 0) return 0;
  LmEpi e0{0, nullptr, nullptr, 0.0, 0.0, nullptr, nullptr, nullptr, nullptr};
  return lm_solve_launch(H, g, L, P, h, info, e0, (cudaStream_t)stream);
}

// Dense damped system beyond the single-CTA solver (apb_chol.cuh): factor once per (H, L), solve per right-hand side.
//   W: device, P*P + 2 doubles (the factor; the tail holds the grid-barrier counter).  info: device int, 0 ok.
extern "C" int apb_chol_factor(const double* H, double L, int P, double* W, int* info, void* stream) {
  if (P <= 0) return 0;
  if (!H || !W || !info) APB_FAIL("apb_chol_factor: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, n_sm = 0, coop = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop || n_sm <= 0) APB_FAIL("apb_chol_factor: the device cannot launch cooperative kernels");
  unsigned int* bar = (unsigned int*)(W + (size_t)P * P);
  CU(cudaMemsetAsync(bar, 0, 2 * sizeof(double), st));
  const int nb = (P + CH_NB - 1) / CH_NB;
  // (no more CTAs than the widest phase has tiles: the barriers cost by the number of arrivals)
  int grid = std::max(1, std::min(n_sm, std::max(nb * (nb - 1) / 2, (int)std::min<long long>((long long)P * P / CH_NT + 1, n_sm))));
  // (APB_CHOL_GRID / APB_CHOL_RELAXED: tuning knobs for experiments)
  if (const char* e = getenv("APB_CHOL_GRID")) grid = std::max(1, std::min(n_sm, atoi(e)));
  int relaxed = 0;
  if (const char* e = getenv("APB_CHOL_RELAXED")) relaxed = atoi(e);
  CholArgs A{H, L, P, W, info, bar, relaxed};
  void* args[] = {&A};
  CU(cudaLaunchCooperativeKernel((void*)k_chol_factor, dim3(grid), dim3(CH_NT), args, 0, st));
  g_launches++;
  return 0;
}

// x = A^-1 rhs with the factor of apb_chol_factor.  rhs, x: device, P (may alias).
extern "C" int apb_chol_solve(const double* W, const double* rhs, int P, double* x, void* stream) {
  if (P <= 0) return 0;
  if (!W || !rhs || !x) APB_FAIL("apb_chol_solve: NULL argument");
  k_chol_solve<<<1, CH_SOLVE_NT, 0, (cudaStream_t)stream>>>(W, rhs, P, x);
  CU(cudaGetLastError());
  g_launches++;
  return 0;
}

// Damped solve of the system of the last apb_normal_eq by block-sparse PCG (apb_solve.cuh).
//   g: device, n_par;  h: device, n_par (out);  info: device, 2 doubles {iterations, |r|/|b|}.
// Returns 1 (not an error) when the plan cannot use it (parameters shared between sources).
extern "C" int apb_lm_solve_sparse(apb_plan_t* p, const double* g, double L, const double* x0, double* h, double* info,
                                   double tol, int max_iter, void* stream) {
  if (!p) APB_FAIL("plan is NULL");
  if (!p->sparse_ok) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t P = (size_t)p->n_par;
  PcgArgs A{p->d_prows, p->n_prows, p->d_ppasses, p->n_ppass, p->d_packed, p->d_pslots, p->d_pitems, p->n_pitems, p->d_multi_rows, p->n_multi,
            p->d_own_slot, p->d_own_off, p->d_bvals, p->d_diagH, p->d_pfac, g, x0, h, p->d_pvec, p->d_pvec + 2 * P,
            p->d_pvec + 4 * P, p->d_pvec + 5 * P, p->d_qpart, p->d_pcg_part, p->d_pcg_bar,
            info, p->n_par, max_iter > 0 ? max_iter : 2000, L, tol > 0.0 ? tol : 1e-14};
  void* args[] = {&A};
  CU(cudaMemsetAsync(p->d_pcg_bar, 0, sizeof(unsigned int), st));
  p->pbegin(K_PCG, st);
  k_pcg_pack<<<296, 256, 0, st>>>(p->d_bvals, p->d_pack_src, p->d_packed, p->n_ppacked);
  CU(cudaLaunchCooperativeKernel((void*)k_pcg, dim3(p->pcg_grid), dim3(PCG_NT), args, 0, st));
  p->pend(st);
  g_launches++;
  return 0;
}

extern "C" long long apb_plan_block_doubles(apb_plan_t* p) {
  if (!p || !p->sparse_ok) return 0;
  return p->n_bvals + (long long)p->n_par;
}

extern "C" int apb_plan_bind_blocks(apb_plan_t* p, double* buf) {
  if (!p) APB_FAIL("plan is NULL");
  if (!p->sparse_ok) APB_FAIL("apb_plan_bind_blocks: the plan has no block-sparse form");
  p->d_bvals = buf ? buf : p->d_bvals_own;
  p->d_diagH = p->d_bvals + p->n_bvals;
  return 0;
}

extern "C" int apb_plan_stats(apb_plan_t* p, apb_stats_t* out) {
  if (!p || !out) APB_FAIL("apb_plan_stats: NULL");
  CU(cudaStreamSynchronize(p->last_stream));
  int cnt[APB_MAX_DEPTH + 2], ovf = 0;
  CU(cudaMemcpy(cnt, p->q.count, sizeof(cnt), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&ovf, p->q.overflow, sizeof(int), cudaMemcpyDeviceToHost));
  memset(out, 0, sizeof(*out));
  out->first_pass_evals = p->first_evals[0];
  // the fused integration kernels only COUNT the entries of depth >= 2 (32-bit atomics): read them as unsigned, a
  // 16k x 16k mosaic passes 2^31
  for (int d = 1; d <= APB_MAX_DEPTH; ++d) out->queued[d] = (d >= 2 && p->use_coop) ? (long long)(unsigned int)cnt[d] : cnt[d];
  out->launches = p->stats.launches;
  out->overflow = ovf;
  unsigned long long cum[2][APB_MAX_DEPTH + 2];
  int last = 0;
  CU(cudaMemcpy(cum, p->q.cum, sizeof(cum), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&last, p->q.last_kind, sizeof(int), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 2; ++k) {
    out->cum_passes[k] = p->cum_passes[k];
    out->cum_first_pass_evals[k] = p->cum_first[k];
    for (int d = 1; d <= APB_MAX_DEPTH; ++d)
      out->cum_queued[k][d] = (long long)cum[k][d] + (last == k ? (long long)(unsigned int)cnt[d] : 0);
  }
  return 0;
}

// ---- measurement helpers ---------------------------------------------------------------------
extern "C" long long apb_launch_count(void) { return g_launches; }

extern "C" int apb_profile(apb_plan_t* p, int enable) {
  if (!p) APB_FAIL("apb_profile: NULL plan");
  p->profiling = enable != 0;
  return 0;
}

extern "C" int apb_profile_read(apb_plan_t* p, apb_kernel_time_t* out, int max_out, int* n_out, int reset) {
  if (!p || !out || !n_out) APB_FAIL("apb_profile_read: NULL");
  CU(cudaStreamSynchronize(p->last_stream));
  for (auto& r : p->prof) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { p->prof_ms[r.id] += ms; p->prof_n[r.id]++; }
    p->ev_pool.push_back(r.a);
    p->ev_pool.push_back(r.b);
  }
  p->prof.clear();
  int n = 0;
  for (int k = 0; k < K_COUNT && n < max_out; ++k) {
    if (!p->prof_n[k]) continue;
    memset(&out[n], 0, sizeof(out[n]));
    strncpy(out[n].name, kKNames[k], sizeof(out[n].name) - 1);
    out[n].launches = p->prof_n[k];
    out[n].total_ms = p->prof_ms[k];
    ++n;
  }
  *n_out = n;
  if (reset) for (int k = 0; k < K_COUNT; ++k) { p->prof_ms[k] = 0; p->prof_n[k] = 0; }
  return 0;
}

// DFMA-stream and copy microbenchmarks: the FP64 ceiling is not in MEASURED_PEAKS.json
__global__ void __launch_bounds__(256) k_bench_dfma(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void __launch_bounds__(256) k_bench_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

extern "C" int apb_bench_peaks(double* dfma_tflops, double* copy_gbs) {
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  const int blocks = 148 * 8, iters = 1 << 15;
  double* buf = nullptr;
  CU(cudaMalloc((void**)&buf, sizeof(double) * blocks * 256));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(e0));
    k_bench_dfma<<<blocks, 256>>>(buf, iters);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms; CU(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  if (dfma_tflops) *dfma_tflops = 2.0 * 8.0 * iters * (double)blocks * 256 / (best * 1e-3) / 1e12;
  CU(cudaFree(buf));
  const size_t n = (size_t)1 << 26;   // 2 x 1 GiB
  double2 *a = nullptr, *b = nullptr;
  CU(cudaMalloc((void**)&a, n * sizeof(double2)));
  CU(cudaMalloc((void**)&b, n * sizeof(double2)));
  CU(cudaMemset(a, 0, n * sizeof(double2)));
  best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(e0));
    k_bench_copy<<<148 * 16, 256>>>(a, b, n);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms; CU(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  if (copy_gbs) *copy_gbs = 2.0 * n * sizeof(double2) / (best * 1e-3) / 1e9;
  CU(cudaFree(a));
  CU(cudaFree(b));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

// ---- one-node all-reduce over NVLink peer memory (apb_comm.cuh) ------------------------------------------------
static size_t comm_bytes(size_t slot_doubles) { return 2 * slot_doubles * sizeof(double) + APB_COMM_MAX_RANKS * sizeof(unsigned long long); }

extern "C" int apb_comm_alloc(size_t max_doubles, void** local_out, void* handle_out) {
  if (!local_out || !handle_out || max_doubles == 0) APB_FAIL("apb_comm_alloc: bad arguments");
  void* buf = nullptr;
  CU(cudaMalloc(&buf, comm_bytes(max_doubles)));
  CU(cudaMemset(buf, 0, comm_bytes(max_doubles)));
  CU(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, buf));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  *local_out = buf;
  return 0;
}

extern "C" int apb_comm_create(int rank, int world, void* local, const void* handles, size_t max_doubles, apb_comm_t** out) {
  if (!out || !local || !handles) APB_FAIL("apb_comm_create: NULL argument");
  if (world < 1 || world > APB_COMM_MAX_RANKS || rank < 0 || rank >= world) APB_FAIL("apb_comm_create: bad rank / world");
  apb_comm* c = new apb_comm();
  c->rank = rank; c->world = world; c->slot_doubles = max_doubles; c->local = local;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { c->peer[r] = local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * (size_t)r, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(&c->peer[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      g_err = std::string("apb_comm_create: cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e);
      for (int q = 0; q < r; ++q) if (q != rank && c->peer[q]) cudaIpcCloseMemHandle(c->peer[q]);
      delete c;
      return -2;
    }
  }
  CU(cudaMalloc((void**)&c->done, sizeof(unsigned int)));
  CU(cudaMemset(c->done, 0, sizeof(unsigned int)));
  int dev = 0, sms = 148;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  c->grid_max = sms;       // one CTA per SM: always co-resident
  *out = c;
  return 0;
}

extern "C" int apb_allreduce(apb_comm_t* c, double* buf, size_t n, void* stream) {
  if (!c || !buf) APB_FAIL("apb_allreduce: NULL argument");
  if (n == 0 || c->world == 1) return 0;
  if (n > c->slot_doubles) APB_FAIL("apb_allreduce: buffer longer than the exchange slots of this communicator");
  CommArgs A;
  memset(&A, 0, sizeof(A));
  c->seq++;
  const size_t slot_off = (c->seq & 1) ? c->slot_doubles : 0;
  for (int r = 0; r < c->world; ++r) {
    A.slot[r] = (double*)c->peer[r] + slot_off;
    A.flags[r] = (unsigned long long*)((double*)c->peer[r] + 2 * c->slot_doubles);
  }
  A.rank = c->rank; A.world = c->world; A.seq = c->seq; A.done = c->done;
  // two elements per thread: the sum is bound by the latency of the peer loads, so spread them wide
  const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)c->grid_max, (n + 511) / 512));
  k_allreduce_peer<<<grid, 256, 0, (cudaStream_t)stream>>>(A, buf, n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) APB_FAIL(std::string("apb_allreduce launch: ") + cudaGetErrorString(e));
  g_launches++;
  return 0;
}

extern "C" int apb_comm_destroy(apb_comm_t* c) {
  if (!c) return 0;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  if (c->local) cudaFree(c->local);
  if (c->done) cudaFree(c->done);
  delete c;
  return 0;
}

#ifdef PCG_TIMING
extern "C" int apb_debug_pcg_clk(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_pcg_clk, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -1;
}
#endif
