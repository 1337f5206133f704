// boomgpu_api.cu -- the C ABI of include/boomgpu.h over the kernels in kernels.cuh.
// One context = one CUDA device, one stream; no CPU fallback anywhere.
#include "../../include/boomgpu.h"

#include <algorithm>
#include <numeric>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "fused_tma.cuh"

using namespace boomgpu;

namespace {

thread_local std::string g_create_error;

struct TimedLaunch { cudaEvent_t a, b; int cls; };

}  // namespace

struct boomgpu_ctx {
  int device = 0;
  int sms = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string error;

  // data
  int model = -1;  // kLogit / kPoisson / kStudentT (plain regression data: y only)
  int upload_kind = -1;  // between boomgpu_upload_begin and _end: the model the chunks belong to
  int64_t n = 0;
  int p = 0;
  int64_t ldx = 0;
  const double *X = nullptr;
  const double *y = nullptr, *ntrials = nullptr, *exposure = nullptr;
  const int64_t *yi = nullptr;
  std::vector<void *> owned;  // device allocations made for uploaded data
  uint64_t row_offset = 0;
  // what the TMA-fed kernels read: X itself, or -- for adopted rows whose pointer / leading dimension TMA cannot describe
  // (odd ldx, base not 16-byte aligned) and p > 64 -- a padded device copy made once per adoption
  const double *Xt = nullptr;
  int64_t ldxt = 0;
  double *Xt_owned = nullptr;

  // X_gamma: the columns last named by boomgpu_select_columns, as an n x sel_ld matrix owned by the context
  double *Xsel = nullptr; int64_t sel_cap = 0; int sel_ld = 0;
  std::vector<int> sel_cols;
  const double *sel_src = nullptr;   // the X it was gathered from (a new upload / adoption invalidates it)
  int *sel_cols_dev = nullptr; int sel_cols_cap = 0;

  // active-set statistics: [G p x ka8 | diag p | xty p | 4 scalars] on the device and in pinned host memory
  double *act_dev = nullptr; int64_t act_cap = 0;
  double *act_pin = nullptr; int64_t act_pin_cap = 0;
  double *col_buf = nullptr; int64_t col_cap = 0;   // w o x_j for boomgpu_weighted_column
  int stats_mode = 0;                                 // 1 while an active-set step runs its imputer pass
  bool latents_valid = false;                         // w_buf / s_buf hold the latents of the last two-pass step

  // mixtures
  LogitMixture mix{};        // host copy
  LogitHot hot{};
  LogitHot ext_hot{};   // Poisson table entry nu = 1
  LogitMixtureDev *mix_dev = nullptr;
  bool have_mix = false;
  PoissonTable tab{};
  bool have_tab = false;
  std::vector<void *> tab_owned;
  // host copy of the table as last installed: re-stating an unchanged table (samplers do, every draw) uploads nothing
  std::vector<int64_t> tab_nu_h; std::vector<int32_t> tab_off_h; std::vector<double> tab_w_h, tab_mu_h, tab_sig_h;
  int64_t tab_cut_h = -1;

  // workspaces
  double *beta_dev = nullptr; int beta_cap = 0;   // [p + 2 doubles | p ints: indices of beta's non-zeros (gather pass)]
  double *beta_pin = nullptr;
  bool xty_only = false;                           // this step wants X'z alone (probit: X'WX is constant)
  int syrk_diag = 0;                               // 0 strip form for whole diagonal regions, 1 unit form everywhere
  int syrk_filter = 0;                             // profiling aid: time the off-diagonal / diagonal regions of the SYRK alone
  int syrk_order = 1;                              // 1 off-diagonal regions first, diagonal last (short CTAs fill the tail); 0 k-slice major;
                                                   // 2 k-slice major over uniform work items (diagonal regions in pairs)
  SyrkItem *syrk_items = nullptr; int syrk_items_nblk = -1, syrk_nitems = 0;   // order 2: the work items of one k-slice
  int syrk_waves = 30;                             // CTAs per SM the split-K aims for
  int syrk_rdiag = 1;                              // 1: the ragged last diagonal region runs in syrk_rdiag_kernel (strip form over its atom columns)
  int syrk_cluster = 0;                            // experiment: launch the SYRK with thread-block clusters of this many CTAs
  int gather = 0;                                  // option: 0 auto (sparse beta -> gather pass), 1 never, 2 whenever beta has a zero
  double *suf_dev = nullptr; int64_t suf_cap = 0;
  double *suf_pin = nullptr; int64_t suf_pin_cap = 0;
  double *partials = nullptr; int64_t partials_cap = 0;
  double *scal_partials = nullptr; int64_t scal_cap = 0;
  double *w_buf = nullptr, *s_buf = nullptr; int64_t ws_cap = 0;
  double *resid_buf = nullptr; int64_t resid_cap = 0;   // Student-t sibling: y - X beta of the last boomgpu_student_loglike(beta != NULL)
  bool resid_valid = false;
  int *err_dev = nullptr;
  int *err_pin = nullptr;
  unsigned int *tail_counter = nullptr;   // single-launch small-p step: CTAs that have written their partial
  int single_launch = 1;                  // option: 0 = separate reduction kernel (two launches)

  // TMA descriptor of X for the single-pass kernel (re-encoded when the data or the tile shape change)
  struct XMap {
    CUtensorMap map;
    const double *X = nullptr; int64_t n = -1, ldx = -1; int p = -1, box_cols = -1, box_rows = -1, promo = -1;
  } xmap_small, xmap_syrk, xmap_sel;
  int tma_promotion = 3;   // option: L2 promotion of the SYRK / panel tensor maps: 0 none, 1 64 B, 2 128 B, 3 256 B

  bool host_out_written = false;   // the last step's reduction wrote the statistics straight into suf_pin (zero copy)

  // NCCL communicator (opaque ncclComm_t), null = single GPU
  void *comm = nullptr;
  int comm_ranks = 1;

  // options / instrumentation
  int path = 0;
  int small_variant = 0;  // 0 auto (TMA kernel when X qualifies), 1 force the cp.async kernel
  bool timing = false;
  int64_t launches = 0;
  std::vector<TimedLaunch> timed;
  std::vector<cudaEvent_t> event_pool;
  double ms_acc[BOOMGPU_NUM_KERNEL_CLASSES] = {0, 0, 0, 0, 0};
  int64_t n_acc[BOOMGPU_NUM_KERNEL_CLASSES] = {0, 0, 0, 0, 0};
};

namespace {

int fail(boomgpu_ctx *c, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->error = buf; else g_create_error = buf;
  return code;
}

#define CU(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return fail(ctx, BOOMGPU_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <class T>
int ensure(boomgpu_ctx *ctx, T **ptr, int64_t *cap, int64_t need) {
  if (*cap >= need && *ptr) return 0;
  if (*ptr) { CU(cudaFree(*ptr)); *ptr = nullptr; *cap = 0; }
  CU(cudaMalloc((void **)ptr, sizeof(T) * (size_t)need));
  *cap = need;
  return 0;
}

void free_data(boomgpu_ctx *ctx) {
  for (void *q : ctx->owned) cudaFree(q);
  ctx->owned.clear();
  ctx->X = ctx->y = ctx->ntrials = ctx->exposure = nullptr;
  ctx->yi = nullptr;
  if (ctx->Xt_owned) { cudaFree(ctx->Xt_owned); ctx->Xt_owned = nullptr; }
  ctx->Xt = nullptr; ctx->ldxt = 0;
  ctx->n = 0; ctx->p = 0; ctx->model = -1; ctx->upload_kind = -1;
  ctx->latents_valid = false;
  ctx->resid_valid = false;
}

// ---- launch bookkeeping -------------------------------------------------------------------
struct LaunchScope {
  boomgpu_ctx *ctx; int cls; cudaEvent_t a = nullptr, b = nullptr;
  LaunchScope(boomgpu_ctx *c, int cls_) : ctx(c), cls(cls_) {
    ++ctx->launches;
    if (ctx->timing) {
      auto get = [&]() {
        cudaEvent_t e;
        if (!ctx->event_pool.empty()) { e = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
      };
      a = get(); b = get();
      cudaEventRecord(a, ctx->stream);
    }
  }
  ~LaunchScope() {
    if (ctx->timing) { cudaEventRecord(b, ctx->stream); ctx->timed.push_back({a, b, cls}); }
  }
};

void drain_timings(boomgpu_ctx *ctx) {
  for (auto &t : ctx->timed) {
    float ms = 0;
    cudaEventSynchronize(t.b);
    if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) { ctx->ms_acc[t.cls] += ms; ctx->n_acc[t.cls] += 1; }
    ctx->event_pool.push_back(t.a);
    ctx->event_pool.push_back(t.b);
  }
  ctx->timed.clear();
}

// ---- NCCL, bound at run time ------------------------------------------------------------------
// Minimal mirror of nccl.h (stable ABI since NCCL 2.0): the 128-byte unique id, result code 0 = success,
// ncclDouble = 8 (ncclFloat64), ncclSum = 0.
struct NcclUniqueId { char internal[128]; };
struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(void **, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string why;
};
const NcclApi &nccl() {
  static NcclApi api = [] {
    NcclApi a;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { a.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return a; }
    a.GetUniqueId = (int (*)(NcclUniqueId *))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(void **, int, NcclUniqueId, int))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
    a.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclAllReduce");
    a.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString;
    if (!a.ok) a.why = "libnccl.so.2 lacks an expected symbol";
    return a;
  }();
  return api;
}
constexpr int kNcclDouble = 8, kNcclSum = 0;

int allreduce_on_stream(boomgpu_ctx *ctx, double *dev, int64_t count) {
  if (!ctx->comm || ctx->comm_ranks <= 1) return 0;
  const int rc = nccl().AllReduce(dev, dev, (size_t)count, kNcclDouble, kNcclSum, ctx->comm, ctx->stream);
  if (rc) return fail(ctx, BOOMGPU_ERR_CUDA, "ncclAllReduce failed: %s", nccl().GetErrorString(rc));
  return 0;
}

SyrkUnitTable make_unit_table() {
  SyrkUnitTable t;
  memset(&t, 0, sizeof(t));
  const int8_t V = 1, D = 2, Y = 4, F = 8;  // valid, diagonal unit, X's duty, fallback X's duty (when unit (ui, ui+1) is cut off by p)
  if (kSyrkConsumerWarps == 8) {
    // off-diagonal region: warp w owns units (w/2, 2(w%2)) and (w/2, 2(w%2)+1) (shared A fragments)
    for (int w = 0; w < 8; ++w) {
      t.u[0][w][0] = {(int8_t)(w / 2), (int8_t)(2 * (w % 2)), 1};
      t.u[0][w][1] = {(int8_t)(w / 2), (int8_t)(2 * (w % 2) + 1), 1};
    }
    // diagonal region: 6 full units (16 DMMA / k-step), 4 diagonal units (10), balanced over the four SM
    // sub-partitions (warp w issues on sub-partition w % 4): 36/36/36/36 + the DFMA of the X's duties.
    t.u[1][0][0] = {0, 1, (int8_t)(V | Y)};
    t.u[1][4][0] = {0, 0, (int8_t)(V | D | F)}; t.u[1][4][1] = {1, 1, (int8_t)(V | D | F)};
    t.u[1][1][0] = {1, 2, (int8_t)(V | Y)};
    t.u[1][5][0] = {0, 2, V};
    t.u[1][2][0] = {2, 3, (int8_t)(V | Y)};
    t.u[1][6][0] = {0, 3, V};
    t.u[1][3][0] = {1, 3, V};
    t.u[1][7][0] = {2, 2, (int8_t)(V | D | F)}; t.u[1][7][1] = {3, 3, (int8_t)(V | D | Y)};
    // ragged last unit column (units (., 3) hold NM < 4 atoms in n).  Off-diagonal: the warps of row blocks 2, 3 take their
    // column pairs in the other order, so each sub-partition (w % 4) gets one whole pair (32 DMMA / k-step) and one cut pair
    // (16 + 4 NM) instead of two of a kind: the region's pace is the busiest sub-partition's.
    for (int w = 0; w < 8; ++w) {
      const int c = (w % 2) ^ (w >= 4 ? 1 : 0);
      t.u[2][w][0] = {(int8_t)(w / 2), (int8_t)(2 * c), 1};
      t.u[2][w][1] = {(int8_t)(w / 2), (int8_t)(2 * c + 1), 1};
    }
    // Diagonal (NM = 3: 28 / 28 / 32 / 32 DMMA per k-step and sub-partition instead of 36 / 32 / 24 / 28)
    t.u[3][0][0] = {0, 1, (int8_t)(V | Y)};  t.u[3][4][0] = {0, 3, V};
    t.u[3][1][0] = {0, 2, V};                t.u[3][5][0] = {1, 3, V};
    t.u[3][2][0] = {1, 2, (int8_t)(V | Y)};  t.u[3][6][0] = {2, 2, (int8_t)(V | D | F)}; t.u[3][6][1] = {3, 3, (int8_t)(V | D | Y)};
    t.u[3][3][0] = {2, 3, (int8_t)(V | Y)};  t.u[3][7][0] = {0, 0, (int8_t)(V | D | F)}; t.u[3][7][1] = {1, 1, (int8_t)(V | D | F)};
  } else {
    // 16 warps x 1 unit.  Off-diagonal: warp w owns unit (w / 4, w % 4): 16 DMMA per k-step on every warp.
    for (int w = 0; w < kSyrkConsumerWarps; ++w) t.u[0][w][0] = {(int8_t)(w / 4), (int8_t)(w % 4), 1};
    // Diagonal: 10 of the 16 warps work; per sub-partition (w % 4): 32 / 32 / 36 / 36 DMMA per k-step.
    t.u[1][0][0] = {0, 1, (int8_t)(V | Y)};  t.u[1][4][0] = {0, 2, V};
    t.u[1][1][0] = {0, 3, V};                t.u[1][5][0] = {1, 2, (int8_t)(V | Y)};
    t.u[1][2][0] = {1, 3, V};                t.u[1][6][0] = {0, 0, (int8_t)(V | D | F)};  t.u[1][10][0] = {1, 1, (int8_t)(V | D | F)};
    t.u[1][3][0] = {2, 3, (int8_t)(V | Y)};  t.u[1][7][0] = {2, 2, (int8_t)(V | D | F)};  t.u[1][11][0] = {3, 3, (int8_t)(V | D | Y)};
    for (int w = 0; w < kSyrkConsumerWarps; ++w) { t.u[2][w][0] = t.u[0][w][0]; t.u[2][w][1] = t.u[0][w][1]; t.u[3][w][0] = t.u[1][w][0]; t.u[3][w][1] = t.u[1][w][1]; }
  }
  return t;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int choose_path(boomgpu_ctx *ctx, int *path) {
  if (ctx->path == 1) {
    if (ctx->p > 64) return fail(ctx, BOOMGPU_ERR_ARG, "option path=1 (fused single pass) needs p <= 64, p = %d", ctx->p);
    *path = 1;
  } else if (ctx->path == 2) {
    *path = 2;
  } else {
    *path = ctx->p <= 64 ? 1 : 2;
  }
  if (*path == 2 && ctx->n >= (int64_t)0x7fffffc0)
    return fail(ctx, BOOMGPU_ERR_ARG, "the two-pass path addresses rows with 32-bit TMA coordinates: fewer than 2^31 rows per context "
                "(n = %lld); shard the rows over more contexts", (long long)ctx->n);
  return 0;
}

// The two-pass path describes X to TMA, which needs a 16-byte aligned base and row pitch.  Uploaded rows are padded for
// that; ADOPTED rows that do not qualify (e.g. a contiguous n x 501 tensor) are copied once into a padded device buffer
// (a second copy of X in HBM -- the price of an unaligned adoption; documented in boomgpu.h).
int ensure_tma_view(boomgpu_ctx *ctx) {
  if (ctx->Xt) return 0;
  if ((ctx->ldx % 2 == 0) && aligned16(ctx->X)) { ctx->Xt = ctx->X; ctx->ldxt = ctx->ldx; return 0; }
  const int64_t ldd = ((int64_t)ctx->p + 7) / 8 * 8;
  CU(cudaMalloc((void **)&ctx->Xt_owned, sizeof(double) * (size_t)std::max<int64_t>(ctx->n * ldd, 1)));
  if (ldd != ctx->p) CU(cudaMemsetAsync(ctx->Xt_owned, 0, sizeof(double) * (size_t)(ctx->n * ldd), ctx->stream));
  CU(cudaMemcpy2DAsync(ctx->Xt_owned, sizeof(double) * ldd, ctx->X, sizeof(double) * ctx->ldx, sizeof(double) * ctx->p, (size_t)ctx->n,
                       cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->Xt = ctx->Xt_owned; ctx->ldxt = ldd;
  return 0;
}

template <int MODEL>
struct SmallLauncher {
  template <int NB>
  static cudaError_t go(boomgpu_ctx *ctx, const RowData &d, const DrawParams &prm, const RowOut &out, int *nparts) {
    auto kern = fused_small_kernel<NB, MODEL>;
    const size_t smem = small_smem_bytes(NB);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kSmallThreads, smem);
    if (e != cudaSuccess) return e;
    occ = std::max(occ, 1);
    const int R = small_rows(NB);
    const int64_t nchunks = (d.n + R - 1) / R;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nchunks, (int64_t)ctx->sms * occ));
    const int64_t need = (int64_t)grid * small_partial_len(NB);
    if (ctx->partials_cap < need) {
      if (ctx->partials) cudaFree(ctx->partials);
      ctx->partials = nullptr; ctx->partials_cap = 0;
      e = cudaMalloc((void **)&ctx->partials, sizeof(double) * (size_t)need);
      if (e != cudaSuccess) return e;
      ctx->partials_cap = need;
    }
    const uint32_t magic = (uint32_t)(0x100000000ull / (uint64_t)d.p) + 1u;
    const int vec2 = (d.p % 2 == 0) && (d.ldx % 2 == 0) && aligned16(d.X);
    {
      LaunchScope ls(ctx, 0);
      kern<<<grid, kSmallThreads, smem, ctx->stream>>>(d, prm, out, ctx->beta_dev, ctx->partials, ctx->err_dev, magic, vec2);
    }
    *nparts = grid;
    return cudaGetLastError();
  }
  static cudaError_t dispatch(boomgpu_ctx *ctx, int nb, const RowData &d, const DrawParams &prm, const RowOut &out, int *nparts) {
    switch (nb) {
      case 1: return go<1>(ctx, d, prm, out, nparts);
      case 2: return go<2>(ctx, d, prm, out, nparts);
      case 3: return go<3>(ctx, d, prm, out, nparts);
      case 4: return go<4>(ctx, d, prm, out, nparts);
      case 5: return go<5>(ctx, d, prm, out, nparts);
      case 6: return go<6>(ctx, d, prm, out, nparts);
      case 7: return go<7>(ctx, d, prm, out, nparts);
      case 8: return go<8>(ctx, d, prm, out, nparts);
    }
    return cudaErrorInvalidValue;
  }
};

// ---- TMA descriptor: row-major X as a 2-D tensor {p, n}, box {8 NB + 4, 32} (wider than p: pad columns read as zero)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int ensure_xmap(boomgpu_ctx *ctx, boomgpu_ctx::XMap &m, const double *X, int64_t ldx, int box_cols, int box_rows, int promo = 3) {
  if (m.X == X && m.n == ctx->n && m.ldx == ldx && m.p == ctx->p && m.box_cols == box_cols && m.box_rows == box_rows && m.promo == promo) return 0;
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(ctx, BOOMGPU_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)ctx->p, (cuuint64_t)ctx->n};
  const cuuint64_t strides[1] = {(cuuint64_t)ldx * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = encode(&m.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(X), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                      : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, BOOMGPU_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for n=%lld p=%d ldx=%lld box=%dx%d", (int)r,
                                     (long long)ctx->n, ctx->p, (long long)ldx, box_cols, box_rows);
  m.X = X; m.n = ctx->n; m.ldx = ldx; m.p = ctx->p; m.box_cols = box_cols; m.box_rows = box_rows; m.promo = promo;
  return 0;
}

// X can be described to TMA: 16-byte aligned base and row pitch, rows addressable with 32-bit coordinates
bool tma_ok(const boomgpu_ctx *ctx) { return (ctx->ldx % 2 == 0) && aligned16(ctx->X) && ctx->n < (int64_t)0x7fffffc0; }

template <int MODEL>
struct TmaLauncher {
  template <int NB>
  static int go(boomgpu_ctx *ctx, const RowData &d, const DrawParams &prm, const RowOut &out, int *nparts, const TailParams &tail) {
    // wide tiles, the two draw models: accumulators parked in tensor memory between DMMA phases -> 12 warps instead of 8
    // (registers are allocated per warpgroup on sm_100: 10 warps get the 168 registers of 12, so there is no 10-warp form)
    // Measured (profiles/bench_r02/run27_tmem_park.txt, ms per launch, 8 warps -> 12 warps parked): Poisson p = 40 / 48 / 50 / 64
    // 0.561 -> 0.483, 0.658 -> 0.588, 0.233 -> 0.222, 0.988 -> 1.063; logit p = 40 / 48 / 56 / 64 0.832 -> 0.728, 1.012 -> 0.911,
    // 1.24 -> 2.06, 1.62 -> 3.74 (at NB >= 7 the logit draw spills inside the k-steps at 168 registers).  So: automatic for
    // NB = 5, 6 and for the Poisson model at NB = 7; small_variant = 2 never parks, 4 parks wherever the form exists.
    if constexpr (NB >= 5 && (MODEL == kLogit || MODEL == kPoisson)) {
      const bool faster = NB <= 6 || (NB == 7 && MODEL == kPoisson);
      if (ctx->small_variant == 4 || (ctx->small_variant == 0 && faster)) return go_impl<NB, 12, true>(ctx, d, prm, out, nparts, tail);
    }
    return go_impl<NB, tma_warps(NB), false>(ctx, d, prm, out, nparts, tail);
  }
  template <int NB, int NW, bool PARK>
  static int go_impl(boomgpu_ctx *ctx, const RowData &d, const DrawParams &prm, const RowOut &out, int *nparts, const TailParams &tail) {
    auto kern = fused_tma_kernel<NB, MODEL, NW, PARK>;
    BetaParam bp;
    memset(&bp, 0, sizeof(bp));
    memcpy(bp.b, ctx->beta_pin, sizeof(double) * d.p);
    const size_t smem = tma_smem_bytes(NB, NW, PARK);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (int rc = ensure_xmap(ctx, ctx->xmap_small, ctx->X, ctx->ldx, tma_padw(NB), tma_slice_rows(NB))) return rc;
    const int64_t nslices = (d.n + tma_slice_rows(NB) - 1) / tma_slice_rows(NB);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nslices + NW - 1) / NW, (int64_t)ctx->sms));
    if (ensure(ctx, &ctx->partials, &ctx->partials_cap, (int64_t)grid * tma_partial_len(NB))) return BOOMGPU_ERR_CUDA;
    {
      LaunchScope ls(ctx, 0);
      kern<<<grid, 32 * NW, smem, ctx->stream>>>(ctx->xmap_small.map, d, prm, out, bp, ctx->partials, ctx->err_dev, tail);
    }
    CU(cudaGetLastError());
    *nparts = grid;
    return 0;
  }
  static int dispatch(boomgpu_ctx *ctx, int nb, const RowData &d, const DrawParams &prm, const RowOut &out, int *nparts,
                      const TailParams &tail) {
    switch (nb) {
      case 1: return go<1>(ctx, d, prm, out, nparts, tail);
      case 2: return go<2>(ctx, d, prm, out, nparts, tail);
      case 3: return go<3>(ctx, d, prm, out, nparts, tail);
      case 4: return go<4>(ctx, d, prm, out, nparts, tail);
      case 5: return go<5>(ctx, d, prm, out, nparts, tail);
      case 6: return go<6>(ctx, d, prm, out, nparts, tail);
      case 7: return go<7>(ctx, d, prm, out, nparts, tail);
      case 8: return go<8>(ctx, d, prm, out, nparts, tail);
    }
    return fail(ctx, BOOMGPU_ERR_ARG, "bad column block count %d", nb);
  }
};

// the warp-specialised wide-tile kernel (fused_ws_kernel): 40 < p <= 64
template <int MODEL>
struct WsLauncher {
  template <int NB>
  static int go(boomgpu_ctx *ctx, const RowData &d, const DrawParams &prm, const RowOut &out, int *nparts, const TailParams &tail) {
    auto kern = fused_ws_kernel<NB, MODEL>;
    BetaParam bp;
    memset(&bp, 0, sizeof(bp));
    memcpy(bp.b, ctx->beta_pin, sizeof(double) * d.p);
    const size_t smem = ws_smem_bytes(NB);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (int rc = ensure_xmap(ctx, ctx->xmap_small, ctx->X, ctx->ldx, tma_padw(NB), 32)) return rc;
    const int64_t nslices = (d.n + 31) / 32;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nslices + kWsRings - 1) / kWsRings, (int64_t)ctx->sms));
    if (ensure(ctx, &ctx->partials, &ctx->partials_cap, (int64_t)grid * tma_partial_len(NB))) return BOOMGPU_ERR_CUDA;
    {
      LaunchScope ls(ctx, 0);
      kern<<<grid, 32 * kWsWarps, smem, ctx->stream>>>(ctx->xmap_small.map, d, prm, out, bp, ctx->partials, ctx->err_dev, tail);
    }
    CU(cudaGetLastError());
    *nparts = grid;
    return 0;
  }
  static int dispatch(boomgpu_ctx *ctx, int nb, const RowData &d, const DrawParams &prm, const RowOut &out, int *nparts,
                      const TailParams &tail) {
    switch (nb) {
      case 5: return go<5>(ctx, d, prm, out, nparts, tail);
      case 6: return go<6>(ctx, d, prm, out, nparts, tail);
      case 7: return go<7>(ctx, d, prm, out, nparts, tail);
      case 8: return go<8>(ctx, d, prm, out, nparts, tail);
    }
    return fail(ctx, BOOMGPU_ERR_ARG, "bad column block count %d for the warp-specialised kernel", nb);
  }
};

template <int MODEL>
cudaError_t launch_impute_rows(boomgpu_ctx *ctx, const RowData &d, const DrawParams &prm, const RowOut &out, int *nparts, int nnz) {
  const bool vec2 = (d.ldx % 2 == 0) && aligned16(d.X);
  const bool gather = nnz >= 0;
  const size_t smem = gather ? sizeof(double) * (size_t)(nnz + 1) + sizeof(int) * (size_t)(nnz + 2) : sizeof(double) * (size_t)(d.p + 2);
  const int64_t ngroups = (d.n + 31) / 32;
  int occ = 1;
  cudaError_t e;
  if (gather) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, impute_rows_gather_kernel<MODEL>, kImputeThreads, smem);
  else if (vec2) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, impute_rows_kernel<MODEL, true>, kImputeThreads, smem);
  else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, impute_rows_kernel<MODEL, false>, kImputeThreads, smem);
  if (e != cudaSuccess) return e;
  occ = std::max(occ, 1);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ngroups + 7) / 8, (int64_t)ctx->sms * occ));
  if (ctx->scal_cap < (int64_t)grid * 4) {
    if (ctx->scal_partials) cudaFree(ctx->scal_partials);
    ctx->scal_partials = nullptr; ctx->scal_cap = 0;
    e = cudaMalloc((void **)&ctx->scal_partials, sizeof(double) * (size_t)grid * 4);
    if (e != cudaSuccess) return e;
    ctx->scal_cap = (int64_t)grid * 4;
  }
  {
    LaunchScope ls(ctx, 1);
    if (gather)
      impute_rows_gather_kernel<MODEL><<<grid, kImputeThreads, smem, ctx->stream>>>(
          d, prm, out, ctx->beta_dev, reinterpret_cast<const int *>(ctx->beta_dev + d.p + 2), nnz, ctx->w_buf, ctx->s_buf,
          ctx->scal_partials, ctx->err_dev);
    else if (vec2)
      impute_rows_kernel<MODEL, true><<<grid, kImputeThreads, smem, ctx->stream>>>(d, prm, out, ctx->beta_dev, ctx->w_buf,
                                                                                  ctx->s_buf, ctx->scal_partials, ctx->err_dev);
    else
      impute_rows_kernel<MODEL, false><<<grid, kImputeThreads, smem, ctx->stream>>>(d, prm, out, ctx->beta_dev, ctx->w_buf,
                                                                                   ctx->s_buf, ctx->scal_partials, ctx->err_dev);
  }
  *nparts = grid;
  return cudaGetLastError();
}

int ensure_ws(boomgpu_ctx *ctx) {
  const int64_t npad = ((ctx->n + kSyrkKB - 1) / kSyrkKB) * kSyrkKB + kSyrkKB;
  if (ctx->ws_cap < npad) {
    if (ctx->w_buf) { CU(cudaFree(ctx->w_buf)); ctx->w_buf = nullptr; }
    if (ctx->s_buf) { CU(cudaFree(ctx->s_buf)); ctx->s_buf = nullptr; }
    CU(cudaMalloc((void **)&ctx->w_buf, sizeof(double) * (size_t)npad));
    CU(cudaMalloc((void **)&ctx->s_buf, sizeof(double) * (size_t)npad));
    ctx->ws_cap = npad;
  }
  // rows beyond n must weigh nothing
  CU(cudaMemsetAsync(ctx->w_buf + ctx->n, 0, sizeof(double) * (size_t)(npad - ctx->n), ctx->stream));
  CU(cudaMemsetAsync(ctx->s_buf + ctx->n, 0, sizeof(double) * (size_t)(npad - ctx->n), ctx->stream));
  return 0;
}

// pass 2 + reduction into suf (device)
// Order 2 of the SYRK grid: the work items of one k-slice in launch order.  Off-diagonal regions are listed super-tile by
// super-tile (12 x 12 regions: about the 148 CTAs that run at one time, which then touch 24 distinct panels instead of the
// 1 + 148 of a row-by-row order -- what made C4's shape read X 18 times), every diagonal super-tile followed by its
// diagonal regions in pairs (2q, 2q + 1); an odd last diagonal region stands alone.
int ensure_syrk_items(boomgpu_ctx *ctx, int nblk) {
  if (ctx->syrk_items && ctx->syrk_items_nblk == nblk) return 0;
  std::vector<SyrkItem> items;
  constexpr int T = 12;
  const int nsuper = (nblk + T - 1) / T;
  for (int SI = 0; SI < nsuper; ++SI)
    for (int SJ = SI; SJ < nsuper; ++SJ) {
      for (int I = SI * T; I < std::min(nblk, (SI + 1) * T); ++I)
        for (int J = std::max(I + 1, SJ * T); J < std::min(nblk, (SJ + 1) * T); ++J) items.push_back({(int16_t)I, (int16_t)J, 0, 0});
      if (SJ == SI) {
        const int lo = SI * T, hi = std::min(nblk, (SI + 1) * T);
        int I = lo;
        for (; I + 1 < hi; I += 2) items.push_back({(int16_t)I, (int16_t)(I + 1), 2, 0});
        if (I < hi) items.push_back({(int16_t)I, (int16_t)I, 1, 0});
      }
    }
  if (ctx->syrk_items) { CU(cudaFree(ctx->syrk_items)); ctx->syrk_items = nullptr; }
  CU(cudaMalloc((void **)&ctx->syrk_items, sizeof(SyrkItem) * items.size()));
  CU(cudaMemcpyAsync(ctx->syrk_items, items.data(), sizeof(SyrkItem) * items.size(), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));   // items is a local
  ctx->syrk_items_nblk = nblk; ctx->syrk_nitems = (int)items.size();
  return 0;
}

int launch_syrk(boomgpu_ctx *ctx, double *suf) {
  SyrkParams sp;
  sp.X = ctx->Xt; sp.ldx = ctx->ldxt; sp.n = ctx->n; sp.p = ctx->p;
  sp.w = ctx->w_buf; sp.s = ctx->s_buf;
  sp.nblk = (ctx->p + 127) / 128;
  sp.nregions = sp.nblk * (sp.nblk + 1) / 2;
  // ~30 CTAs per SM for dynamic balance (measured, profiles/bench_r02/run22_syrk_order.jsonl: 20 -> 30 is worth 0.5-1.5 %
  // with the off-diagonal-first order; more only adds partial tiles), at least 2048 rows per CTA
  sp.order = (ctx->syrk_order == 2 && kSyrkConsumerWarps != 8) ? 1 : ctx->syrk_order;
  sp.items = nullptr; sp.nitems = 0;
  int per_slice = sp.nregions;                     // CTAs per k-slice
  if (sp.order == 2) {
    if (int rc = ensure_syrk_items(ctx, sp.nblk)) return rc;
    sp.items = ctx->syrk_items; sp.nitems = ctx->syrk_nitems;
    per_slice = sp.nitems;
  }
  // one region (64 < p <= 128): uniform CTAs need no fine split for balance, and every k-slice costs a partial tile to write and
  // to sum -- 4 CTAs per SM instead of 30 (run 55: p = 70 / 100 / 128 3.26 / 4.68 / 2.79 -> 2.67 / 4.04 / 2.43 ms)
  const int waves = sp.nregions == 1 ? std::min(ctx->syrk_waves, 4) : ctx->syrk_waves;
  int64_t ksplit = (waves * (int64_t)ctx->sms + per_slice - 1) / per_slice;
  if (sp.order == 2) {
    // uniform CTAs: a grid that is a whole number of waves has no tail at all
    const int64_t unit = ctx->sms / std::gcd((int64_t)ctx->sms, (int64_t)per_slice);
    ksplit = std::max<int64_t>(unit, (ksplit + unit / 2) / unit * unit);
  }
  ksplit = std::min<int64_t>(ksplit, std::max<int64_t>(1, ctx->n / 2048));
  ksplit = std::max<int64_t>(ksplit, 1);
  int64_t rows = (ctx->n + ksplit - 1) / ksplit;
  rows = ((rows + kSyrkKB - 1) / kSyrkKB) * kSyrkKB;
  ksplit = std::max<int64_t>(1, (ctx->n + rows - 1) / rows);
  sp.ksplit = (int)ksplit;
  sp.rows_per_slice = rows;
  const int64_t need = ksplit * sp.nregions * kSyrkTileLen;
  if (ensure(ctx, &ctx->partials, &ctx->partials_cap, need)) return BOOMGPU_ERR_CUDA;
  sp.partials = ctx->partials;
  sp.filter = ctx->syrk_filter;
  sp.diag_form = ctx->syrk_diag;
  // the ragged last diagonal region (p not a multiple of 128) in its own kernel, unless order 2 has paired it with its neighbour
  const int rem_cols = ((ctx->p + 7) & ~7) - 128 * (sp.nblk - 1);
  const int ragged_atoms = rem_cols < 128 ? rem_cols / 8 : 0;
  const bool side = ragged_atoms > 0 && ctx->syrk_rdiag && kSyrkConsumerWarps == 8 && ctx->syrk_filter == 0 && ctx->syrk_diag == 0 &&
                    !(sp.order == 2 && sp.nblk % 2 == 0);
  // ... and, when that block is narrow (<= 12 atom columns; wider ones do as well in the balanced unit form), its off-diagonal regions
  const bool side_off = side && sp.nblk > 1 && ragged_atoms <= 12 && sp.order != 2;
  sp.skip_ragged_diag = side ? (side_off ? 3 : 1) : 0;
  static const SyrkUnitTable table = make_unit_table();
  if (int rc = ensure_xmap(ctx, ctx->xmap_syrk, ctx->Xt, ctx->ldxt, kSyrkPanelLd, kSyrkKB, ctx->tma_promotion)) return rc;
  CU(cudaFuncSetAttribute(syrk_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrkSmemBytes));
  if (!(side && sp.nregions == 1)) {   // (64 < p < 128: the ragged diagonal region is the only one)
    LaunchScope ls(ctx, 2);
    const unsigned grid = (unsigned)(ksplit * per_slice);
    if (ctx->syrk_cluster > 1 && grid % (unsigned)ctx->syrk_cluster == 0) {
      // experiment (option "syrk_cluster"): co-schedule c consecutive CTAs -- with the off-diagonal-first order the regions of
      // one k-slice that share panels -- as a thread-block cluster, so that they start together and meet in L2
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kSyrkThreads); cfg.dynamicSmemBytes = kSyrkSmemBytes; cfg.stream = ctx->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)ctx->syrk_cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      if (ctx->syrk_cluster > 8) cudaFuncSetAttribute(syrk_dmma_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      CU(cudaLaunchKernelEx(&cfg, syrk_dmma_kernel, ctx->xmap_syrk.map, sp, table));
    } else {
      syrk_dmma_kernel<<<grid, kSyrkThreads, kSyrkSmemBytes, ctx->stream>>>(ctx->xmap_syrk.map, sp, table);
    }
  }
  CU(cudaGetLastError());
  if (side) {
    LaunchScope ls(ctx, 2);
    const unsigned grid = (unsigned)(ksplit * (side_off ? sp.nblk : 1));
#define BOOMGPU_RDIAG(A_)                                                                                                   \
  case A_:                                                                                                                  \
    CU(cudaFuncSetAttribute(syrk_rdiag_kernel<A_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrkSmemBytes));      \
    syrk_rdiag_kernel<A_><<<grid, kSyrkThreads, kSyrkSmemBytes, ctx->stream>>>(ctx->xmap_syrk.map, sp);                      \
    break;
    switch (ragged_atoms) {
      BOOMGPU_RDIAG(1) BOOMGPU_RDIAG(2) BOOMGPU_RDIAG(3) BOOMGPU_RDIAG(4) BOOMGPU_RDIAG(5) BOOMGPU_RDIAG(6) BOOMGPU_RDIAG(7)
      BOOMGPU_RDIAG(8) BOOMGPU_RDIAG(9) BOOMGPU_RDIAG(10) BOOMGPU_RDIAG(11) BOOMGPU_RDIAG(12) BOOMGPU_RDIAG(13) BOOMGPU_RDIAG(14)
      BOOMGPU_RDIAG(15)
      default: return fail(ctx, BOOMGPU_ERR_ARG, "internal: ragged diagonal region with %d atom columns", ragged_atoms);
    }
#undef BOOMGPU_RDIAG
    CU(cudaGetLastError());
  }
  {
    LaunchScope ls(ctx, 3);
    const int64_t total = (int64_t)sp.nregions * kSyrkTileLen;
    reduce_syrk_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(sp, suf);
  }
  CU(cudaGetLastError());
  return 0;
}

// out[0 .. p) = X' svec (svec: n doubles on the device)
int launch_xts_vec(boomgpu_ctx *ctx, const double *svec, double *out) {
  const int p = ctx->p;
  const int p2 = (p + 1) / 2;
  const int pairs = (p2 + kXtsThreads - 1) / kXtsThreads;
  if (pairs > kXtsMaxPairs) return fail(ctx, BOOMGPU_ERR_ARG, "p = %d is too wide for the X's kernel", p);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ctx->n + 255) / 256, (int64_t)ctx->sms * 4));
  if (ensure(ctx, &ctx->partials, &ctx->partials_cap, (int64_t)grid * 2 * p2)) return BOOMGPU_ERR_CUDA;
  {
    LaunchScope ls(ctx, 2);
    auto go = [&](auto tag) {
      constexpr int P = decltype(tag)::value;
      xts_kernel<P><<<grid, kXtsThreads, 0, ctx->stream>>>(ctx->Xt, ctx->ldxt, ctx->n, p, svec, ctx->partials);
    };
    if (pairs <= 1) go(std::integral_constant<int, 1>());
    else if (pairs <= 2) go(std::integral_constant<int, 2>());
    else if (pairs <= 4) go(std::integral_constant<int, 4>());
    else if (pairs <= 8) go(std::integral_constant<int, 8>());
    else if (pairs <= 16) go(std::integral_constant<int, 16>());
    else go(std::integral_constant<int, 32>());
  }
  CU(cudaGetLastError());
  {
    LaunchScope ls(ctx, 3);
    reduce_xts_kernel<<<(p + 255) / 256, 256, 0, ctx->stream>>>(ctx->partials, grid, p, out);
  }
  CU(cudaGetLastError());
  return 0;
}

// X's alone into suf[p*p .. p*p + p) (the matrix part is zeroed: it is not recomputed)
int launch_xts(boomgpu_ctx *ctx, double *suf) {
  CU(cudaMemsetAsync(suf, 0, sizeof(double) * (size_t)ctx->p * ctx->p, ctx->stream));
  return launch_xts_vec(ctx, ctx->s_buf, suf + (int64_t)ctx->p * ctx->p);
}

// G = X' diag(w) X_A, diag, X's from the latents in w_buf / s_buf: [G p x ka8 | diag p | xty p] at out (device)
int launch_panel(boomgpu_ctx *ctx, double *out, int *ka8_out) {
  const int k = (int)ctx->sel_cols.size();
  static const int sizes[] = {1, 2, 3, 4, 6, 8, 12, 16};
  int nba = 0;
  for (int v : sizes) if (8 * v >= k) { nba = v; break; }
  if (!nba) return fail(ctx, BOOMGPU_ERR_ARG, "active set of %d columns exceeds the supported 128", k);
  PanelParams pp;
  pp.n = ctx->n; pp.p = ctx->p; pp.nblk = (ctx->p + 127) / 128; pp.ka8 = 8 * nba;
  pp.w = ctx->w_buf; pp.s = ctx->s_buf;
  int64_t ksplit = (8 * (int64_t)ctx->sms + pp.nblk - 1) / pp.nblk;
  ksplit = std::max<int64_t>(1, std::min<int64_t>(ksplit, std::max<int64_t>(1, ctx->n / 2048)));
  int64_t rows = (ctx->n + ksplit - 1) / ksplit;
  rows = ((rows + kSyrkKB - 1) / kSyrkKB) * kSyrkKB;
  ksplit = std::max<int64_t>(1, (ctx->n + rows - 1) / rows);
  pp.ksplit = (int)ksplit; pp.rows_per_slice = rows;
  if (ensure(ctx, &ctx->partials, &ctx->partials_cap, ksplit * pp.nblk * kPanelTileLen)) return BOOMGPU_ERR_CUDA;
  pp.partials = ctx->partials;
  if (int rc = ensure_xmap(ctx, ctx->xmap_syrk, ctx->Xt, ctx->ldxt, kSyrkPanelLd, kSyrkKB, ctx->tma_promotion)) return rc;
  {   // the gathered matrix as a tensor {k columns, n rows}: the box is wider than k, the excess reads as zero
    boomgpu_ctx::XMap &m = ctx->xmap_sel;
    const int p_save = ctx->p;
    ctx->p = k;
    const int rc = ensure_xmap(ctx, m, ctx->Xsel, ctx->sel_ld, 8 * nba + 4, kSyrkKB, ctx->tma_promotion);
    ctx->p = p_save;
    if (rc) return rc;
  }
  const unsigned grid = (unsigned)(ksplit * pp.nblk);
  {
    LaunchScope ls(ctx, 2);
    auto go = [&](auto tag) -> cudaError_t {
      constexpr int NBA = decltype(tag)::value;
      cudaError_t e = cudaFuncSetAttribute(panel_dmma_kernel<NBA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrkSmemBytes);
      if (e != cudaSuccess) return e;
      panel_dmma_kernel<NBA><<<grid, kSyrkThreads, kSyrkSmemBytes, ctx->stream>>>(ctx->xmap_syrk.map, ctx->xmap_sel.map, pp);
      return cudaGetLastError();
    };
    cudaError_t e;
    switch (nba) {
      case 1: e = go(std::integral_constant<int, 1>()); break;
      case 2: e = go(std::integral_constant<int, 2>()); break;
      case 3: e = go(std::integral_constant<int, 3>()); break;
      case 4: e = go(std::integral_constant<int, 4>()); break;
      case 6: e = go(std::integral_constant<int, 6>()); break;
      case 8: e = go(std::integral_constant<int, 8>()); break;
      case 12: e = go(std::integral_constant<int, 12>()); break;
      default: e = go(std::integral_constant<int, 16>()); break;
    }
    if (e != cudaSuccess) return fail(ctx, BOOMGPU_ERR_CUDA, "panel_dmma_kernel launch failed: %s", cudaGetErrorString(e));
  }
  {
    LaunchScope ls(ctx, 3);
    const int64_t total = (int64_t)pp.nblk * 128 * (pp.ka8 + 2);
    reduce_panel_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, ctx->stream>>>(
        pp, out, out + (int64_t)ctx->p * pp.ka8, out + (int64_t)ctx->p * pp.ka8 + ctx->p);
  }
  CU(cudaGetLastError());
  *ka8_out = pp.ka8;
  return 0;
}

// The whole device step for MODEL; leaves the packed statistics at suf (device).
template <int MODEL>
int run_step(boomgpu_ctx *ctx, const double *beta_host, const DrawParams &prm, const RowOut &out,
             const double *w_in, const double *s_in, double *suf, double *host_out = nullptr) {
  int path = 0;
  if (int rc = choose_path(ctx, &path)) return rc;
  const int p = ctx->p;
  ctx->host_out_written = false;
  const bool tma_small = path == 1 && tma_ok(ctx) && ctx->small_variant != 1;
  if (ctx->beta_cap < p + 2) {
    if (ctx->beta_dev) { CU(cudaFree(ctx->beta_dev)); ctx->beta_dev = nullptr; }
    if (ctx->beta_pin) { CU(cudaFreeHost(ctx->beta_pin)); ctx->beta_pin = nullptr; }
    const size_t bytes = sizeof(double) * (size_t)(p + 2) + sizeof(int) * (size_t)(p + 2);
    CU(cudaMalloc((void **)&ctx->beta_dev, bytes));
    CU(cudaMallocHost((void **)&ctx->beta_pin, bytes));
    ctx->beta_cap = p + 2;
  }
  if (MODEL != kSupplied) {
    memcpy(ctx->beta_pin, beta_host, sizeof(double) * p);
  } else {
    memset(ctx->beta_pin, 0, sizeof(double) * p);
  }
  // sparse beta (spike and slab): the included columns, for the gather pass.  Worth it while nnz 32-byte sectors per row
  // are fewer bytes than the 8 p byte row, i.e. nnz < p / 4 (auto); the dense pass otherwise.
  int nnz = -1;
  if (path == 2 && MODEL != kSupplied && ctx->gather != 1) {
    int *idx = reinterpret_cast<int *>(ctx->beta_pin + p + 2);
    int c = 0;
    for (int j = 0; j < p; ++j) if (ctx->beta_pin[j] != 0.0) idx[c++] = j;
    if (ctx->gather == 2 ? c < p : 4 * c < p) nnz = c;
  }
  if (!tma_small)   // the TMA small-p kernel takes beta as a kernel parameter
    CU(cudaMemcpyAsync(ctx->beta_dev, ctx->beta_pin, sizeof(double) * (size_t)(p + 2) + (nnz >= 0 ? sizeof(int) * (size_t)nnz : 0),
                       cudaMemcpyHostToDevice, ctx->stream));

  RowData d;
  d.X = ctx->X; d.ldx = ctx->ldx; d.n = ctx->n; d.p = p;
  d.y = ctx->y; d.ntrials = ctx->ntrials; d.yi = ctx->yi; d.exposure = ctx->exposure;
  d.w_in = w_in; d.s_in = s_in; d.row_offset = ctx->row_offset;

  if (ctx->n == 0) {
    CU(cudaMemsetAsync(suf, 0, sizeof(double) * (size_t)boomgpu_suf_len(p), ctx->stream));
    return 0;
  }
  if (path == 1) {
    ctx->latents_valid = false;
    int nparts = 0;
    const int nb = (p + 7) / 8;
    bool reduced = false;
    if (tma_small) {
      // TMA-fed warp-autonomous kernel (fused_tma.cuh); single launch: its last CTA also sums the partials
      TailParams tail{suf, host_out, ctx->single_launch ? ctx->tail_counter : nullptr};
      if (nb >= 5 && ctx->small_variant == 3) {   // wide tiles, opt-in: accumulate warps + draw warps (fused_ws_kernel)
        if (int rc = WsLauncher<MODEL>::dispatch(ctx, nb, d, prm, out, &nparts, tail)) return rc;
      } else if (int rc = TmaLauncher<MODEL>::dispatch(ctx, nb, d, prm, out, &nparts, tail)) return rc;
      reduced = tail.counter != nullptr;
      if (reduced) ctx->host_out_written = host_out != nullptr;
    } else {
      // X cannot be described to TMA (odd leading dimension / unaligned adopted pointer): cp.async variant
      cudaError_t e = SmallLauncher<MODEL>::dispatch(ctx, nb, d, prm, out, &nparts);
      if (e != cudaSuccess) return fail(ctx, BOOMGPU_ERR_CUDA, "fused_small_kernel launch failed: %s", cudaGetErrorString(e));
    }
    if (!reduced) {
      LaunchScope ls(ctx, 3);
      const int total = p * (p + 1) / 2 + p + 4;   // one warp per output element
      reduce_partials_kernel<<<(total + 7) / 8, 256, 0, ctx->stream>>>(ctx->partials, nparts, nb, p, suf, host_out, ctx->err_dev);
      ctx->host_out_written = host_out != nullptr;
    }
    CU(cudaGetLastError());
  } else {
    if (int rc = ensure_tma_view(ctx)) return rc;
    d.X = ctx->Xt; d.ldx = ctx->ldxt;
    if (int rc = ensure_ws(ctx)) return rc;
    int nparts = 0;
    if (MODEL == kSupplied) {
      CU(cudaMemcpyAsync(ctx->w_buf, w_in, sizeof(double) * (size_t)ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
      CU(cudaMemcpyAsync(ctx->s_buf, s_in, sizeof(double) * (size_t)ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
      CU(cudaMemsetAsync(suf + (int64_t)p * p + p, 0, sizeof(double) * 4, ctx->stream));
    } else {
      cudaError_t e = launch_impute_rows<MODEL>(ctx, d, prm, out, &nparts, nnz);
      if (e != cudaSuccess) return fail(ctx, BOOMGPU_ERR_CUDA, "impute_rows_kernel launch failed: %s", cudaGetErrorString(e));
      LaunchScope ls(ctx, 3);
      reduce_scalars_kernel<<<1, 32, 0, ctx->stream>>>(ctx->scal_partials, nparts, suf + (int64_t)p * p + p);
    }
    CU(cudaGetLastError());
    ctx->latents_valid = MODEL != kSupplied;
    if (ctx->stats_mode == 1) return 0;   // active-set step: the caller launches the panel product
    if (ctx->xty_only && MODEL == kProbit) { if (int rc = launch_xts(ctx, suf)) return rc; }
    else if (int rc = launch_syrk(ctx, suf)) return rc;
  }
  return 0;
}

// the Philox counter carries 16 bits of the per-row slot (draws.cuh: uniform_pair) and the per-trial branch uses slot = trial
// index, so a row may run at most 65535 trials one by one; larger n_i must take the CLT branch
int check_clt(boomgpu_ctx *ctx, int clt_threshold) {
  if (clt_threshold < 0 || clt_threshold >= 65536)
    return fail(ctx, BOOMGPU_ERR_ARG, "clt_threshold = %d: must lie in [0, 65535] (the random stream has 16 bits of per-row trial index)", clt_threshold);
  return 0;
}

int check_ready(boomgpu_ctx *ctx, int model) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (ctx->model != model || !ctx->X)
    return fail(ctx, BOOMGPU_ERR_STATE, "no %s data uploaded to this context", model == kLogit ? "binomial" : "Poisson");
  if (model == kLogit && !ctx->have_mix) return fail(ctx, BOOMGPU_ERR_STATE, "boomgpu_set_logit_mixture has not been called");
  if (model == kPoisson && !ctx->have_tab) return fail(ctx, BOOMGPU_ERR_STATE, "boomgpu_set_poisson_table has not been called");
  return 0;
}

int ensure_suf(boomgpu_ctx *ctx) {
  const int64_t len = boomgpu_suf_len(ctx->p);
  if (ensure(ctx, &ctx->suf_dev, &ctx->suf_cap, len)) return BOOMGPU_ERR_CUDA;
  if (ctx->suf_pin_cap < len + 1) {   // + 1: the validation flag rides behind the statistics on the zero-copy path
    if (ctx->suf_pin) { CU(cudaFreeHost(ctx->suf_pin)); ctx->suf_pin = nullptr; }
    CU(cudaMallocHost((void **)&ctx->suf_pin, sizeof(double) * (size_t)(len + 1)));
    ctx->suf_pin_cap = len + 1;
  }
  return 0;
}

int finish_and_check(boomgpu_ctx *ctx) {
  CU(cudaMemcpyAsync(ctx->err_pin, ctx->err_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const int flag = *ctx->err_pin;
  if (flag) {
    CU(cudaMemsetAsync(ctx->err_dev, 0, sizeof(int), ctx->stream));
    if (flag & 1) return fail(ctx, BOOMGPU_ERR_DATA, "a count y is missing from the Poisson mixture table "
                              "(call NormalMixtureApproximationTable::approximate(y) for every distinct y before upload)");
    return fail(ctx, BOOMGPU_ERR_DATA, "invalid observation on the device: successes > trials, a negative count/exposure, "
                "or a non-finite linear predictor");
  }
  return 0;
}

bool is_pinned_host(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// statistics of the step just launched -> host, then wait and report device-side validation errors.
//   * single GPU + small-p kernel: the reduction already wrote them (and the flag) into the host-mapped suf_pin;
//   * xtx_direct (a page-locked destination for the p x p matrix, large p): the matrix is copied straight into it and
//     only the [p | 4] tail goes through suf_pin; *matrix_in_place tells the caller not to copy the matrix again.
int fetch_suf(boomgpu_ctx *ctx, double *xtx_direct = nullptr, bool *matrix_in_place = nullptr) {
  const int p = ctx->p;
  const int64_t len = boomgpu_suf_len(p);
  if (matrix_in_place) *matrix_in_place = false;
  if (ctx->host_out_written) {
    CU(cudaStreamSynchronize(ctx->stream));
    const int flag = (int)ctx->suf_pin[len];
    if (flag) {
      CU(cudaMemsetAsync(ctx->err_dev, 0, sizeof(int), ctx->stream));
      if (flag & 1) return fail(ctx, BOOMGPU_ERR_DATA, "a count y is missing from the Poisson mixture table "
                                "(call NormalMixtureApproximationTable::approximate(y) for every distinct y before upload)");
      return fail(ctx, BOOMGPU_ERR_DATA, "invalid observation on the device: successes > trials, a negative count/exposure, "
                  "or a non-finite linear predictor");
    }
    return 0;
  }
  const size_t mat = (size_t)p * p;
  if (xtx_direct && mat * sizeof(double) >= ((size_t)1 << 20) && is_pinned_host(xtx_direct)) {
    CU(cudaMemcpyAsync(xtx_direct, ctx->suf_dev, sizeof(double) * mat, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ctx->suf_pin + mat, ctx->suf_dev + mat, sizeof(double) * (size_t)(p + 4), cudaMemcpyDeviceToHost, ctx->stream));
    if (matrix_in_place) *matrix_in_place = true;
    return finish_and_check(ctx);
  }
  CU(cudaMemcpyAsync(ctx->suf_pin, ctx->suf_dev, sizeof(double) * (size_t)len, cudaMemcpyDeviceToHost, ctx->stream));
  return finish_and_check(ctx);
}

DrawParams make_prm(boomgpu_ctx *ctx, int clt, uint64_t seed, uint64_t iteration) {
  DrawParams prm;
  prm.hot = ctx->hot;
  prm.ext = ctx->ext_hot;
  prm.mix = ctx->mix_dev;
  prm.tab = ctx->tab;
  prm.key.seed = seed;
  prm.key.iteration = iteration;
  philox_round_keys(prm.key);
  prm.clt_threshold = clt;
  prm.log_alpha = 0.0;
  prm.t_inv_sigma = 0.0; prm.t_nu = 0.0;
  return prm;
}

template <class T>
int upload_array(boomgpu_ctx *ctx, const T *host, int64_t count, const T **dev_out) {
  T *dptr = nullptr;
  CU(cudaMalloc((void **)&dptr, sizeof(T) * (size_t)std::max<int64_t>(count, 1)));
  ctx->owned.push_back(dptr);
  if (count) CU(cudaMemcpyAsync(dptr, host, sizeof(T) * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
  *dev_out = dptr;
  return 0;
}

// A large X in pageable host memory (C3: 40 GB): a plain cudaMemcpy2D stages it through the driver's own bounce buffers at
// ~10 GB/s (4.2 s at C3).  Here: three page-locked staging buffers of 64 MB, filled by a few host threads (memcpy of
// whole rows, already in the device's padded pitch) while the previous ones are in flight on the stream.
int upload_x_staged(boomgpu_ctx *ctx, double *dX, int64_t ldd, int64_t n, int p, const double *X, int64_t ldx) {
  constexpr int kBufs = 3;
  const size_t row_bytes = sizeof(double) * (size_t)p;
  const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)(((size_t)64 << 20) / row_bytes));
  double *stage[kBufs] = {nullptr, nullptr, nullptr};
  cudaEvent_t done[kBufs] = {nullptr, nullptr, nullptr};
  bool in_flight[kBufs] = {false, false, false};
  int rc = 0;
  for (int b = 0; b < kBufs && !rc; ++b) {
    if (cudaMallocHost((void **)&stage[b], row_bytes * (size_t)chunk_rows) != cudaSuccess ||
        cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming) != cudaSuccess)
      rc = fail(ctx, BOOMGPU_ERR_CUDA, "staging buffers for the upload: %s", cudaGetErrorString(cudaGetLastError()));
  }
  const unsigned hw = std::thread::hardware_concurrency();
  const int nthreads = (int)std::max(1u, std::min(8u, hw ? hw : 1u));
  int b = 0;
  for (int64_t r0 = 0; r0 < n && !rc; r0 += chunk_rows, b = (b + 1) % kBufs) {
    const int64_t rows = std::min(chunk_rows, n - r0);
    if (in_flight[b] && cudaEventSynchronize(done[b]) != cudaSuccess) { rc = fail(ctx, BOOMGPU_ERR_CUDA, "upload: event wait failed"); break; }
    double *dst = stage[b];
    auto copy_rows = [=](int64_t a, int64_t z) {
      if (ldx == p) { memcpy(dst + (size_t)a * p, X + (size_t)(r0 + a) * ldx, row_bytes * (size_t)(z - a)); return; }
      for (int64_t i = a; i < z; ++i) memcpy(dst + (size_t)i * p, X + (size_t)(r0 + i) * ldx, row_bytes);
    };
    if (nthreads == 1 || rows < 4 * nthreads) {
      copy_rows(0, rows);
    } else {
      std::vector<std::thread> pool;
      const int64_t per = (rows + nthreads - 1) / nthreads;
      for (int t = 0; t < nthreads; ++t) {
        const int64_t a = t * per, z = std::min(rows, a + per);
        if (a < z) pool.emplace_back(copy_rows, a, z);
      }
      for (auto &th : pool) th.join();
    }
    if (cudaMemcpy2DAsync(dX + (size_t)r0 * ldd, sizeof(double) * ldd, dst, row_bytes, row_bytes, (size_t)rows, cudaMemcpyHostToDevice,
                          ctx->stream) != cudaSuccess ||
        cudaEventRecord(done[b], ctx->stream) != cudaSuccess)
      rc = fail(ctx, BOOMGPU_ERR_CUDA, "upload: %s", cudaGetErrorString(cudaGetLastError()));
    in_flight[b] = true;
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc) rc = fail(ctx, BOOMGPU_ERR_CUDA, "upload: %s", cudaGetErrorString(cudaGetLastError()));
  for (int i = 0; i < kBufs; ++i) {
    if (done[i]) cudaEventDestroy(done[i]);
    if (stage[i]) cudaFreeHost(stage[i]);
  }
  return rc;
}

int upload_x(boomgpu_ctx *ctx, int64_t n, int p, const double *X, int64_t ldx) {
  // device layout: row major, leading dimension rounded up to 8 doubles (64 B), pad columns zero
  const int64_t ldd = ((int64_t)p + 7) / 8 * 8;
  double *dX = nullptr;
  CU(cudaMalloc((void **)&dX, sizeof(double) * (size_t)std::max<int64_t>(n * ldd, 1)));
  ctx->owned.push_back(dX);
  if (n) {
    if (ldd != p) CU(cudaMemsetAsync(dX, 0, sizeof(double) * (size_t)(n * ldd), ctx->stream));
    const size_t bytes = sizeof(double) * (size_t)n * (size_t)p;
    if (bytes < ((size_t)256 << 20) || is_pinned_host(X)) {
      CU(cudaMemcpy2DAsync(dX, sizeof(double) * ldd, X, sizeof(double) * ldx, sizeof(double) * p, (size_t)n,
                           cudaMemcpyHostToDevice, ctx->stream));
    } else if (int rc = upload_x_staged(ctx, dX, ldd, n, p, X, ldx)) {
      return rc;
    }
  }
  ctx->X = dX; ctx->ldx = ldd; ctx->n = n; ctx->p = p;
  return 0;
}

int check_dims(boomgpu_ctx *ctx, int64_t n, int p, int64_t ldx, const void *X) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (n < 0 || p <= 0 || ldx < p || (!X && n > 0)) return fail(ctx, BOOMGPU_ERR_ARG, "bad dimensions n=%lld p=%d ldx=%lld", (long long)n, p, (long long)ldx);
  if (p > 16384) return fail(ctx, BOOMGPU_ERR_ARG, "p = %d exceeds the supported maximum 16384", p);
  return 0;
}

// log likelihood + gradient + Hessian in ONE pass of the step kernels (MODEL = kLogitLL / kPoissonLL):
// gradient = X's with s = dl/d eta, Hessian = -X'WX with w = -d2l/d eta2, log likelihood in the y'Wy slot.
template <int MODEL>
int loglike_derivs_impl(boomgpu_ctx *ctx, int model, const double *beta, double log_alpha, double *loglike, double *gradient,
                               double *hessian) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (ctx->model != model || !ctx->X) return fail(ctx, BOOMGPU_ERR_STATE, "no matching data uploaded to this context");
  if (!beta || !loglike) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  if (int rc = ensure_suf(ctx)) return rc;
  DrawParams prm = make_prm(ctx, 0, 0, 0);
  prm.log_alpha = log_alpha;
  RowOut out{nullptr, nullptr, nullptr, nullptr};
  // rows sharded over ranks: every rank must see the likelihood of ALL rows (the mode finder and the MH moves built on it
  // run replicated on every rank and must stay in step), so the packed result is all-reduced like a Gibbs step's statistics
  const bool sharded = ctx->comm && ctx->comm_ranks > 1;
  if (int rc = run_step<MODEL>(ctx, beta, prm, out, nullptr, nullptr, ctx->suf_dev, sharded ? nullptr : ctx->suf_pin)) return rc;
  const int p = ctx->p;
  if (int rc = allreduce_on_stream(ctx, ctx->suf_dev, boomgpu_suf_len(p))) return rc;
  if (int rc = fetch_suf(ctx)) return rc;
  *loglike = ctx->n == 0 ? 0.0 : ctx->suf_pin[(size_t)p * p + p + 1];
  if (gradient) memcpy(gradient, ctx->suf_pin + (size_t)p * p, sizeof(double) * p);
  if (hessian) for (size_t e = 0; e < (size_t)p * p; ++e) hessian[e] = -ctx->suf_pin[e];
  return 0;
}

template <int MODEL>
int loglike_derivs_device_impl(boomgpu_ctx *ctx, int model, const double *beta, double log_alpha, double *suf_dev) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (ctx->model != model || !ctx->X) return fail(ctx, BOOMGPU_ERR_STATE, "no matching data uploaded to this context");
  if (!beta || !suf_dev) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  DrawParams prm = make_prm(ctx, 0, 0, 0);
  prm.log_alpha = log_alpha;
  RowOut out{nullptr, nullptr, nullptr, nullptr};
  return run_step<MODEL>(ctx, beta, prm, out, nullptr, nullptr, suf_dev);
}

// Runs f with the context looking at X_gamma (n x k) instead of X, then restores it.
template <int MODEL>
int step_active_impl(boomgpu_ctx *ctx, int model, const double *beta, int clt_threshold, uint64_t seed, uint64_t iteration,
                            const int32_t *active, int k, double *G, double *diag, double *xty, double scalars[4],
                            double t_sigma = 0.0, double t_nu = 0.0) {
  if (int rc = check_ready(ctx, model)) return rc;
  if (!beta || !active || k < 1 || !G || !diag || !xty) return fail(ctx, BOOMGPU_ERR_ARG, "null / empty argument");
  if (ctx->p <= 64) return fail(ctx, BOOMGPU_ERR_ARG, "the active-set step is for p > 64 (p = %d runs the single-pass kernel)", ctx->p);
  if (k > 128) return fail(ctx, BOOMGPU_ERR_ARG, "active set of %d columns exceeds the supported 128", k);
  DeviceGuard g(ctx->device);
  if (int rc = boomgpu_select_columns(ctx, active, k)) return rc;
  if (int rc = ensure_suf(ctx)) return rc;
  const int p = ctx->p;
  const int64_t cap = (int64_t)p * 128 + 2 * (int64_t)p + 4;
  if (ensure(ctx, &ctx->act_dev, &ctx->act_cap, cap)) return BOOMGPU_ERR_CUDA;
  if (ctx->act_pin_cap < cap) {
    if (ctx->act_pin) { CU(cudaFreeHost(ctx->act_pin)); ctx->act_pin = nullptr; }
    CU(cudaMallocHost((void **)&ctx->act_pin, sizeof(double) * (size_t)cap));
    ctx->act_pin_cap = cap;
  }
  RowOut out{nullptr, nullptr, nullptr, nullptr};
  const int path_save = ctx->path;
  ctx->path = 2; ctx->stats_mode = 1;
  DrawParams prm = make_prm(ctx, clt_threshold, seed, iteration);
  if (MODEL == kStudentT) { prm.t_inv_sigma = 1.0 / t_sigma; prm.t_nu = t_nu; }
  int rc = run_step<MODEL>(ctx, beta, prm, out, nullptr, nullptr, ctx->suf_dev);
  ctx->path = path_save; ctx->stats_mode = 0;
  if (rc) return rc;
  int ka8 = 0;
  if ((rc = launch_panel(ctx, ctx->act_dev, &ka8))) return rc;
  const int64_t len = (int64_t)p * ka8 + 2 * (int64_t)p;
  // the four scalars of the imputer pass sit at suf_dev[p*p + p ..]: bring them behind the panel result
  CU(cudaMemcpyAsync(ctx->act_dev + len, ctx->suf_dev + (int64_t)p * p + p, sizeof(double) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  if ((rc = allreduce_on_stream(ctx, ctx->act_dev, len + 4))) return rc;
  CU(cudaMemcpyAsync(ctx->act_pin, ctx->act_dev, sizeof(double) * (size_t)(len + 4), cudaMemcpyDeviceToHost, ctx->stream));
  if ((rc = finish_and_check(ctx))) return rc;
  for (int j = 0; j < p; ++j) memcpy(G + (size_t)j * k, ctx->act_pin + (size_t)j * ka8, sizeof(double) * k);
  memcpy(diag, ctx->act_pin + (size_t)p * ka8, sizeof(double) * p);
  memcpy(xty, ctx->act_pin + (size_t)p * ka8 + p, sizeof(double) * p);
  if (scalars) memcpy(scalars, ctx->act_pin + len, sizeof(double) * 4);
  return 0;
}

template <class F>
int with_selected_columns(boomgpu_ctx *ctx, F f) {
  if (!ctx->Xsel || ctx->sel_src != ctx->X || ctx->sel_cols.empty())
    return fail(ctx, BOOMGPU_ERR_STATE, "no columns selected for the current data (boomgpu_select_columns)");
  const double *X = ctx->X, *Xt = ctx->Xt; const int64_t ldx = ctx->ldx, ldxt = ctx->ldxt; const int p = ctx->p;
  double *Xt_owned = ctx->Xt_owned;
  ctx->X = ctx->Xsel; ctx->ldx = ctx->sel_ld; ctx->p = (int)ctx->sel_cols.size();
  ctx->Xt = nullptr; ctx->ldxt = 0; ctx->Xt_owned = nullptr;    // the selected matrix is padded: TMA describes it in place
  const int rc = f();
  ctx->X = X; ctx->ldx = ldx; ctx->p = p; ctx->Xt = Xt; ctx->ldxt = ldxt; ctx->Xt_owned = Xt_owned;
  return rc;
}

}  // namespace

// =============================================================================================
extern "C" {

const char *boomgpu_version(void) { return "boomgpu 0.1 (sm_100a)"; }

int boomgpu_create(boomgpu_ctx **out, int device) {
  boomgpu_ctx *ctx = nullptr;
  if (!out) return fail(nullptr, BOOMGPU_ERR_ARG, "null ctx pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, BOOMGPU_ERR_CUDA, "no CUDA device available (%s); boomgpu has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(nullptr, BOOMGPU_ERR_ARG, "device %d out of range (%d devices)", device, count);
  ctx = new boomgpu_ctx;
  ctx->device = device;
  DeviceGuard g(device);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, BOOMGPU_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  }
  if (prop.major < 9) {
    delete ctx;
    return fail(nullptr, BOOMGPU_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  }
  ctx->sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc((void **)&ctx->err_dev, sizeof(int)) != cudaSuccess ||
      cudaMallocHost((void **)&ctx->err_pin, sizeof(int)) != cudaSuccess ||
      cudaMalloc((void **)&ctx->tail_counter, sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(ctx->tail_counter, 0, sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(ctx->err_dev, 0, sizeof(int)) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, BOOMGPU_ERR_CUDA, "context set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  ctx->own_stream = true;
  *out = ctx;
  return 0;
}

void boomgpu_destroy(boomgpu_ctx *ctx) {
  if (!ctx) return;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  drain_timings(ctx);
  for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
  if (ctx->comm) nccl().CommDestroy(ctx->comm);
  free_data(ctx);
  for (void *q : ctx->tab_owned) cudaFree(q);
  cudaFree(ctx->mix_dev);
  cudaFree(ctx->beta_dev); cudaFreeHost(ctx->beta_pin);
  cudaFree(ctx->suf_dev); cudaFreeHost(ctx->suf_pin);
  cudaFree(ctx->partials); cudaFree(ctx->scal_partials); cudaFree(ctx->resid_buf); cudaFree(ctx->syrk_items);
  cudaFree(ctx->w_buf); cudaFree(ctx->s_buf);
  cudaFree(ctx->Xsel); cudaFree(ctx->sel_cols_dev);
  cudaFree(ctx->act_dev); cudaFreeHost(ctx->act_pin); cudaFree(ctx->col_buf);
  cudaFree(ctx->err_dev); cudaFreeHost(ctx->err_pin); cudaFree(ctx->tail_counter);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *boomgpu_last_error(const boomgpu_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int boomgpu_set_stream(boomgpu_ctx *ctx, void *cuda_stream) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  DeviceGuard g(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) { CU(cudaStreamDestroy(ctx->stream)); ctx->own_stream = false; }
  ctx->stream = (cudaStream_t)cuda_stream;
  return 0;
}

int boomgpu_set_row_offset(boomgpu_ctx *ctx, uint64_t first_global_row) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  ctx->row_offset = first_global_row;
  return 0;
}

int boomgpu_set_option(boomgpu_ctx *ctx, const char *name, int64_t value) {
  if (!ctx || !name) return BOOMGPU_ERR_ARG;
  if (!strcmp(name, "path")) {
    if (value < 0 || value > 2) return fail(ctx, BOOMGPU_ERR_ARG, "path must be 0, 1 or 2");
    ctx->path = (int)value;
    return 0;
  }
  if (!strcmp(name, "timing")) { ctx->timing = value != 0; return 0; }
  if (!strcmp(name, "single_launch")) { ctx->single_launch = value != 0; return 0; }
  if (!strcmp(name, "syrk_diag")) { ctx->syrk_diag = value != 0; return 0; }
  if (!strcmp(name, "syrk_filter")) { ctx->syrk_filter = (int)value; return 0; }
  if (!strcmp(name, "syrk_rdiag")) { ctx->syrk_rdiag = value != 0; return 0; }
  if (!strcmp(name, "tma_promotion")) { ctx->tma_promotion = value < 0 || value > 3 ? 3 : (int)value; return 0; }
  if (!strcmp(name, "syrk_order")) { ctx->syrk_order = value < 0 || value > 2 ? 1 : (int)value; return 0; }
  if (!strcmp(name, "syrk_cluster")) {
    if (value < 0 || value > 16) return fail(ctx, BOOMGPU_ERR_ARG, "syrk_cluster must be in 0..16");
    ctx->syrk_cluster = (int)value; return 0;
  }
  if (!strcmp(name, "syrk_waves")) {
    if (value < 1 || value > 256) return fail(ctx, BOOMGPU_ERR_ARG, "syrk_waves must be in 1..256");
    ctx->syrk_waves = (int)value; return 0;
  }
  if (!strcmp(name, "gather")) {
    if (value < 0 || value > 2) return fail(ctx, BOOMGPU_ERR_ARG, "gather must be 0 (auto), 1 (never) or 2 (whenever beta has a zero)");
    ctx->gather = (int)value;
    return 0;
  }
  if (!strcmp(name, "small_variant")) {
    if (value < 0 || value > 4) return fail(ctx, BOOMGPU_ERR_ARG, "small_variant must be 0 .. 4");
    ctx->small_variant = (int)value;
    return 0;
  }
  return fail(ctx, BOOMGPU_ERR_ARG, "unknown option '%s'", name);
}

int boomgpu_upload_binomial(boomgpu_ctx *ctx, int64_t n, int p, const double *X, int64_t ldx, const double *y,
                            const double *ntrials) {
  if (int rc = check_dims(ctx, n, p, ldx, X)) return rc;
  if (n > 0 && (!y || !ntrials)) return fail(ctx, BOOMGPU_ERR_ARG, "null y / ntrials");
  DeviceGuard g(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  free_data(ctx);
  if (int rc = upload_x(ctx, n, p, X, ldx)) return rc;
  if (int rc = upload_array(ctx, y, n, &ctx->y)) return rc;
  if (int rc = upload_array(ctx, ntrials, n, &ctx->ntrials)) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->model = kLogit;
  return 0;
}

int boomgpu_upload_poisson(boomgpu_ctx *ctx, int64_t n, int p, const double *X, int64_t ldx, const int64_t *y,
                           const double *exposure) {
  if (int rc = check_dims(ctx, n, p, ldx, X)) return rc;
  if (n > 0 && (!y || !exposure)) return fail(ctx, BOOMGPU_ERR_ARG, "null y / exposure");
  DeviceGuard g(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  free_data(ctx);
  if (int rc = upload_x(ctx, n, p, X, ldx)) return rc;
  if (int rc = upload_array(ctx, y, n, &ctx->yi)) return rc;
  if (int rc = upload_array(ctx, exposure, n, &ctx->exposure)) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->model = kPoisson;
  return 0;
}

// Chunked upload: the caller packs a few thousand rows at a time out of its own representation (BOOM keeps one heap object
// per observation) instead of materialising a second n x p copy on the host.
int boomgpu_upload_begin(boomgpu_ctx *ctx, int poisson, int64_t n, int p) {
  if (int rc = check_dims(ctx, n, p, p, n > 0 ? (const void *)ctx : nullptr)) return rc;
  DeviceGuard g(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  free_data(ctx);
  const int64_t ldd = ((int64_t)p + 7) / 8 * 8;
  double *dX = nullptr, *daux = nullptr;
  void *dy = nullptr;
  const size_t n1 = (size_t)std::max<int64_t>(n, 1);
  CU(cudaMalloc((void **)&dX, sizeof(double) * n1 * ldd));
  ctx->owned.push_back(dX);
  CU(cudaMalloc(&dy, 8 * n1));
  ctx->owned.push_back(dy);
  CU(cudaMalloc((void **)&daux, sizeof(double) * n1));
  ctx->owned.push_back(daux);
  if (ldd != p && n) CU(cudaMemsetAsync(dX, 0, sizeof(double) * (size_t)(n * ldd), ctx->stream));
  ctx->X = dX; ctx->ldx = ldd; ctx->n = n; ctx->p = p;
  if (poisson == 1) { ctx->yi = (const int64_t *)dy; ctx->exposure = daux; }
  else { ctx->y = (const double *)dy; ctx->ntrials = daux; }
  ctx->model = -1;                 // not usable until boomgpu_upload_end
  ctx->upload_kind = poisson == 1 ? kPoisson : poisson == 2 ? kStudentT : kLogit;   // 2: plain regression rows (y only)
  return 0;
}

int boomgpu_upload_rows(boomgpu_ctx *ctx, int64_t row0, int64_t nrows, const double *X, int64_t ldx, const void *y, const double *aux) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (ctx->upload_kind < 0 || !ctx->X) return fail(ctx, BOOMGPU_ERR_STATE, "boomgpu_upload_rows without boomgpu_upload_begin");
  if (row0 < 0 || nrows < 0 || row0 + nrows > ctx->n || ldx < ctx->p ||
      (nrows > 0 && (!X || !y || (!aux && ctx->upload_kind != kStudentT))))
    return fail(ctx, BOOMGPU_ERR_ARG, "bad row range [%lld, %lld) of %lld", (long long)row0, (long long)(row0 + nrows), (long long)ctx->n);
  if (nrows == 0) return 0;
  DeviceGuard g(ctx->device);
  double *dX = const_cast<double *>(ctx->X) + row0 * ctx->ldx;
  CU(cudaMemcpy2DAsync(dX, sizeof(double) * ctx->ldx, X, sizeof(double) * ldx, sizeof(double) * ctx->p, (size_t)nrows,
                       cudaMemcpyHostToDevice, ctx->stream));
  void *dy = ctx->upload_kind == kPoisson ? (void *)(const_cast<int64_t *>(ctx->yi) + row0) : (void *)(const_cast<double *>(ctx->y) + row0);
  double *da = const_cast<double *>(ctx->upload_kind == kPoisson ? ctx->exposure : ctx->ntrials) + row0;
  CU(cudaMemcpyAsync(dy, y, 8 * (size_t)nrows, cudaMemcpyHostToDevice, ctx->stream));
  if (aux) CU(cudaMemcpyAsync(da, aux, sizeof(double) * (size_t)nrows, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));   // the caller re-uses its chunk buffers
  return 0;
}

int boomgpu_upload_end(boomgpu_ctx *ctx) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (ctx->upload_kind < 0 || !ctx->X) return fail(ctx, BOOMGPU_ERR_STATE, "boomgpu_upload_end without boomgpu_upload_begin");
  DeviceGuard g(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->model = ctx->upload_kind;
  ctx->upload_kind = -1;
  return 0;
}

int boomgpu_adopt_binomial(boomgpu_ctx *ctx, int64_t n, int p, const double *dX, int64_t ldx, const double *dy,
                           const double *dntrials) {
  if (int rc = check_dims(ctx, n, p, ldx, dX)) return rc;
  DeviceGuard g(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  free_data(ctx);
  ctx->X = dX; ctx->ldx = ldx; ctx->n = n; ctx->p = p; ctx->y = dy; ctx->ntrials = dntrials;
  ctx->model = kLogit;
  return 0;
}

int boomgpu_adopt_poisson(boomgpu_ctx *ctx, int64_t n, int p, const double *dX, int64_t ldx, const int64_t *dy,
                          const double *dexposure) {
  if (int rc = check_dims(ctx, n, p, ldx, dX)) return rc;
  DeviceGuard g(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  free_data(ctx);
  ctx->X = dX; ctx->ldx = ldx; ctx->n = n; ctx->p = p; ctx->yi = dy; ctx->exposure = dexposure;
  ctx->model = kPoisson;
  return 0;
}

// plain regression rows (y double, no trials / exposure): the Student-t sibling (TRegressionModel, Models/Glm/TRegression.cpp:43-54)
int boomgpu_upload_regression(boomgpu_ctx *ctx, int64_t n, int p, const double *X, int64_t ldx, const double *y) {
  if (int rc = check_dims(ctx, n, p, ldx, X)) return rc;
  if (n > 0 && !y) return fail(ctx, BOOMGPU_ERR_ARG, "null y");
  DeviceGuard g(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  free_data(ctx);
  if (int rc = upload_x(ctx, n, p, X, ldx)) return rc;
  if (int rc = upload_array(ctx, y, n, &ctx->y)) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->model = kStudentT;
  return 0;
}

int boomgpu_adopt_regression(boomgpu_ctx *ctx, int64_t n, int p, const double *dX, int64_t ldx, const double *dy) {
  if (int rc = check_dims(ctx, n, p, ldx, dX)) return rc;
  DeviceGuard g(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  free_data(ctx);
  ctx->X = dX; ctx->ldx = ldx; ctx->n = n; ctx->p = p; ctx->y = dy;
  ctx->model = kStudentT;
  return 0;
}

int boomgpu_comm_unique_id(char id[BOOMGPU_COMM_ID_BYTES]) {
  if (!id) return fail(nullptr, BOOMGPU_ERR_ARG, "null id");
  if (!nccl().ok) return fail(nullptr, BOOMGPU_ERR_STATE, "%s", nccl().why.c_str());
  NcclUniqueId u;
  const int rc = nccl().GetUniqueId(&u);
  if (rc) return fail(nullptr, BOOMGPU_ERR_CUDA, "ncclGetUniqueId failed: %s", nccl().GetErrorString(rc));
  memcpy(id, u.internal, BOOMGPU_COMM_ID_BYTES);
  return 0;
}

int boomgpu_comm_init(boomgpu_ctx *ctx, const char id[BOOMGPU_COMM_ID_BYTES], int nranks, int rank) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (!id || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, BOOMGPU_ERR_ARG, "bad communicator request (rank %d of %d)", rank, nranks);
  if (!nccl().ok) return fail(ctx, BOOMGPU_ERR_STATE, "%s", nccl().why.c_str());
  DeviceGuard g(ctx->device);
  if (ctx->comm) { nccl().CommDestroy(ctx->comm); ctx->comm = nullptr; ctx->comm_ranks = 1; }
  NcclUniqueId u;
  memcpy(u.internal, id, BOOMGPU_COMM_ID_BYTES);
  const int rc = nccl().CommInitRank(&ctx->comm, nranks, u, rank);
  if (rc) { ctx->comm = nullptr; return fail(ctx, BOOMGPU_ERR_CUDA, "ncclCommInitRank failed: %s", nccl().GetErrorString(rc)); }
  ctx->comm_ranks = nranks;
  return 0;
}

int boomgpu_comm_destroy(boomgpu_ctx *ctx) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  DeviceGuard g(ctx->device);
  if (ctx->comm) {
    CU(cudaStreamSynchronize(ctx->stream));
    nccl().CommDestroy(ctx->comm);
    ctx->comm = nullptr; ctx->comm_ranks = 1;
  }
  return 0;
}

int boomgpu_allreduce(boomgpu_ctx *ctx, double *dev, int64_t count) {
  if (!ctx || !dev || count < 0) return BOOMGPU_ERR_ARG;
  DeviceGuard g(ctx->device);
  return allreduce_on_stream(ctx, dev, count);
}

int boomgpu_set_logit_mixture(boomgpu_ctx *ctx, int K, const double *mu, const double *sigma, const double *weights) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (K < 1 || K > kMaxLogitK || !sigma || !weights) return fail(ctx, BOOMGPU_ERR_ARG, "mixture needs 1 <= K <= %d", kMaxLogitK);
  LogitMixture m;
  memset(&m, 0, sizeof(m));
  m.K = K;
  for (int k = 0; k < K; ++k) {
    if (!(sigma[k] > 0) || !(weights[k] > 0)) return fail(ctx, BOOMGPU_ERR_ARG, "mixture sigma and weights must be positive");
    m.mu[k] = mu ? mu[k] : 0.0;
    m.sigma[k] = sigma[k];
    m.inv_sigma[k] = 1.0 / sigma[k];
    m.weights[k] = weights[k];
    m.lconst[k] = std::log(weights[k]) - kLnSqrt2Pi - std::log(sigma[k]);
    m.inv_sigsq[k] = 1.0 / (sigma[k] * sigma[k]);
  }
  if (ctx->have_mix && !memcmp(&m, &ctx->mix, sizeof(m))) return 0;   // unchanged (samplers re-state it every draw)
  ctx->mix = m;
  ctx->have_mix = false;
  LogitHot &h = ctx->hot;
  memset(&h, 0, sizeof(h));
  h.K = K;
  h.center = m.mu[0];
  double lmax = -1e300;
  for (int k = 0; k < K; ++k) lmax = std::max(lmax, m.lconst[k] * 1.4426950408889634);
  h.zero_mean = 1; h.wide = 0;
  for (int k = 0; k < K; ++k) {
    h.lconst2[k] = (float)(m.lconst[k] * 1.4426950408889634);
    h.l0[k] = (float)(m.lconst[k] * 1.4426950408889634 - lmax);
    h.hs2[k] = (float)(-0.5 * 1.4426950408889634 * m.inv_sigsq[k]);
    h.mu_c[k] = (float)(m.mu[k] - h.center);
    h.inv_sigsq[k] = m.inv_sigsq[k];
    h.mu_d[k] = m.mu[k];
    h.logw[k] = std::log(m.inv_sigsq[k]);
    if (m.mu[k] != m.mu[0]) h.zero_mean = 0;
    if (m.sigma[k] > m.sigma[h.wide]) h.wide = k;
  }
  if (h.wide != K - 1) h.zero_mean = 0;   // the scale-mixture fast path takes the LAST component as the widest (the logit table is sorted by sigma)
  DeviceGuard g(ctx->device);
  if (!ctx->mix_dev) CU(cudaMalloc((void **)&ctx->mix_dev, sizeof(LogitMixtureDev)));
  CU(cudaStreamSynchronize(ctx->stream));   // no step may still be reading the old mixture
  LogitMixtureDev both;
  both.full = m; both.hot = h;
  CU(cudaMemcpy(ctx->mix_dev, &both, sizeof(both), cudaMemcpyHostToDevice));
  ctx->have_mix = true;
  return 0;
}

int boomgpu_set_poisson_table(boomgpu_ctx *ctx, int ntab, const int64_t *nu, const int32_t *offset, const double *weights,
                              const double *mu, const double *sigma, int64_t gaussian_cutoff) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (ntab < 1 || !nu || !offset || !weights || !mu || !sigma) return fail(ctx, BOOMGPU_ERR_ARG, "bad Poisson table");
  int e1 = -1;
  for (int e = 0; e < ntab; ++e) {
    if (e > 0 && nu[e] < nu[e - 1]) return fail(ctx, BOOMGPU_ERR_ARG, "Poisson table must be sorted by nu");
    const int K = offset[e + 1] - offset[e];
    if (K < 1 || K > kMaxLogitK) return fail(ctx, BOOMGPU_ERR_ARG, "table entry nu=%lld has %d components (max %d)", (long long)nu[e], K, kMaxLogitK);
    if (nu[e] == 1 && e1 < 0) e1 = e;
  }
  if (e1 < 0) return fail(ctx, BOOMGPU_ERR_ARG, "Poisson table has no entry for nu = 1");
  const int total = offset[ntab];
  if (ctx->have_tab && gaussian_cutoff == ctx->tab_cut_h && (size_t)ntab == ctx->tab_nu_h.size() && (size_t)total == ctx->tab_w_h.size() &&
      !memcmp(nu, ctx->tab_nu_h.data(), sizeof(int64_t) * ntab) && !memcmp(offset, ctx->tab_off_h.data(), sizeof(int32_t) * (ntab + 1)) &&
      !memcmp(weights, ctx->tab_w_h.data(), sizeof(double) * total) && !memcmp(mu, ctx->tab_mu_h.data(), sizeof(double) * total) &&
      !memcmp(sigma, ctx->tab_sig_h.data(), sizeof(double) * total))
    return 0;   // unchanged
  std::vector<double> inv_sigma(total), lconst(total);
  std::vector<float> mu_f(total), lconst2_f(total), hs2_f(total);
  std::vector<double> inv_sigsq(total), logw(total);
  for (int i = 0; i < total; ++i) {
    if (!(sigma[i] > 0) || !(weights[i] > 0)) return fail(ctx, BOOMGPU_ERR_ARG, "table sigma and weights must be positive");
    inv_sigma[i] = 1.0 / sigma[i];
    lconst[i] = std::log(weights[i]) - kLnSqrt2Pi - std::log(sigma[i]);
    lconst2_f[i] = (float)(lconst[i] * 1.4426950408889634);
    hs2_f[i] = (float)(-0.5 * 1.4426950408889634 * inv_sigma[i] * inv_sigma[i]);
    inv_sigsq[i] = 1.0 / (sigma[i] * sigma[i]);
    logw[i] = std::log(inv_sigsq[i]);
  }
  // dense index for the small counts, and the nu = 1 entry in kernel-parameter form
  const int dense_n = (int)std::min<int64_t>(nu[ntab - 1] + 1, 1 << 16);
  std::vector<int32_t> dense((size_t)std::max(dense_n, 1), -1);
  for (int e = ntab - 1; e >= 0; --e) if (nu[e] >= 0 && nu[e] < dense_n) dense[(size_t)nu[e]] = e;   // first of duplicates wins
  LogitHot &xh = ctx->ext_hot;
  memset(&xh, 0, sizeof(xh));
  xh.K = offset[e1 + 1] - offset[e1];
  xh.center = mu[offset[e1]];
  for (int k = 0; k < xh.K; ++k) {
    const int i = offset[e1] + k;
    xh.lconst2[k] = lconst2_f[i]; xh.hs2[k] = hs2_f[i]; xh.mu_c[k] = (float)(mu[i] - xh.center);
    xh.inv_sigsq[k] = inv_sigsq[i]; xh.mu_d[k] = mu[i]; xh.logw[k] = logw[i];
  }
  for (int e = 0; e < ntab; ++e)
    for (int i = offset[e]; i < offset[e + 1]; ++i) mu_f[i] = (float)(mu[i] - mu[offset[e]]);
  DeviceGuard g(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  for (void *q : ctx->tab_owned) cudaFree(q);
  ctx->tab_owned.clear();
  ctx->have_tab = false;
  auto up = [&](const void *src, size_t bytes, void **dst) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    ctx->tab_owned.push_back(*dst);
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  PoissonTable &t = ctx->tab;
  t.ntab = ntab; t.gaussian_cutoff = gaussian_cutoff; t.e1 = e1;
  CU(up(nu, sizeof(int64_t) * ntab, (void **)&t.nu));
  CU(up(offset, sizeof(int32_t) * (ntab + 1), (void **)&t.offset));
  CU(up(mu, sizeof(double) * total, (void **)&t.mu));
  CU(up(sigma, sizeof(double) * total, (void **)&t.sigma));
  CU(up(inv_sigma.data(), sizeof(double) * total, (void **)&t.inv_sigma));
  CU(up(lconst.data(), sizeof(double) * total, (void **)&t.lconst));
  CU(up(mu_f.data(), sizeof(float) * total, (void **)&t.mu_f));
  CU(up(lconst2_f.data(), sizeof(float) * total, (void **)&t.lconst2_f));
  CU(up(hs2_f.data(), sizeof(float) * total, (void **)&t.hs2_f));
  CU(up(inv_sigsq.data(), sizeof(double) * total, (void **)&t.inv_sigsq));
  CU(up(logw.data(), sizeof(double) * total, (void **)&t.logw));
  CU(up(dense.data(), sizeof(int32_t) * dense.size(), (void **)&t.dense));
  t.dense_n = dense_n;
  ctx->have_tab = true;
  ctx->tab_nu_h.assign(nu, nu + ntab); ctx->tab_off_h.assign(offset, offset + ntab + 1);
  ctx->tab_w_h.assign(weights, weights + total); ctx->tab_mu_h.assign(mu, mu + total); ctx->tab_sig_h.assign(sigma, sigma + total);
  ctx->tab_cut_h = gaussian_cutoff;
  return 0;
}

int boomgpu_poisson_counts_present(boomgpu_ctx *ctx, unsigned char *present, int64_t len) {
  if (!ctx || !present || len <= 0) return BOOMGPU_ERR_ARG;
  if (ctx->model != kPoisson || (!ctx->yi && ctx->n > 0)) return fail(ctx, BOOMGPU_ERR_STATE, "no Poisson data uploaded to this context");
  DeviceGuard g(ctx->device);
  unsigned char *d = nullptr;
  CU(cudaMalloc((void **)&d, (size_t)len));
  cudaError_t e = cudaMemsetAsync(d, 0, (size_t)len, ctx->stream);
  if (e == cudaSuccess && ctx->n > 0) {
    LaunchScope ls(ctx, 4);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ctx->n + 255) / 256, (int64_t)ctx->sms * 8));
    counts_present_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->yi, ctx->n, d, len);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(present, d, (size_t)len, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); else cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(ctx, BOOMGPU_ERR_CUDA, "counts_present failed: %s", cudaGetErrorString(e));
  return 0;
}

int boomgpu_pin_host(void *ptr, uint64_t bytes) {
  if (!ptr || !bytes) return fail(nullptr, BOOMGPU_ERR_ARG, "boomgpu_pin_host: null range");
  cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, BOOMGPU_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e)); }
  return 0;
}
int boomgpu_unpin_host(void *ptr) {
  if (!ptr) return 0;
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, BOOMGPU_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e)); }
  return 0;
}

int64_t boomgpu_suf_len(int p) { return (int64_t)p * p + p + 4; }

int boomgpu_logit_step_device(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed, uint64_t iteration,
                              double *suf_dev) {
  if (int rc = check_ready(ctx, kLogit)) return rc;
  if (int rc = check_clt(ctx, clt_threshold)) return rc;
  if (!beta || !suf_dev) return fail(ctx, BOOMGPU_ERR_ARG, "null beta / suf_dev");
  DeviceGuard g(ctx->device);
  RowOut out{nullptr, nullptr, nullptr, nullptr};
  return run_step<kLogit>(ctx, beta, make_prm(ctx, clt_threshold, seed, iteration), out, nullptr, nullptr, suf_dev);
}

int boomgpu_poisson_step_device(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *suf_dev) {
  if (int rc = check_ready(ctx, kPoisson)) return rc;
  if (!beta || !suf_dev) return fail(ctx, BOOMGPU_ERR_ARG, "null beta / suf_dev");
  DeviceGuard g(ctx->device);
  RowOut out{nullptr, nullptr, nullptr, nullptr};
  return run_step<kPoisson>(ctx, beta, make_prm(ctx, 0, seed, iteration), out, nullptr, nullptr, suf_dev);
}

int boomgpu_synchronize(boomgpu_ctx *ctx) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  DeviceGuard g(ctx->device);
  return finish_and_check(ctx);
}

int boomgpu_suf_buffer(boomgpu_ctx *ctx, double **suf_dev) {
  if (!ctx || !suf_dev) return BOOMGPU_ERR_ARG;
  if (!ctx->X) return fail(ctx, BOOMGPU_ERR_STATE, "no data uploaded to this context");
  DeviceGuard g(ctx->device);
  if (int rc = ensure_suf(ctx)) return rc;
  *suf_dev = ctx->suf_dev;
  return 0;
}

int boomgpu_download(boomgpu_ctx *ctx, const double *src_dev, double *dst_host, int64_t count) {
  if (!ctx || !src_dev || !dst_host || count < 0) return BOOMGPU_ERR_ARG;
  DeviceGuard g(ctx->device);
  if (ctx->suf_pin_cap < count) {
    if (ctx->suf_pin) { CU(cudaFreeHost(ctx->suf_pin)); ctx->suf_pin = nullptr; }
    CU(cudaMallocHost((void **)&ctx->suf_pin, sizeof(double) * (size_t)count));
    ctx->suf_pin_cap = count;
  }
  CU(cudaMemcpyAsync(ctx->suf_pin, src_dev, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
  if (int rc = finish_and_check(ctx)) return rc;
  memcpy(dst_host, ctx->suf_pin, sizeof(double) * (size_t)count);
  return 0;
}

int boomgpu_logit_step(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed, uint64_t iteration,
                       double *xtx, double *xty, int64_t *sample_size) {
  if (int rc = check_ready(ctx, kLogit)) return rc;
  if (int rc = check_clt(ctx, clt_threshold)) return rc;
  if (!beta || !xtx || !xty) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  if (int rc = ensure_suf(ctx)) return rc;
  const bool sharded = ctx->comm && ctx->comm_ranks > 1;
  RowOut out{nullptr, nullptr, nullptr, nullptr};
  if (int rc = run_step<kLogit>(ctx, beta, make_prm(ctx, clt_threshold, seed, iteration), out, nullptr, nullptr, ctx->suf_dev,
                                sharded ? nullptr : ctx->suf_pin))
    return rc;
  if (int rc = allreduce_on_stream(ctx, ctx->suf_dev, boomgpu_suf_len(ctx->p))) return rc;
  bool in_place = false;
  if (int rc = fetch_suf(ctx, xtx, &in_place)) return rc;
  const int p = ctx->p;
  if (!in_place) memcpy(xtx, ctx->suf_pin, sizeof(double) * (size_t)p * p);
  memcpy(xty, ctx->suf_pin + (size_t)p * p, sizeof(double) * p);
  if (sample_size) *sample_size = (int64_t)llround(ctx->suf_pin[(size_t)p * p + p]);
  return 0;
}

int boomgpu_poisson_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *xtwx, double *xtwy,
                         double scalars[4]) {
  if (int rc = check_ready(ctx, kPoisson)) return rc;
  if (!beta || !xtwx || !xtwy) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  if (int rc = ensure_suf(ctx)) return rc;
  const bool sharded = ctx->comm && ctx->comm_ranks > 1;
  RowOut out{nullptr, nullptr, nullptr, nullptr};
  if (int rc = run_step<kPoisson>(ctx, beta, make_prm(ctx, 0, seed, iteration), out, nullptr, nullptr, ctx->suf_dev,
                                  sharded ? nullptr : ctx->suf_pin))
    return rc;
  if (int rc = allreduce_on_stream(ctx, ctx->suf_dev, boomgpu_suf_len(ctx->p))) return rc;
  bool in_place = false;
  if (int rc = fetch_suf(ctx, xtwx, &in_place)) return rc;
  const int p = ctx->p;
  if (!in_place) memcpy(xtwx, ctx->suf_pin, sizeof(double) * (size_t)p * p);
  memcpy(xtwy, ctx->suf_pin + (size_t)p * p, sizeof(double) * p);
  if (scalars) memcpy(scalars, ctx->suf_pin + (size_t)p * p + p, sizeof(double) * 4);
  return 0;
}

// ---- active-set statistics ----------------------------------------------------------------------------
int boomgpu_logit_step_active(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed, uint64_t iteration,
                              const int32_t *active, int k, double *G, double *diag, double *xty, int64_t *sample_size) {
  if (ctx) if (int rc = check_clt(ctx, clt_threshold)) return rc;
  double sc[4] = {0, 0, 0, 0};
  const int rc = step_active_impl<kLogit>(ctx, kLogit, beta, clt_threshold, seed, iteration, active, k, G, diag, xty, sc);
  if (!rc && sample_size) *sample_size = (int64_t)llround(sc[0]);
  return rc;
}
int boomgpu_poisson_step_active(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, const int32_t *active, int k,
                                double *G, double *diag, double *xty, double scalars[4]) {
  return step_active_impl<kPoisson>(ctx, kPoisson, beta, 0, seed, iteration, active, k, G, diag, xty, scalars);
}

int boomgpu_weighted_column(boomgpu_ctx *ctx, int j, double *column) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (!ctx->X || !ctx->latents_valid) return fail(ctx, BOOMGPU_ERR_STATE, "no latents on the device: run a two-pass / active-set step first");
  if (j < 0 || j >= ctx->p || !column) return fail(ctx, BOOMGPU_ERR_ARG, "bad column request");
  DeviceGuard g(ctx->device);
  if (ensure(ctx, &ctx->col_buf, &ctx->col_cap, std::max<int64_t>(ctx->n, 1))) return BOOMGPU_ERR_CUDA;
  const int p = ctx->p;
  if (ensure(ctx, &ctx->act_dev, &ctx->act_cap, (int64_t)p * 128 + 2 * (int64_t)p + 4)) return BOOMGPU_ERR_CUDA;
  if (ctx->act_pin_cap < p) {
    if (ctx->act_pin) { CU(cudaFreeHost(ctx->act_pin)); ctx->act_pin = nullptr; }
    CU(cudaMallocHost((void **)&ctx->act_pin, sizeof(double) * ((size_t)p * 128 + 2 * (size_t)p + 4)));
    ctx->act_pin_cap = (int64_t)p * 128 + 2 * (int64_t)p + 4;
  }
  {
    LaunchScope ls(ctx, 4);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ctx->n + 255) / 256, (int64_t)ctx->sms * 8));
    weight_column_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->Xt, ctx->ldxt, ctx->n, j, ctx->w_buf, ctx->col_buf);
  }
  CU(cudaGetLastError());
  if (int rc = launch_xts_vec(ctx, ctx->col_buf, ctx->act_dev)) return rc;
  if (int rc = allreduce_on_stream(ctx, ctx->act_dev, p)) return rc;
  CU(cudaMemcpyAsync(ctx->act_pin, ctx->act_dev, sizeof(double) * (size_t)p, cudaMemcpyDeviceToHost, ctx->stream));
  if (int rc = finish_and_check(ctx)) return rc;
  memcpy(column, ctx->act_pin, sizeof(double) * p);
  return 0;
}

int boomgpu_full_statistics(boomgpu_ctx *ctx, double *xtx, double *xty) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (!ctx->X || !ctx->latents_valid) return fail(ctx, BOOMGPU_ERR_STATE, "no latents on the device: run a two-pass / active-set step first");
  if (!xtx || !xty) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  if (int rc = ensure_suf(ctx)) return rc;
  ctx->host_out_written = false;
  if (int rc = launch_syrk(ctx, ctx->suf_dev)) return rc;
  const int p = ctx->p;
  if (int rc = allreduce_on_stream(ctx, ctx->suf_dev, (int64_t)p * p + p)) return rc;
  bool in_place = false;
  if (int rc = fetch_suf(ctx, xtx, &in_place)) return rc;
  if (!in_place) memcpy(xtx, ctx->suf_pin, sizeof(double) * (size_t)p * p);
  memcpy(xty, ctx->suf_pin + (size_t)p * p, sizeof(double) * p);
  return 0;
}

// ---- probit sibling (BinomialProbitSpikeSlabSampler) ---------------------------------------------------
static int check_probit(boomgpu_ctx *ctx, int clt_threshold) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (ctx->model != kLogit || !ctx->X) return fail(ctx, BOOMGPU_ERR_STATE, "no binomial data uploaded to this context");
  return check_clt(ctx, clt_threshold);
}

int boomgpu_probit_step_device(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed, uint64_t iteration,
                               double *suf_dev, int xty_only) {
  if (int rc = check_probit(ctx, clt_threshold)) return rc;
  if (!beta || !suf_dev) return fail(ctx, BOOMGPU_ERR_ARG, "null beta / suf_dev");
  DeviceGuard g(ctx->device);
  RowOut out{nullptr, nullptr, nullptr, nullptr};
  ctx->xty_only = xty_only != 0;
  const int rc = run_step<kProbit>(ctx, beta, make_prm(ctx, clt_threshold, seed, iteration), out, nullptr, nullptr, suf_dev);
  ctx->xty_only = false;
  return rc;
}

int boomgpu_probit_step(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed, uint64_t iteration, double *xtx,
                        double *xtz, int64_t *sample_size) {
  if (int rc = check_probit(ctx, clt_threshold)) return rc;
  if (!beta || !xtz) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  if (int rc = ensure_suf(ctx)) return rc;
  const bool sharded = ctx->comm && ctx->comm_ranks > 1;
  RowOut out{nullptr, nullptr, nullptr, nullptr};
  ctx->xty_only = xtx == nullptr;
  int rc = run_step<kProbit>(ctx, beta, make_prm(ctx, clt_threshold, seed, iteration), out, nullptr, nullptr, ctx->suf_dev,
                             sharded ? nullptr : ctx->suf_pin);
  ctx->xty_only = false;
  if (rc) return rc;
  const int p = ctx->p;
  const size_t mat = (size_t)p * p;
  if (xtx) {
    if ((rc = allreduce_on_stream(ctx, ctx->suf_dev, boomgpu_suf_len(p)))) return rc;
    bool in_place = false;
    if ((rc = fetch_suf(ctx, xtx, &in_place))) return rc;
    if (!in_place) memcpy(xtx, ctx->suf_pin, sizeof(double) * mat);
  } else {   // X'z and the scalars only: p + 4 doubles travel
    if ((rc = allreduce_on_stream(ctx, ctx->suf_dev + mat, p + 4))) return rc;
    if (!ctx->host_out_written)
      CU(cudaMemcpyAsync(ctx->suf_pin + mat, ctx->suf_dev + mat, sizeof(double) * (size_t)(p + 4), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->host_out_written) { if ((rc = fetch_suf(ctx))) return rc; }
    else if ((rc = finish_and_check(ctx))) return rc;
  }
  memcpy(xtz, ctx->suf_pin + mat, sizeof(double) * p);
  if (sample_size) *sample_size = (int64_t)llround(ctx->suf_pin[mat + p]);
  return 0;
}

int boomgpu_probit_draw(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed, uint64_t iteration, double *sum_z_out) {
  if (int rc = check_probit(ctx, clt_threshold)) return rc;
  if (!beta || !sum_z_out) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  if (int rc = ensure_suf(ctx)) return rc;
  double *ds = nullptr;
  CU(cudaMalloc((void **)&ds, sizeof(double) * (size_t)std::max<int64_t>(ctx->n, 1)));
  RowOut out{nullptr, ds, nullptr, nullptr};
  int rc = run_step<kProbit>(ctx, beta, make_prm(ctx, clt_threshold, seed, iteration), out, nullptr, nullptr, ctx->suf_dev);
  if (!rc && cudaMemcpyAsync(sum_z_out, ds, sizeof(double) * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    rc = fail(ctx, BOOMGPU_ERR_CUDA, "copy of draws failed");
  if (!rc) rc = finish_and_check(ctx); else cudaStreamSynchronize(ctx->stream);
  cudaFree(ds);
  return rc;
}

// ---- Student-t sibling (TRegressionSampler) ------------------------------------------------------------
static int check_student(boomgpu_ctx *ctx, double sigma, double nu) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (ctx->model != kStudentT || !ctx->X) return fail(ctx, BOOMGPU_ERR_STATE, "no regression data uploaded to this context");
  if (!(sigma > 0) || !(nu > 0) || !std::isfinite(sigma) || !std::isfinite(nu))
    return fail(ctx, BOOMGPU_ERR_ARG, "sigma = %g, nu = %g: both must be positive and finite", sigma, nu);
  return 0;
}

static DrawParams make_student_prm(boomgpu_ctx *ctx, double sigma, double nu, uint64_t seed, uint64_t iteration) {
  DrawParams prm = make_prm(ctx, 0, seed, iteration);
  prm.t_inv_sigma = 1.0 / sigma; prm.t_nu = nu;
  return prm;
}

int boomgpu_student_step_device(boomgpu_ctx *ctx, const double *beta, double sigma, double nu, uint64_t seed, uint64_t iteration,
                                double *suf_dev) {
  if (int rc = check_student(ctx, sigma, nu)) return rc;
  if (!beta || !suf_dev) return fail(ctx, BOOMGPU_ERR_ARG, "null beta / suf_dev");
  DeviceGuard g(ctx->device);
  RowOut out{nullptr, nullptr, nullptr, nullptr};
  return run_step<kStudentT>(ctx, beta, make_student_prm(ctx, sigma, nu, seed, iteration), out, nullptr, nullptr, suf_dev);
}

int boomgpu_student_step(boomgpu_ctx *ctx, const double *beta, double sigma, double nu, uint64_t seed, uint64_t iteration,
                         double *xtwx, double *xtwy, double scalars[4]) {
  if (int rc = check_student(ctx, sigma, nu)) return rc;
  if (!beta || !xtwx || !xtwy) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  if (int rc = ensure_suf(ctx)) return rc;
  const bool sharded = ctx->comm && ctx->comm_ranks > 1;
  RowOut out{nullptr, nullptr, nullptr, nullptr};
  if (int rc = run_step<kStudentT>(ctx, beta, make_student_prm(ctx, sigma, nu, seed, iteration), out, nullptr, nullptr, ctx->suf_dev,
                                   sharded ? nullptr : ctx->suf_pin))
    return rc;
  if (int rc = allreduce_on_stream(ctx, ctx->suf_dev, boomgpu_suf_len(ctx->p))) return rc;
  bool in_place = false;
  if (int rc = fetch_suf(ctx, xtwx, &in_place)) return rc;
  const int p = ctx->p;
  if (!in_place) memcpy(xtwx, ctx->suf_pin, sizeof(double) * (size_t)p * p);
  memcpy(xtwy, ctx->suf_pin + (size_t)p * p, sizeof(double) * p);
  if (scalars) memcpy(scalars, ctx->suf_pin + (size_t)p * p + p, sizeof(double) * 4);
  return 0;
}

// active-set form of the Student-t step (as boomgpu_logit_step_active): G = X'WX[:, active], the diagonal and X'Wy for the weights
// drawn at (beta, sigma, nu); the weights stay on the device for boomgpu_weighted_column / boomgpu_full_statistics
int boomgpu_student_step_active(boomgpu_ctx *ctx, const double *beta, double sigma, double nu, uint64_t seed, uint64_t iteration,
                                const int32_t *active, int k, double *G, double *diag, double *xty, double scalars[4]) {
  if (int rc = check_student(ctx, sigma, nu)) return rc;
  return step_active_impl<kStudentT>(ctx, kStudentT, beta, 0, seed, iteration, active, k, G, diag, xty, scalars, sigma, nu);
}

int boomgpu_student_draw(boomgpu_ctx *ctx, const double *beta, double sigma, double nu, uint64_t seed, uint64_t iteration,
                         double *weight_out) {
  if (int rc = check_student(ctx, sigma, nu)) return rc;
  if (!beta || !weight_out) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  if (int rc = ensure_suf(ctx)) return rc;
  double *dw = nullptr;
  CU(cudaMalloc((void **)&dw, sizeof(double) * (size_t)std::max<int64_t>(ctx->n, 1)));
  RowOut out{dw, nullptr, nullptr, nullptr};
  int rc = run_step<kStudentT>(ctx, beta, make_student_prm(ctx, sigma, nu, seed, iteration), out, nullptr, nullptr, ctx->suf_dev);
  if (!rc && cudaMemcpyAsync(weight_out, dw, sizeof(double) * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    rc = fail(ctx, BOOMGPU_ERR_CUDA, "copy of draws failed");
  if (!rc) rc = finish_and_check(ctx); else cudaStreamSynchronize(ctx->stream);
  cudaFree(dw);
  return rc;
}

// Observed-data log likelihood sum_i log dstudent(y_i; x_i'beta, sigma, nu).  beta != NULL: one pass over X stores the
// residuals; beta == NULL: the residuals of the previous call are reused (8 n bytes per evaluation) -- the slice sampler on
// nu evaluates the likelihood at several nu with beta fixed.  All-reduced when the context is sharded.
int boomgpu_student_loglike(boomgpu_ctx *ctx, const double *beta, double sigma, double nu, double *loglike) {
  if (int rc = check_student(ctx, sigma, nu)) return rc;
  if (!loglike) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  if (!beta && !ctx->resid_valid) return fail(ctx, BOOMGPU_ERR_STATE, "boomgpu_student_loglike(beta = NULL) before a call with beta");
  DeviceGuard g(ctx->device);
  const int p = ctx->p;
  const int64_t n = ctx->n;
  if (n == 0) { *loglike = 0.0; ctx->resid_valid = true; return 0; }   // sharded callers add their own shards
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 1023) / 1024, (int64_t)ctx->sms * 8));
  if (ctx->resid_cap < n + grid + 8) ctx->resid_valid = false;
  if (!beta && !ctx->resid_valid) return fail(ctx, BOOMGPU_ERR_STATE, "boomgpu_student_loglike(beta = NULL) before a call with beta");
  if (int rc = ensure(ctx, &ctx->resid_buf, &ctx->resid_cap, n + grid + 8)) return rc;   // [residuals | per-CTA partials | result, n]
  double *parts = ctx->resid_buf + n;
  if (beta) {
    ctx->resid_valid = false;
    if (ctx->beta_cap < p + 2) {   // the step's beta staging buffers
      if (ctx->beta_dev) { CU(cudaFree(ctx->beta_dev)); ctx->beta_dev = nullptr; }
      if (ctx->beta_pin) { CU(cudaFreeHost(ctx->beta_pin)); ctx->beta_pin = nullptr; }
      const size_t bytes = sizeof(double) * (size_t)(p + 2) + sizeof(int) * (size_t)(p + 2);
      CU(cudaMalloc((void **)&ctx->beta_dev, bytes));
      CU(cudaMallocHost((void **)&ctx->beta_pin, bytes));
      ctx->beta_cap = p + 2;
    }
    memcpy(ctx->beta_pin, beta, sizeof(double) * p);
    CU(cudaMemcpyAsync(ctx->beta_dev, ctx->beta_pin, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
    RowData d;
    memset(&d, 0, sizeof(d));
    d.X = ctx->X; d.ldx = ctx->ldx; d.n = n; d.p = p; d.y = ctx->y;
    LaunchScope ls(ctx, 4);
    if (p <= 128) {   // G lanes per row
      const int rgrid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sms * 8));
      const size_t sm = sizeof(double) * p;
      if (p <= 1) residual_rows_kernel<1><<<rgrid, 256, sm, ctx->stream>>>(d, ctx->beta_dev, ctx->resid_buf);
      else if (p <= 2) residual_rows_kernel<2><<<rgrid, 256, sm, ctx->stream>>>(d, ctx->beta_dev, ctx->resid_buf);
      else if (p <= 4) residual_rows_kernel<4><<<rgrid, 256, sm, ctx->stream>>>(d, ctx->beta_dev, ctx->resid_buf);
      else if (p <= 8) residual_rows_kernel<8><<<rgrid, 256, sm, ctx->stream>>>(d, ctx->beta_dev, ctx->resid_buf);
      else if (p <= 16) residual_rows_kernel<16><<<rgrid, 256, sm, ctx->stream>>>(d, ctx->beta_dev, ctx->resid_buf);
      else residual_rows_kernel<32><<<rgrid, 256, sm, ctx->stream>>>(d, ctx->beta_dev, ctx->resid_buf);
    } else {
      const int rgrid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, (int64_t)ctx->sms * 8));
      residual_kernel<<<rgrid, 256, sizeof(double) * p, ctx->stream>>>(d, ctx->beta_dev, ctx->resid_buf);
    }
  }
  {
    LaunchScope ls(ctx, 4);
    student_loglike_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->resid_buf, n, 1.0 / sigma, nu, parts);
  }
  {
    LaunchScope ls(ctx, 3);
    reduce_sum_warp_kernel<<<1, 32, 0, ctx->stream>>>(parts, grid, parts + grid);
  }
  CU(cudaGetLastError());
  // [row part | n]: both sum over the shards
  const double nd = (double)n;
  CU(cudaMemcpyAsync(parts + grid + 1, &nd, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (int rc = allreduce_on_stream(ctx, parts + grid, 2)) return rc;
  double res[2] = {0, 0};
  CU(cudaMemcpyAsync(res, parts + grid, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->resid_valid = true;
  const double kLogPi = 1.1447298858494001741434;
  *loglike = res[0] + res[1] * (lgamma(0.5 * (nu + 1.0)) - lgamma(0.5 * nu) - 0.5 * (log(nu) + kLogPi) - log(sigma));
  return 0;
}

int boomgpu_accumulate(boomgpu_ctx *ctx, const double *weight, const double *weighted_value, double *xtx, double *xty) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (!ctx->X) return fail(ctx, BOOMGPU_ERR_STATE, "no data uploaded to this context");
  if (!weight || !weighted_value || !xtx || !xty) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  if (int rc = ensure_suf(ctx)) return rc;
  double *dw = nullptr, *ds = nullptr;
  const size_t bytes = sizeof(double) * (size_t)std::max<int64_t>(ctx->n, 1);
  CU(cudaMalloc((void **)&dw, bytes));
  CU(cudaMalloc((void **)&ds, bytes));
  int rc = 0;
  cudaError_t e1 = cudaMemcpyAsync(dw, weight, sizeof(double) * (size_t)ctx->n, cudaMemcpyHostToDevice, ctx->stream);
  cudaError_t e2 = cudaMemcpyAsync(ds, weighted_value, sizeof(double) * (size_t)ctx->n, cudaMemcpyHostToDevice, ctx->stream);
  if (e1 != cudaSuccess || e2 != cudaSuccess) rc = fail(ctx, BOOMGPU_ERR_CUDA, "copy of latents failed");
  RowOut out{nullptr, nullptr, nullptr, nullptr};
  DrawParams prm = make_prm(ctx, 0, 0, 0);
  if (!rc) rc = run_step<kSupplied>(ctx, nullptr, prm, out, dw, ds, ctx->suf_dev);
  const int p = ctx->p;
  if (!rc && cudaMemcpyAsync(ctx->suf_pin, ctx->suf_dev, sizeof(double) * (size_t)boomgpu_suf_len(p), cudaMemcpyDeviceToHost,
                             ctx->stream) != cudaSuccess)
    rc = fail(ctx, BOOMGPU_ERR_CUDA, "copy of statistics failed");
  if (!rc) rc = finish_and_check(ctx); else cudaStreamSynchronize(ctx->stream);
  cudaFree(dw); cudaFree(ds);
  if (rc) return rc;
  memcpy(xtx, ctx->suf_pin, sizeof(double) * (size_t)p * p);
  memcpy(xty, ctx->suf_pin + (size_t)p * p, sizeof(double) * p);
  return 0;
}

int boomgpu_logit_draw(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed, uint64_t iteration,
                       double *sum_out, double *info_out) {
  if (int rc = check_ready(ctx, kLogit)) return rc;
  if (int rc = check_clt(ctx, clt_threshold)) return rc;
  if (!beta || !sum_out || !info_out) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  if (int rc = ensure_suf(ctx)) return rc;
  double *dw = nullptr, *ds = nullptr;
  const size_t bytes = sizeof(double) * (size_t)std::max<int64_t>(ctx->n, 1);
  CU(cudaMalloc((void **)&dw, bytes));
  CU(cudaMalloc((void **)&ds, bytes));
  RowOut out{dw, ds, nullptr, nullptr};
  int rc = run_step<kLogit>(ctx, beta, make_prm(ctx, clt_threshold, seed, iteration), out, nullptr, nullptr, ctx->suf_dev);
  if (!rc && (cudaMemcpyAsync(info_out, dw, sizeof(double) * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
              cudaMemcpyAsync(sum_out, ds, sizeof(double) * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess))
    rc = fail(ctx, BOOMGPU_ERR_CUDA, "copy of draws failed");
  if (!rc) rc = finish_and_check(ctx); else cudaStreamSynchronize(ctx->stream);
  cudaFree(dw); cudaFree(ds);
  return rc;
}

int boomgpu_poisson_draw(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *out6, int32_t *kout2) {
  if (int rc = check_ready(ctx, kPoisson)) return rc;
  if (!beta || !out6) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  if (int rc = ensure_suf(ctx)) return rc;
  double *d6 = nullptr; int32_t *dk = nullptr;
  const size_t n1 = (size_t)std::max<int64_t>(ctx->n, 1);
  CU(cudaMalloc((void **)&d6, sizeof(double) * 6 * n1));
  CU(cudaMalloc((void **)&dk, sizeof(int32_t) * 2 * n1));
  CU(cudaMemsetAsync(d6, 0, sizeof(double) * 6 * n1, ctx->stream));
  RowOut out{nullptr, nullptr, d6, dk};
  int rc = run_step<kPoisson>(ctx, beta, make_prm(ctx, 0, seed, iteration), out, nullptr, nullptr, ctx->suf_dev);
  if (!rc && cudaMemcpyAsync(out6, d6, sizeof(double) * 6 * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    rc = fail(ctx, BOOMGPU_ERR_CUDA, "copy of draws failed");
  if (!rc && kout2 && cudaMemcpyAsync(kout2, dk, sizeof(int32_t) * 2 * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    rc = fail(ctx, BOOMGPU_ERR_CUDA, "copy of indicators failed");
  if (!rc) rc = finish_and_check(ctx); else cudaStreamSynchronize(ctx->stream);
  cudaFree(d6); cudaFree(dk);
  return rc;
}

static int loglike_impl(boomgpu_ctx *ctx, int model, const double *beta, double *loglike) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (ctx->model != model || !ctx->X) return fail(ctx, BOOMGPU_ERR_STATE, "no matching data uploaded to this context");
  if (!beta || !loglike) return fail(ctx, BOOMGPU_ERR_ARG, "null argument");
  DeviceGuard g(ctx->device);
  const int p = ctx->p;
  double *dbeta = nullptr, *parts = nullptr;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ctx->n + 7) / 8, (int64_t)ctx->sms * 8));
  CU(cudaMalloc((void **)&dbeta, sizeof(double) * p));
  CU(cudaMalloc((void **)&parts, sizeof(double) * (grid + 1)));
  CU(cudaMemcpyAsync(dbeta, beta, sizeof(double) * p, cudaMemcpyHostToDevice, ctx->stream));
  RowData d;
  memset(&d, 0, sizeof(d));
  d.X = ctx->X; d.ldx = ctx->ldx; d.n = ctx->n; d.p = p;
  d.y = ctx->y; d.ntrials = ctx->ntrials; d.yi = ctx->yi; d.exposure = ctx->exposure;
  {
    LaunchScope ls(ctx, 4);
    if (model == kLogit) loglike_kernel<kLogit><<<grid, 256, sizeof(double) * p, ctx->stream>>>(d, dbeta, parts);
    else loglike_kernel<kPoisson><<<grid, 256, sizeof(double) * p, ctx->stream>>>(d, dbeta, parts);
  }
  {
    LaunchScope ls(ctx, 3);
    reduce_sum_kernel<<<1, 32, 0, ctx->stream>>>(parts, grid, parts + grid);
  }
  cudaError_t e = cudaGetLastError();
  double result = 0;
  if (e == cudaSuccess && allreduce_on_stream(ctx, parts + grid, 1)) { cudaStreamSynchronize(ctx->stream); cudaFree(dbeta); cudaFree(parts); return BOOMGPU_ERR_CUDA; }
  if (e == cudaSuccess) e = cudaMemcpyAsync(&result, parts + grid, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dbeta); cudaFree(parts);
  if (e != cudaSuccess) return fail(ctx, BOOMGPU_ERR_CUDA, "log likelihood failed: %s", cudaGetErrorString(e));
  *loglike = ctx->n == 0 ? 0.0 : result;
  return 0;
}

int boomgpu_binomial_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double log_alpha, double *loglike, double *gradient,
                                    double *hessian) {
  return loglike_derivs_impl<kLogitLL>(ctx, kLogit, beta, log_alpha, loglike, gradient, hessian);
}
int boomgpu_poisson_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double *loglike, double *gradient, double *hessian) {
  return loglike_derivs_impl<kPoissonLL>(ctx, kPoisson, beta, 0.0, loglike, gradient, hessian);
}

int boomgpu_binomial_loglike_derivs_device(boomgpu_ctx *ctx, const double *beta, double log_alpha, double *suf_dev) {
  return loglike_derivs_device_impl<kLogitLL>(ctx, kLogit, beta, log_alpha, suf_dev);
}
int boomgpu_poisson_loglike_derivs_device(boomgpu_ctx *ctx, const double *beta, double *suf_dev) {
  return loglike_derivs_device_impl<kPoissonLL>(ctx, kPoisson, beta, 0.0, suf_dev);
}

int boomgpu_select_columns(boomgpu_ctx *ctx, const int32_t *cols, int k) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  if (!ctx->X || ctx->model < 0) return fail(ctx, BOOMGPU_ERR_STATE, "no data uploaded to this context");
  if (k < 0 || k > ctx->p || (k > 0 && !cols)) return fail(ctx, BOOMGPU_ERR_ARG, "bad column selection (k = %d of p = %d)", k, ctx->p);
  for (int j = 0; j < k; ++j)
    if (cols[j] < 0 || cols[j] >= ctx->p) return fail(ctx, BOOMGPU_ERR_ARG, "selected column %d out of range", cols[j]);
  if (ctx->Xsel && ctx->sel_src == ctx->X && (int)ctx->sel_cols.size() == k && (k == 0 || !memcmp(cols, ctx->sel_cols.data(), sizeof(int) * k)))
    return 0;   // already there
  ctx->sel_cols.assign(cols, cols + k);
  ctx->sel_src = ctx->X;
  if (k == 0) return 0;
  DeviceGuard g(ctx->device);
  const int ld = (k + 7) / 8 * 8;
  if (int rc = ensure(ctx, &ctx->Xsel, &ctx->sel_cap, std::max<int64_t>(ctx->n, 1) * ld)) { ctx->sel_cols.clear(); return rc; }
  if (ctx->sel_cols_cap < k) {
    if (ctx->sel_cols_dev) { CU(cudaFree(ctx->sel_cols_dev)); ctx->sel_cols_dev = nullptr; ctx->sel_cols_cap = 0; }
    CU(cudaMalloc((void **)&ctx->sel_cols_dev, sizeof(int) * (size_t)ctx->p));
    ctx->sel_cols_cap = ctx->p;
  }
  ctx->sel_ld = ld;
  CU(cudaMemcpyAsync(ctx->sel_cols_dev, cols, sizeof(int) * k, cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->n > 0) {
    LaunchScope ls(ctx, 4);
    const int64_t total = ctx->n * ld;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sms * 16));
    select_columns_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->X, ctx->ldx, ctx->n, ctx->sel_cols_dev, k, ctx->Xsel, ld);
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));   // cols is the caller's buffer
  return 0;
}

int boomgpu_binomial_loglike_derivs_selected(boomgpu_ctx *ctx, const double *beta_selected, double log_alpha, double *loglike,
                                             double *gradient, double *hessian) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  return with_selected_columns(ctx, [&]() { return loglike_derivs_impl<kLogitLL>(ctx, kLogit, beta_selected, log_alpha, loglike, gradient, hessian); });
}
int boomgpu_poisson_loglike_derivs_selected(boomgpu_ctx *ctx, const double *beta_selected, double *loglike, double *gradient,
                                            double *hessian) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  return with_selected_columns(ctx, [&]() { return loglike_derivs_impl<kPoissonLL>(ctx, kPoisson, beta_selected, 0.0, loglike, gradient, hessian); });
}
int boomgpu_binomial_loglike_derivs_selected_device(boomgpu_ctx *ctx, const double *beta_selected, double log_alpha, double *suf_dev) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  return with_selected_columns(ctx, [&]() { return loglike_derivs_device_impl<kLogitLL>(ctx, kLogit, beta_selected, log_alpha, suf_dev); });
}

int boomgpu_binomial_loglike(boomgpu_ctx *ctx, const double *beta, double *loglike) { return loglike_impl(ctx, kLogit, beta, loglike); }
int boomgpu_poisson_loglike(boomgpu_ctx *ctx, const double *beta, double *loglike) { return loglike_impl(ctx, kPoisson, beta, loglike); }

int64_t boomgpu_kernel_launches(const boomgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int boomgpu_get_timings(boomgpu_ctx *ctx, double ms[BOOMGPU_NUM_KERNEL_CLASSES], int64_t launches[BOOMGPU_NUM_KERNEL_CLASSES],
                        int reset) {
  if (!ctx) return BOOMGPU_ERR_ARG;
  DeviceGuard g(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  drain_timings(ctx);
  for (int c = 0; c < BOOMGPU_NUM_KERNEL_CLASSES; ++c) {
    if (ms) ms[c] = ctx->ms_acc[c];
    if (launches) launches[c] = ctx->n_acc[c];
    if (reset) { ctx->ms_acc[c] = 0; ctx->n_acc[c] = 0; }
  }
  return 0;
}

}  // extern "C"
