// Host side of the C ABI (include/lsqr_b200.h): context, device buffers, and the orchestration
// that replaces RANSAC<T,S>::compute (parametersEstimators/RANSAC.hxx:4-249).  No arithmetic of
// the path happens on the host: sampling, minimal solves, consensus, arg-max, consensus set and
// least squares (including the small eigen / LM controllers) are all kernels.  The host only
// sequences launches, evaluates the scalar stop rule (RANSAC.hxx:107-110) between rounds and
// moves results.  There is no CPU fallback: every entry point fails with LSQR_ERR_CUDA when no
// device is usable.
#include "../../include/lsqr_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "engine.h"

using namespace lsqr;

namespace {
constexpr int kSmall = 1024, kSmIn = 0, kSmOut = 32, kSmLm = 64, kSmEst = 640, kSmSink = 800;   // layout of lsqr_ctx::small_dev
}

namespace {

struct DataSet {
  double* soa64 = nullptr;
  float* soa32 = nullptr;
  uint32_t* maskbits = nullptr;
  size_t ld = 0, cap = 0;   // cap = allocated leading dimension
  uint32_t n = 0;
  int D = 0, capD = 0;
  double center[kMaxDim] = {0};
  bool moments_valid = false;  // rb.moments holds the LS moments of the stored consensus set
  bool mask_valid = false;
  DataView view() const {
    DataView v;
    v.soa64 = soa64; v.soa32 = soa32; v.ld = ld; v.n = n;
    for (int i = 0; i < kMaxDim; i++) v.center[i] = center[i];
    return v;
  }
};

}  // namespace

struct lsqr_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  uint64_t launches = 0;

  int model = -1;
  EstCfg cfg{};
  double delta = 0, aux = 0;
  int ls_type = 1;

  DataSet main, scratch;
  unsigned char* staging = nullptr; size_t staging_cap = 0;

  // hypothesis buffers
  size_t hcap = 0;
  int32_t* subsets = nullptr; double* hyp64 = nullptr; float* hyp32 = nullptr; uint32_t* counts = nullptr;
  int32_t* list_dev = nullptr; size_t list_cap = 0;
  double* params_in_dev = nullptr; size_t params_in_cap = 0;
  unsigned long long* key_dev = nullptr;  // [0] key, [1] n_valid (as u32 in low half)
  double* small_dev = nullptr;            // kSmall doubles: parameters in @kSmIn, solve out @kSmOut, LM state @kSmLm, estimate() input @kSmEst, sink @kSmSink
  double* center_dev = nullptr;           // kMaxDim doubles
  double* center_partials = nullptr;      // 256 * kMaxDim
  // refine
  RefineBuffers rb{};
  // pinned host scratch
  double* pin = nullptr;                  // 64 doubles
  double* weights_dev = nullptr; size_t weights_cap = 0;   // lsqr_weighted_least_squares
  double* bt_data = nullptr; size_t bt_data_cap = 0;       // lsqr_ransac_batch: packed problems, offsets, results
  uint64_t* bt_off = nullptr; size_t bt_off_cap = 0;
  double* bt_prm = nullptr; size_t bt_prm_cap = 0;
  uint32_t* bt_cnt = nullptr; size_t bt_cnt_cap = 0;
  uint8_t* mask_dev = nullptr; size_t mask_dev_cap = 0;   // consensus set, one byte per datum (device)
  uint8_t* mask_pin = nullptr; size_t mask_pin_cap = 0;   // pinned bounce buffer for its download
  cudaEvent_t ev[6]{};

  // sharding
  int rank = 0, world = 1;
  lsqr_allreduce_max_u64_fn max_fn = nullptr;
  lsqr_allreduce_sum_f64_fn sum_fn = nullptr;
  void* comm_user = nullptr;

  // stats of the last refine
  double refine_kernel_ms = 0, refine_bytes = 0;
  int lm_iterations = 0;
};

namespace {

int fail(lsqr_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess) return fail(ctx, LSQR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)
#define CKL()                                                                                            \
  do {                                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                                \
    if (e__ != cudaSuccess) return fail(ctx, LSQR_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__)); \
  } while (0)

template <class T> int ensure(lsqr_ctx* ctx, T** p, size_t* cap, size_t want) {
  if (*cap >= want && *p) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  CK(cudaMalloc((void**)p, want * sizeof(T)));
  *cap = want;
  return 0;
}

size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

int ensure_dataset(lsqr_ctx* ctx, DataSet& ds, int D, uint32_t n) {
  const size_t ld = round_up(std::max<size_t>(n, 1), kTilePad);
  if (ds.cap < ld || ds.capD < D) {
    if (ds.soa64) cudaFree(ds.soa64);
    if (ds.soa32) cudaFree(ds.soa32);
    if (ds.maskbits) cudaFree(ds.maskbits);
    ds.soa64 = nullptr; ds.soa32 = nullptr; ds.maskbits = nullptr; ds.cap = 0;
    const int Dc = std::max(D, ds.capD);
    CK(cudaMalloc((void**)&ds.soa64, sizeof(double) * Dc * ld));
    CK(cudaMalloc((void**)&ds.soa32, sizeof(float) * Dc * ld));
    CK(cudaMalloc((void**)&ds.maskbits, sizeof(uint32_t) * (ld / 32)));
    ds.cap = ld; ds.capD = Dc;
  }
  ds.ld = ld; ds.n = n; ds.D = D;
  ds.moments_valid = false; ds.mask_valid = false;
  return 0;
}

int ensure_hyp(lsqr_ctx* ctx, size_t H) {
  const size_t want = round_up(std::max<size_t>(H, 1), 256);
  if (ctx->hcap >= want) return 0;
  if (ctx->subsets) cudaFree(ctx->subsets);
  if (ctx->hyp64) cudaFree(ctx->hyp64);
  if (ctx->hyp32) cudaFree(ctx->hyp32);
  if (ctx->counts) cudaFree(ctx->counts);
  ctx->subsets = nullptr; ctx->hyp64 = nullptr; ctx->hyp32 = nullptr; ctx->counts = nullptr; ctx->hcap = 0;
  CK(cudaMalloc((void**)&ctx->subsets, sizeof(int32_t) * LSQR_MAX_SUBSET * want));
  CK(cudaMalloc((void**)&ctx->hyp64, sizeof(double) * LSQR_MAX_PARAMS * want));
  CK(cudaMalloc((void**)&ctx->hyp32, sizeof(float) * 12 * want));
  CK(cudaMalloc((void**)&ctx->counts, sizeof(uint32_t) * want));
  ctx->hcap = want;
  return 0;
}

// AoS on the device -> SoA fp64 + centre + fp32 copy.
int build_layouts(lsqr_ctx* ctx, DataSet& ds, const unsigned char* aos_dev, size_t stride) {
  launch_ingest(ds.D, aos_dev, stride, ds.n, ds.soa64, ds.ld, ctx->stream); ctx->launches++;
  launch_center(ctx->model, ds.soa64, ds.ld, ds.n, ctx->center_partials, ctx->center_dev, ctx->stream); ctx->launches += 2;
  launch_make32(ds.D, ds.soa64, ctx->center_dev, ds.soa32, ds.ld, ctx->stream); ctx->launches++;
  CKL();
  CK(cudaMemsetAsync(ds.maskbits, 0, sizeof(uint32_t) * (ds.ld / 32), ctx->stream));
  CK(cudaMemcpyAsync(ctx->pin, ctx->center_dev, sizeof(double) * kMaxDim, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < kMaxDim; i++) ds.center[i] = ctx->pin[i];
  return 0;
}

int upload_to(lsqr_ctx* ctx, DataSet& ds, const void* aos, size_t n, size_t stride, bool on_device) {
  if (ctx->model < 0) return fail(ctx, LSQR_ERR_STATE, "lsqr_set_estimator must be called before uploading data");
  const ModelInfo mi = model_info(ctx->model);
  if (n > 0xFFFFFFF0ull) return fail(ctx, LSQR_ERR_ARG, "too many data records");
  if (stride < sizeof(double) * mi.D || (stride % sizeof(double)) != 0) return fail(ctx, LSQR_ERR_ARG, "stride must be a multiple of 8 and >= 8*dim");
  if (n && !aos) return fail(ctx, LSQR_ERR_ARG, "null data pointer");
  CK(cudaSetDevice(ctx->device));
  if (int rc = ensure_dataset(ctx, ds, mi.D, (uint32_t)n)) return rc;
  const unsigned char* src = static_cast<const unsigned char*>(aos);
  if (!on_device && n) {
    if (int rc = ensure(ctx, &ctx->staging, &ctx->staging_cap, n * stride)) return rc;
    CK(cudaMemcpyAsync(ctx->staging, aos, n * stride, cudaMemcpyHostToDevice, ctx->stream));
    src = ctx->staging;
  }
  return build_layouts(ctx, ds, src, stride);
}

// Number of subsets as the reference counts them (RANSAC::choose, RANSAC.hxx:254-280): the binomial
// coefficient evaluated in double precision as a ratio of two running products, saturating at
// UINT_MAX when either the result or an intermediate does not fit.  The stop rule clamps the
// number of tries with this value (RANSAC.hxx:41,110), so the same saturation must happen here.
unsigned int choose_ref(unsigned int n, unsigned int m) {
  const unsigned int small = std::min(m, n - m);         // multiply the shorter run of factors
  double num = 1.0, den = 1.0;
  for (unsigned int f = n - small + 1; f <= n && f != 0; f++) num *= static_cast<double>(f);
  for (unsigned int f = 1; f <= small; f++) den *= static_cast<double>(f);
  const double c = num / den;
  const double dmax = std::numeric_limits<double>::max(), umax = static_cast<double>(std::numeric_limits<unsigned int>::max());
  if (num > dmax || den > dmax || c > umax) return std::numeric_limits<unsigned int>::max();
  return static_cast<unsigned int>(c);
}

uint64_t choose_exact(uint64_t n, uint64_t k) {
  if (k > n) return 0;
  long double r = 1;
  for (uint64_t i = 0; i < k; i++) r = r * (long double)(n - i) / (long double)(i + 1);
  if (r > 1.8e19L) return ~0ull;
  return (uint64_t)(r + 0.5L);
}

constexpr size_t kHypBatch = (size_t)1 << 22;

int score_impl(lsqr_ctx* ctx, const lsqr_score_args* a, lsqr_score_result* res) {
  DataSet& ds = ctx->main;
  if (!ds.soa64) return fail(ctx, LSQR_ERR_STATE, "no data uploaded");
  if (!a || !res) return fail(ctx, LSQR_ERR_ARG, "null argument");
  const ModelInfo mi = model_info(ctx->model);
  if (a->count > 0xFFFFFFFEull) return fail(ctx, LSQR_ERR_ARG, "at most 2^32-2 hypotheses per request");
  if (a->sampler < 0 || a->sampler > 3) return fail(ctx, LSQR_ERR_ARG, "bad sampler");
  if (a->sampler == LSQR_SAMPLE_LIST && !a->subsets) return fail(ctx, LSQR_ERR_ARG, "subset list missing");
  if (a->sampler == LSQR_SAMPLE_PARAMS && !a->params) return fail(ctx, LSQR_ERR_ARG, "parameter list missing");
  CK(cudaSetDevice(ctx->device));
  memset(res, 0, sizeof(*res));
  const uint64_t H = a->count;
  const uint64_t lo = H * (uint64_t)ctx->rank / (uint64_t)ctx->world, hi = H * (uint64_t)(ctx->rank + 1) / (uint64_t)ctx->world;
  const bool can_sample = ds.n >= (uint32_t)mi.K || a->sampler == LSQR_SAMPLE_PARAMS;
  const DataView dv = ds.view();
  cudaStream_t s = ctx->stream;
  CK(cudaMemsetAsync(ctx->key_dev, 0, 2 * sizeof(unsigned long long), s));
  CK(cudaEventRecord(ctx->ev[0], s));
  float cons_ms = 0.f;
  if (can_sample) {
    for (uint64_t b0 = lo; b0 < hi; b0 += kHypBatch) {
      const uint32_t B = (uint32_t)std::min<uint64_t>(kHypBatch, hi - b0);
      if (int rc = ensure_hyp(ctx, B)) return rc;
      SolveArgs sa{};
      sa.model = ctx->model; sa.sampler = a->sampler; sa.seed = a->seed; sa.first = a->first + b0; sa.H = B; sa.hld = ctx->hcap;
      sa.subsets = ctx->subsets; sa.hyp64 = ctx->hyp64; sa.n_valid = reinterpret_cast<uint32_t*>(ctx->key_dev + 1);
      if (a->sampler == LSQR_SAMPLE_LIST) {
        if (int rc = ensure(ctx, &ctx->list_dev, &ctx->list_cap, (size_t)B * mi.K)) return rc;
        CK(cudaMemcpyAsync(ctx->list_dev, a->subsets + b0 * mi.K, sizeof(int32_t) * (size_t)B * mi.K, cudaMemcpyHostToDevice, s));
        sa.list = ctx->list_dev;
      } else if (a->sampler == LSQR_SAMPLE_PARAMS) {
        if (int rc = ensure(ctx, &ctx->params_in_dev, &ctx->params_in_cap, (size_t)B * mi.P)) return rc;
        CK(cudaMemcpyAsync(ctx->params_in_dev, a->params + b0 * mi.P, sizeof(double) * (size_t)B * mi.P, cudaMemcpyHostToDevice, s));
        sa.params_in = ctx->params_in_dev;
      }
      launch_solve(sa, dv, ctx->cfg, s); ctx->launches++;
      if (a->precision == LSQR_FP32) { launch_hoist32(ctx->model, ctx->hyp64, ctx->hcap, B, dv, ctx->cfg, ctx->hyp32, s); ctx->launches++; }
      CK(cudaMemsetAsync(ctx->counts, 0, sizeof(uint32_t) * B, s));
      CK(cudaEventRecord(ctx->ev[2], s));
      ctx->launches += launch_consensus(ctx->model, a->precision, dv, ctx->hyp64, ctx->hyp32, ctx->hcap, B, ctx->cfg, ctx->counts, ctx->num_sms, s);
      CK(cudaEventRecord(ctx->ev[3], s));
      launch_argmax(ctx->counts, B, (uint32_t)b0, ctx->key_dev, s); ctx->launches++;
      CKL();
      if (a->out_counts) CK(cudaMemcpyAsync(a->out_counts + b0, ctx->counts, sizeof(uint32_t) * B, cudaMemcpyDeviceToHost, s));
      if (a->out_params) {
        // device layout is [P][hld]; gather rows on the host side of the copy
        std::vector<double> tmp((size_t)B * mi.P);
        CK(cudaMemcpy2DAsync(tmp.data(), sizeof(double) * B, ctx->hyp64, sizeof(double) * ctx->hcap, sizeof(double) * B, mi.P, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (uint32_t h = 0; h < B; h++) for (int j = 0; j < mi.P; j++) a->out_params[(b0 + h) * mi.P + j] = tmp[(size_t)j * B + h];
      }
      CK(cudaEventSynchronize(ctx->ev[3]));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
      cons_ms += ms;
    }
  }
  if (ctx->world > 1) {
    if (!ctx->max_fn) return fail(ctx, LSQR_ERR_COMM, "sharded context without an all-reduce hook");
    if (ctx->max_fn(ctx->comm_user, reinterpret_cast<uint64_t*>(ctx->key_dev), (void*)s)) return fail(ctx, LSQR_ERR_COMM, "max all-reduce hook failed");
  }
  CK(cudaEventRecord(ctx->ev[1], s));
  unsigned long long hk[2] = {0, 0};
  CK(cudaMemcpyAsync(hk, ctx->key_dev, sizeof(hk), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  float total_ms = 0.f;
  CK(cudaEventElapsedTime(&total_ms, ctx->ev[0], ctx->ev[1]));
  res->score_ms = total_ms; res->consensus_ms = cons_ms;
  res->n_valid = (uint32_t)(hk[1] & 0xFFFFFFFFull);
  res->best_count = (uint32_t)(hk[0] >> 32);
  for (int j = 0; j < LSQR_MAX_SUBSET; j++) res->best_subset[j] = -1;
  if (res->best_count == 0) { res->best_index = 0; return LSQR_OK; }
  const uint64_t rel = 0xFFFFFFFFull - (hk[0] & 0xFFFFFFFFull);
  res->best_index = a->first + rel;
  // Re-derive the winner's subset and parameters (same kernel, one hypothesis) so that every rank
  // holds them regardless of which shard produced the winner.
  if (int rc = ensure_hyp(ctx, 1)) return rc;
  SolveArgs sa{};
  sa.model = ctx->model; sa.sampler = a->sampler; sa.seed = a->seed; sa.first = res->best_index; sa.H = 1; sa.hld = ctx->hcap;
  sa.subsets = ctx->subsets; sa.hyp64 = ctx->hyp64; sa.n_valid = reinterpret_cast<uint32_t*>(ctx->key_dev + 1);
  if (a->sampler == LSQR_SAMPLE_LIST) {
    if (int rc = ensure(ctx, &ctx->list_dev, &ctx->list_cap, (size_t)mi.K)) return rc;
    CK(cudaMemcpyAsync(ctx->list_dev, a->subsets + rel * mi.K, sizeof(int32_t) * mi.K, cudaMemcpyHostToDevice, s));
    sa.list = ctx->list_dev;
  } else if (a->sampler == LSQR_SAMPLE_PARAMS) {
    if (int rc = ensure(ctx, &ctx->params_in_dev, &ctx->params_in_cap, (size_t)mi.P)) return rc;
    CK(cudaMemcpyAsync(ctx->params_in_dev, a->params + rel * mi.P, sizeof(double) * mi.P, cudaMemcpyHostToDevice, s));
    sa.params_in = ctx->params_in_dev;
  }
  launch_solve(sa, dv, ctx->cfg, s); ctx->launches++;
  CKL();
  double hp[LSQR_MAX_PARAMS]; int32_t hs[LSQR_MAX_SUBSET];
  CK(cudaMemcpy2DAsync(hp, sizeof(double), ctx->hyp64, sizeof(double) * ctx->hcap, sizeof(double), mi.P, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpy2DAsync(hs, sizeof(int32_t), ctx->subsets, sizeof(int32_t) * ctx->hcap, sizeof(int32_t), mi.K, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (int j = 0; j < mi.P; j++) res->best_params[j] = hp[j];
  for (int j = 0; j < mi.K; j++) res->best_subset[j] = hs[j];
  return LSQR_OK;
}

int reduce_moments(lsqr_ctx* ctx, int nm) {
  launch_reduce_partials(ctx->rb, nm, ctx->stream); ctx->launches++;
  if (ctx->world > 1) {
    if (!ctx->sum_fn) return fail(ctx, LSQR_ERR_COMM, "sharded context without a sum all-reduce hook");
    if (ctx->sum_fn(ctx->comm_user, ctx->rb.moments, nm, (void*)ctx->stream)) return fail(ctx, LSQR_ERR_COMM, "sum all-reduce hook failed");
  }
  return 0;
}

void shard_range(const lsqr_ctx* ctx, uint32_t n, uint32_t* begin, uint32_t* end) {
  if (ctx->world <= 1) { *begin = 0; *end = n; return; }
  const uint64_t words = ((uint64_t)n + 31) / 32;
  const uint64_t w0 = words * (uint64_t)ctx->rank / (uint64_t)ctx->world, w1 = words * (uint64_t)(ctx->rank + 1) / (uint64_t)ctx->world;
  *begin = (uint32_t)std::min<uint64_t>(w0 * 32, n);
  *end = (uint32_t)std::min<uint64_t>(w1 * 32, n);
}

int consensus_impl(lsqr_ctx* ctx, DataSet& ds, const double* params, uint32_t* out_count) {
  if (!ds.soa64) return fail(ctx, LSQR_ERR_STATE, "no data uploaded");
  const ModelInfo mi = model_info(ctx->model);
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  for (int j = 0; j < mi.P; j++) ctx->pin[32 + j] = params[j];
  CK(cudaMemcpyAsync(ctx->small_dev + kSmIn, ctx->pin + 32, sizeof(double) * mi.P, cudaMemcpyHostToDevice, s));
  uint32_t b, e;
  shard_range(ctx, ds.n, &b, &e);
  const int nm = moments_count(ctx->model, false);
  CK(cudaEventRecord(ctx->ev[4], s));
  ctx->rb.maskbits = ds.maskbits;
  launch_mask_moments(ctx->model, ds.view(), b, e, ctx->small_dev + kSmIn, 1, nullptr, ctx->cfg, ctx->rb, s); ctx->launches++;
  CK(cudaEventRecord(ctx->ev[5], s));
  if (int rc = reduce_moments(ctx, nm)) return rc;
  CKL();
  CK(cudaMemcpyAsync(ctx->pin, ctx->rb.moments, sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
  ctx->refine_kernel_ms = ms;
  ctx->refine_bytes = (double)(e - b) * mi.D * sizeof(double) + (double)(e - b) / 8.0;
  ds.moments_valid = true; ds.mask_valid = true;
  if (out_count) *out_count = (uint32_t)(ctx->pin[0] + 0.5);
  return LSQR_OK;
}

int refine_impl(lsqr_ctx* ctx, DataSet& ds, int use_mask, double* out_params, int* n_params) {
  if (!ds.soa64) return fail(ctx, LSQR_ERR_STATE, "no data uploaded");
  if (use_mask && !ds.mask_valid) return fail(ctx, LSQR_ERR_STATE, "no consensus set stored; call lsqr_consensus first");
  const ModelInfo mi = model_info(ctx->model);
  cudaStream_t s = ctx->stream;
  CK(cudaSetDevice(ctx->device));
  const DataView dv = ds.view();
  ctx->rb.maskbits = ds.maskbits;
  uint32_t b, e;
  shard_range(ctx, ds.n, &b, &e);
  const int nm = moments_count(ctx->model, false);
  ctx->lm_iterations = 0;
  if (!(use_mask && ds.moments_valid)) {
    CK(cudaEventRecord(ctx->ev[4], s));
    launch_mask_moments(ctx->model, dv, b, e, nullptr, use_mask ? 2 : 0, nullptr, ctx->cfg, ctx->rb, s); ctx->launches++;
    CK(cudaEventRecord(ctx->ev[5], s));
    if (int rc = reduce_moments(ctx, nm)) return rc;
    ds.moments_valid = false;
  }
  double* out_dev = ctx->small_dev + kSmOut;
  // iterative refinement: geometric circle / sphere fit, iterative cross-wire calibration (ls_type 1 in both)
  const bool geometric = (ctx->model == CIRCLE2 || ctx->model == SPHERE3 || ctx->model == SPHERE4 || ctx->model == USXW || ctx->model == USCP) && ctx->ls_type == LSQR_LS_GEOMETRIC;
  launch_solve_moments(ctx->model, dv, ctx->rb.moments, geometric ? 1 : 0, out_dev, s); ctx->launches++;
  if (geometric) {
    // SphereParametersEstimator.hxx:224-230: algebraic fit as the start, then Levenberg-Marquardt.
    double* st = ctx->small_dev + kSmLm;
    const int nlm = moments_count(ctx->model, true);
    launch_lm_init(out_dev, st, s); ctx->launches++;
    ds.moments_valid = false;
    // The controller stops itself (500 / 5000 evaluations at most).  Passes are enqueued kLmAhead evaluations ahead of the
    // status word, so the host synchronises once per chunk instead of once per evaluation; passes behind a stop are no-ops.
    constexpr int kLmAhead = 8;
    for (int it = 0; it < 5200; it += kLmAhead) {
      for (int k = 0; k < kLmAhead; k++) {
        launch_mask_moments(ctx->model, dv, b, e, nullptr, use_mask ? 2 : 0, st, ctx->cfg, ctx->rb, s); ctx->launches++;
        if (int rc = reduce_moments(ctx, nlm)) return rc;
        launch_lm_update(ctx->model, ctx->rb.moments, st, s); ctx->launches++;
      }
      CK(cudaMemcpyAsync(ctx->pin, st + lm_status_offset(), 3 * sizeof(double), cudaMemcpyDeviceToHost, s));   // status, phase, evaluations
      CK(cudaStreamSynchronize(s));
      ctx->lm_iterations = (int)ctx->pin[2];
      if (ctx->pin[0] != 0.0) break;
    }
    launch_lm_finish(ctx->model, dv, st, out_dev, s); ctx->launches++;
  }
  CKL();
  CK(cudaMemcpyAsync(ctx->pin, out_dev, sizeof(double) * (1 + LSQR_MAX_PARAMS), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const int np = (int)ctx->pin[0];
  if (n_params) *n_params = np;
  if (out_params) for (int j = 0; j < np && j < mi.P; j++) out_params[j] = ctx->pin[1 + j];
  return LSQR_OK;
}

int get_mask_impl(lsqr_ctx* ctx, DataSet& ds, uint8_t* out_bytes) {
  if (!ds.mask_valid) return fail(ctx, LSQR_ERR_STATE, "no consensus set stored");
  if (!out_bytes || ds.n == 0) return LSQR_OK;
  // persistent device + pinned buffers: a cudaMalloc/cudaFree pair per call costs up to 0.3 s after the
  // thousands of launches of a large request, and a pageable destination is staged by the driver
  if (int rc = ensure(ctx, &ctx->mask_dev, &ctx->mask_dev_cap, (size_t)ds.n)) return rc;
  if (ctx->mask_pin_cap < ds.n) {
    if (ctx->mask_pin) cudaFreeHost(ctx->mask_pin);
    ctx->mask_pin = nullptr; ctx->mask_pin_cap = 0;
    CK(cudaMallocHost((void**)&ctx->mask_pin, ds.n));
    ctx->mask_pin_cap = ds.n;
  }
  launch_expand_mask(ds.maskbits, ds.n, ctx->mask_dev, ctx->stream); ctx->launches++;
  // a caller buffer that is itself page-locked takes the DMA directly
  cudaPointerAttributes attr{};
  const bool direct = cudaPointerGetAttributes(&attr, out_bytes) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  if (!direct) cudaGetLastError();   // an unregistered pointer is not an error here
  cudaError_t e1 = cudaMemcpyAsync(direct ? out_bytes : ctx->mask_pin, ctx->mask_dev, ds.n, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  if (!direct && e1 == cudaSuccess && e2 == cudaSuccess) memcpy(out_bytes, ctx->mask_pin, ds.n);
  if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(ctx, LSQR_ERR_CUDA, "mask download failed");
  return LSQR_OK;
}

// winner -> consensus set -> least squares: RANSAC.hxx:129-138 / :175-185
int finish_ransac(lsqr_ctx* ctx, const lsqr_score_result& best, uint8_t* out_mask, lsqr_compute_result* res) {
  res->best_count = best.best_count; res->best_index = best.best_index;
  if (best.best_count == 0) { res->n_params = 0; res->fraction = 0.0; return LSQR_OK; }
  uint32_t cnt = 0;
  if (int rc = consensus_impl(ctx, ctx->main, best.best_params, &cnt)) return rc;
  res->best_count = cnt;
  if (out_mask) if (int rc = get_mask_impl(ctx, ctx->main, out_mask)) return rc;
  int np = 0;
  if (int rc = refine_impl(ctx, ctx->main, 1, res->params, &np)) return rc;
  res->n_params = np;
  res->fraction = (double)cnt / (double)ctx->main.n;
  return LSQR_OK;
}

}  // namespace

extern "C" {

int lsqr_model_info(int model, int* dim, int* nparams, int* k) {
  if (model < 0 || model >= LSQR_NUM_MODELS) return LSQR_ERR_ARG;
  const ModelInfo mi = model_info(model);
  if (dim) *dim = mi.D;
  if (nparams) *nparams = mi.P;
  if (k) *k = mi.K;
  return LSQR_OK;
}

int lsqr_ctx_create(lsqr_ctx** out, int device) {
  if (!out) return LSQR_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) return LSQR_ERR_CUDA;
  lsqr_ctx* ctx = new lsqr_ctx();
  ctx->device = device;
  auto bail = [&](int code) { lsqr_ctx_destroy(ctx); return code; };
  if (cudaSetDevice(device) != cudaSuccess) return bail(LSQR_ERR_CUDA);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(LSQR_ERR_CUDA);
  if (prop.major < 10) { fprintf(stderr, "lsqr_b200: device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major, prop.minor); return bail(LSQR_ERR_CUDA); }
  ctx->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(LSQR_ERR_CUDA);
  ctx->own_stream = true;
  ctx->rb.blocks = ctx->num_sms * mask_moments_ctas_per_sm();   // one wave of resident CTAs
  bool ok = cudaMalloc((void**)&ctx->key_dev, 4 * sizeof(unsigned long long)) == cudaSuccess &&
            cudaMalloc((void**)&ctx->small_dev, kSmall * sizeof(double)) == cudaSuccess &&
            cudaMalloc((void**)&ctx->center_dev, kMaxDim * sizeof(double)) == cudaSuccess &&
            cudaMalloc((void**)&ctx->center_partials, 256 * kMaxDim * sizeof(double)) == cudaSuccess &&
            cudaMalloc((void**)&ctx->rb.partials, sizeof(double) * kMaxMoments * ctx->rb.blocks) == cudaSuccess &&
            cudaMalloc((void**)&ctx->rb.moments, sizeof(double) * kMaxMoments) == cudaSuccess &&
            cudaMallocHost((void**)&ctx->pin, 64 * sizeof(double)) == cudaSuccess;
  for (int i = 0; ok && i < 6; i++) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
  if (!ok) return bail(LSQR_ERR_CUDA);
  *out = ctx;
  return LSQR_OK;
}

void lsqr_ctx_destroy(lsqr_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (DataSet* ds : {&ctx->main, &ctx->scratch}) { cudaFree(ds->soa64); cudaFree(ds->soa32); cudaFree(ds->maskbits); }
  cudaFree(ctx->bt_data); cudaFree(ctx->bt_off); cudaFree(ctx->bt_prm); cudaFree(ctx->bt_cnt);
  cudaFree(ctx->weights_dev); cudaFree(ctx->mask_dev); if (ctx->mask_pin) cudaFreeHost(ctx->mask_pin);
  cudaFree(ctx->staging); cudaFree(ctx->subsets); cudaFree(ctx->hyp64); cudaFree(ctx->hyp32); cudaFree(ctx->counts);
  cudaFree(ctx->list_dev); cudaFree(ctx->params_in_dev); cudaFree(ctx->key_dev); cudaFree(ctx->small_dev);
  cudaFree(ctx->center_dev); cudaFree(ctx->center_partials); cudaFree(ctx->rb.partials); cudaFree(ctx->rb.moments);
  if (ctx->pin) cudaFreeHost(ctx->pin);
  for (int i = 0; i < 6; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* lsqr_last_error(const lsqr_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int lsqr_ctx_set_stream(lsqr_ctx* ctx, void* cuda_stream) {
  if (!ctx) return LSQR_ERR_ARG;
  if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return LSQR_OK;
}

uint64_t lsqr_kernel_launches(const lsqr_ctx* ctx) { return ctx ? ctx->launches : 0; }

int lsqr_set_estimator(lsqr_ctx* ctx, int model, double delta, double aux, int ls_type) {
  if (!ctx) return LSQR_ERR_ARG;
  if (model < 0 || model >= LSQR_NUM_MODELS) return fail(ctx, LSQR_ERR_ARG, "unknown model");
  if (ls_type != LSQR_LS_ALGEBRAIC && ls_type != LSQR_LS_GEOMETRIC) return fail(ctx, LSQR_ERR_ARG, "bad least-squares type");  // SphereParametersEstimator.hxx:17-18
  if (ctx->model != model) {  // layouts differ between models: the data must be uploaded again
    for (DataSet* ds : {&ctx->main, &ctx->scratch}) { cudaFree(ds->soa64); cudaFree(ds->soa32); cudaFree(ds->maskbits); *ds = DataSet(); }
  }
  ctx->model = model; ctx->delta = delta; ctx->aux = aux; ctx->ls_type = ls_type;
  // The reference keeps delta*delta for every estimator but the hypersphere, pivot and dense-system ones (e.g.
  // PlaneParametersEstimator.hxx:16 vs SphereParametersEstimator.hxx:20), so the sign of delta is immaterial there; the device
  // code that compares an unsquared residual (fp32 fast mode) gets |delta|.
  const bool distance_threshold = model == LSQR_CIRCLE2 || model == LSQR_SPHERE3 || model == LSQR_SPHERE4 || model == LSQR_PIVOT ||
                                  model == LSQR_DENSE5 || model == LSQR_DENSE6;
  ctx->cfg.delta = distance_threshold ? delta : fabs(delta);
  ctx->cfg.delta2 = delta * delta;
  const double ang = aux > 0 ? aux : 0.017453292519943295769236907684886;  // RayIntersectionParametersEstimator.h:35
  double ce = sin(ang);
  ce *= ce;
  ctx->cfg.cross_eps = ce;
  ctx->main.moments_valid = false;
  return LSQR_OK;
}

int lsqr_upload(lsqr_ctx* ctx, const void* aos, size_t n, size_t stride_bytes) {
  if (!ctx) return LSQR_ERR_ARG;
  return upload_to(ctx, ctx->main, aos, n, stride_bytes, false);
}
int lsqr_upload_device(lsqr_ctx* ctx, const double* dev_packed, size_t n) {
  if (!ctx) return LSQR_ERR_ARG;
  if (ctx->model < 0) return fail(ctx, LSQR_ERR_STATE, "lsqr_set_estimator must be called first");
  return upload_to(ctx, ctx->main, dev_packed, n, sizeof(double) * model_info(ctx->model).D, true);
}

int lsqr_set_shard(lsqr_ctx* ctx, int rank, int world, lsqr_allreduce_max_u64_fn max_fn, lsqr_allreduce_sum_f64_fn sum_fn, void* user) {
  if (!ctx) return LSQR_ERR_ARG;
  if (world < 1 || rank < 0 || rank >= world) return fail(ctx, LSQR_ERR_ARG, "bad rank/world");
  if (world > 1 && (!max_fn || !sum_fn)) return fail(ctx, LSQR_ERR_ARG, "hooks required when world > 1");
  ctx->rank = rank; ctx->world = world; ctx->max_fn = max_fn; ctx->sum_fn = sum_fn; ctx->comm_user = user;
  return LSQR_OK;
}

int lsqr_score(lsqr_ctx* ctx, const lsqr_score_args* args, lsqr_score_result* res) {
  if (!ctx) return LSQR_ERR_ARG;
  return score_impl(ctx, args, res);
}

int lsqr_consensus(lsqr_ctx* ctx, const double* params, uint32_t* out_count) {
  if (!ctx || !params) return LSQR_ERR_ARG;
  return consensus_impl(ctx, ctx->main, params, out_count);
}
int lsqr_get_mask(lsqr_ctx* ctx, uint8_t* out_bytes) {
  if (!ctx) return LSQR_ERR_ARG;
  return get_mask_impl(ctx, ctx->main, out_bytes);
}
int lsqr_refine(lsqr_ctx* ctx, int use_mask, double* out_params, int* n_params) {
  if (!ctx) return LSQR_ERR_ARG;
  return refine_impl(ctx, ctx->main, use_mask, out_params, n_params);
}

int lsqr_ransac(lsqr_ctx* ctx, double prob, int precision, uint64_t seed, uint8_t* out_mask, lsqr_compute_result* res) {
  if (!ctx || !res) return LSQR_ERR_ARG;
  if (ctx->model < 0 || !ctx->main.soa64) return fail(ctx, LSQR_ERR_STATE, "set the estimator and upload data first");
  memset(res, 0, sizeof(*res));
  const ModelInfo mi = model_info(ctx->model);
  const uint32_t n = ctx->main.n;
  // RANSAC.hxx:16-19: fewer data than the minimal subset, or probability outside (0,1) -> return 0
  if (n < (uint32_t)mi.K || prob >= 1.0 || prob <= 0.0) return LSQR_OK;
  const double numerator = log(1.0 - prob);
  const unsigned int all_tries = choose_ref(n, (unsigned)mi.K);  // RANSAC.hxx:41
  // rounds of 256, 1024, 4096 ... hypotheses: the stop rule (:107-110) is re-evaluated between rounds, and a typical problem
  // (inlier ratio ~0.5, k = 3: 65 tries) ends after the first, small one
  uint64_t num_tries = all_tries, done = 0, round = 256;
  lsqr_score_result best{};
  double dev_ms = 0;
  while (done < num_tries) {
    lsqr_score_args a{};
    a.sampler = LSQR_SAMPLE_PHILOX; a.precision = precision; a.seed = seed; a.first = done;
    a.count = std::min<uint64_t>(round, num_tries - done);
    lsqr_score_result r{};
    if (int rc = score_impl(ctx, &a, &r)) return rc;
    dev_ms += r.score_ms;
    done += a.count;
    if (r.best_count > best.best_count) {  // strict '>' : RANSAC.hxx:100
      best = r;
      if (best.best_count == n) break;   // :104-105
      const double denominator = log(1.0 - pow((double)best.best_count / (double)n, (double)mi.K));  // :107
      const double t = numerator / denominator + 0.5;
      uint64_t nt = (t >= 4294967295.0 || !(t == t)) ? 0xFFFFFFFFull : (t < 0 ? 0x80000000ull : (uint64_t)t);  // (int) cast semantics of :108
      num_tries = std::min<uint64_t>(nt, all_tries);  // :110
    }
    round = std::min<uint64_t>(round * 4, (uint64_t)1 << 20);
  }
  res->tries = done;
  int rc = finish_ransac(ctx, best, out_mask, res);
  res->device_ms = dev_ms + ctx->refine_kernel_ms;
  return rc;
}

int lsqr_ransac_exhaustive(lsqr_ctx* ctx, int precision, uint8_t* out_mask, lsqr_compute_result* res) {
  if (!ctx || !res) return LSQR_ERR_ARG;
  if (ctx->model < 0 || !ctx->main.soa64) return fail(ctx, LSQR_ERR_STATE, "set the estimator and upload data first");
  memset(res, 0, sizeof(*res));
  const ModelInfo mi = model_info(ctx->model);
  const uint32_t n = ctx->main.n;
  if (n < (uint32_t)mi.K) return LSQR_OK;  // RANSAC.hxx:168-169
  const uint64_t total = choose_exact(n, mi.K);
  if (total > 0xFFFFFFFEull) return fail(ctx, LSQR_ERR_ARG, "C(N,k) exceeds 2^32-2: the brute-force overload is for small problems (RANSAC.h:107-109)");
  lsqr_score_args a{};
  a.sampler = LSQR_SAMPLE_EXHAUSTIVE; a.precision = precision; a.first = 0; a.count = total;
  lsqr_score_result r{};
  if (int rc = score_impl(ctx, &a, &r)) return rc;
  res->tries = total;
  int rc = finish_ransac(ctx, r, out_mask, res);
  res->device_ms = r.score_ms + ctx->refine_kernel_ms;
  return rc;
}

int lsqr_ransac_batch(lsqr_ctx* ctx, const double* data, const uint64_t* offsets, uint64_t n_problems, int exhaustive, double prob,
                      uint32_t max_tries, uint64_t seed, double* out_params, uint32_t* out_counts, uint8_t* out_masks, double* device_ms) {
  if (!ctx || !data || !offsets || !out_params || !out_counts) return LSQR_ERR_ARG;
  if (ctx->model < 0) return fail(ctx, LSQR_ERR_STATE, "lsqr_set_estimator must be called first");
  if (n_problems == 0) return LSQR_OK;
  if (n_problems > 0x7FFFFFFFull) return fail(ctx, LSQR_ERR_ARG, "too many problems");
  const ModelInfo mi = model_info(ctx->model);
  CK(cudaSetDevice(ctx->device));
  const uint64_t total = offsets[n_problems];
  uint32_t max_n = 1;
  for (uint64_t b = 0; b < n_problems; b++) {
    if (offsets[b + 1] < offsets[b]) return fail(ctx, LSQR_ERR_ARG, "offsets must be non-decreasing");
    max_n = std::max<uint64_t>(max_n, offsets[b + 1] - offsets[b]);
  }
  if ((size_t)max_n * mi.D * sizeof(double) > 200 * 1024) return fail(ctx, LSQR_ERR_ARG, "a problem does not fit in shared memory; use lsqr_ransac for large problems");
  cudaStream_t s = ctx->stream;
  // persistent buffers (a cudaMalloc/cudaFree set per call cost more than the kernel)
  if (int rc = ensure(ctx, &ctx->bt_data, &ctx->bt_data_cap, (size_t)std::max<uint64_t>(total, 1) * mi.D)) return rc;
  if (int rc = ensure(ctx, &ctx->bt_off, &ctx->bt_off_cap, (size_t)n_problems + 1)) return rc;
  if (int rc = ensure(ctx, &ctx->bt_prm, &ctx->bt_prm_cap, (size_t)n_problems * mi.P)) return rc;
  if (int rc = ensure(ctx, &ctx->bt_cnt, &ctx->bt_cnt_cap, (size_t)n_problems)) return rc;
  if (out_masks) if (int rc = ensure(ctx, &ctx->mask_dev, &ctx->mask_dev_cap, (size_t)std::max<uint64_t>(total, 1))) return rc;
  double* d_data = ctx->bt_data; uint64_t* d_off = ctx->bt_off; double* d_prm = ctx->bt_prm; uint32_t* d_cnt = ctx->bt_cnt;
  uint8_t* d_mask = out_masks ? ctx->mask_dev : nullptr;
  auto cleanup = [&]() {};
#define CKB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return fail(ctx, LSQR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)
  CKB(cudaMemcpyAsync(d_data, data, sizeof(double) * total * mi.D, cudaMemcpyHostToDevice, s));
  CKB(cudaMemcpyAsync(d_off, offsets, sizeof(uint64_t) * (n_problems + 1), cudaMemcpyHostToDevice, s));
  BatchArgs ba{};
  ba.model = ctx->model; ba.exhaustive = exhaustive; ba.tries = max_tries; ba.prob = prob; ba.seed = seed;
  ba.data = d_data; ba.offsets = d_off; ba.n_problems = (uint32_t)n_problems; ba.max_n = max_n;
  ba.out_params = d_prm; ba.out_counts = d_cnt; ba.out_masks = d_mask;
  CKB(cudaEventRecord(ctx->ev[0], s));
  if (launch_batch(ba, ctx->cfg, ctx->ls_type, s) < 0) { cleanup(); return fail(ctx, LSQR_ERR_ARG, "batch launch rejected"); }
  ctx->launches++;
  CKB(cudaEventRecord(ctx->ev[1], s));
  CKB(cudaGetLastError());
  CKB(cudaMemcpyAsync(out_params, d_prm, sizeof(double) * n_problems * mi.P, cudaMemcpyDeviceToHost, s));
  CKB(cudaMemcpyAsync(out_counts, d_cnt, sizeof(uint32_t) * n_problems, cudaMemcpyDeviceToHost, s));
  if (out_masks) CKB(cudaMemcpyAsync(out_masks, d_mask, total, cudaMemcpyDeviceToHost, s));
  CKB(cudaStreamSynchronize(s));
  float ms = 0.f;
  CKB(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  if (device_ms) *device_ms = ms;
  cleanup();
#undef CKB
  return LSQR_OK;
}

int lsqr_estimate(lsqr_ctx* ctx, const double* packed, size_t n, double* out_params, int* n_params) {
  if (!ctx || !packed || !out_params || !n_params) return LSQR_ERR_ARG;
  if (ctx->model < 0) return fail(ctx, LSQR_ERR_STATE, "lsqr_set_estimator must be called first");
  const ModelInfo mi = model_info(ctx->model);
  *n_params = 0;
  if (n < (size_t)mi.K) return LSQR_OK;  // e.g. PlaneParametersEstimator.hxx:45-46
  if ((ctx->model == USXW || ctx->model == USCP) && n != (size_t)mi.K) return LSQR_OK;  // SinglePointTargetUSCalibrationParametersEstimator.cxx:21-22, :674-675: exactly k
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  double* in_dev = ctx->small_dev + kSmEst;  // K*D <= 64 doubles
  CK(cudaMemcpyAsync(in_dev, packed, sizeof(double) * mi.K * mi.D, cudaMemcpyHostToDevice, s));
  launch_estimate_one(ctx->model, in_dev, ctx->cfg, ctx->small_dev + kSmOut, s); ctx->launches++;
  CKL();
  CK(cudaMemcpyAsync(ctx->pin, ctx->small_dev + kSmOut, sizeof(double) * (1 + LSQR_MAX_PARAMS), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  *n_params = (int)ctx->pin[0];
  for (int j = 0; j < *n_params; j++) out_params[j] = ctx->pin[1 + j];
  return LSQR_OK;
}

int lsqr_agree(lsqr_ctx* ctx, const double* params, const double* packed, size_t n, uint8_t* out) {
  if (!ctx || !params || !packed || !out) return LSQR_ERR_ARG;
  if (ctx->model < 0) return fail(ctx, LSQR_ERR_STATE, "lsqr_set_estimator must be called first");
  if (n == 0) return LSQR_OK;
  if (n > 0xFFFFFFF0ull) return fail(ctx, LSQR_ERR_ARG, "too many data records");
  const ModelInfo mi = model_info(ctx->model);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (int rc = ensure(ctx, &ctx->staging, &ctx->staging_cap, n * mi.D * sizeof(double) + n)) return rc;
  double* d_in = reinterpret_cast<double*>(ctx->staging);
  uint8_t* d_out = ctx->staging + n * mi.D * sizeof(double);
  CK(cudaMemcpyAsync(d_in, packed, sizeof(double) * n * mi.D, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->small_dev + kSmIn, params, sizeof(double) * mi.P, cudaMemcpyHostToDevice, s));
  launch_agree_many(ctx->model, ctx->small_dev + kSmIn, d_in, (uint32_t)n, ctx->cfg, d_out, s); ctx->launches++;
  CKL();
  CK(cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return LSQR_OK;
}

int lsqr_least_squares(lsqr_ctx* ctx, const double* packed, size_t n, double* out_params, int* n_params) {
  if (!ctx || !out_params || !n_params) return LSQR_ERR_ARG;
  if (ctx->model < 0) return fail(ctx, LSQR_ERR_STATE, "lsqr_set_estimator must be called first");
  const ModelInfo mi = model_info(ctx->model);
  *n_params = 0;
  if (n < (size_t)mi.K && ctx->model != RAY) return LSQR_OK;  // size guards, e.g. PlaneParametersEstimator.hxx:133-134
  if (n == 0) return LSQR_OK;
  // single-GPU by construction: a direct estimator call is not part of a sharded compute()
  const int world = ctx->world, rank = ctx->rank;
  ctx->world = 1; ctx->rank = 0;
  int rc = upload_to(ctx, ctx->scratch, packed, n, sizeof(double) * mi.D, false);
  if (!rc) rc = refine_impl(ctx, ctx->scratch, 0, out_params, n_params);
  ctx->world = world; ctx->rank = rank;
  return rc;
}

int lsqr_weighted_least_squares(lsqr_ctx* ctx, const double* packed, size_t n, const double* weights, double* out_params, int* n_params) {
  if (!ctx || !out_params || !n_params || !weights) return LSQR_ERR_ARG;
  if (ctx->model != ABSOR) return fail(ctx, LSQR_ERR_ARG, "weighted least squares exists for the absolute-orientation estimator only");
  *n_params = 0;
  if (n < 3) return LSQR_OK;   // AbsoluteOrientationParametersEstimator.cxx:214-215
  const int world = ctx->world, rank = ctx->rank;
  ctx->world = 1; ctx->rank = 0;
  int rc = upload_to(ctx, ctx->scratch, packed, n, sizeof(double) * 6, false);
  ctx->world = world; ctx->rank = rank;
  if (rc) return rc;
  cudaStream_t s = ctx->stream;
  if (int rc2 = ensure(ctx, &ctx->weights_dev, &ctx->weights_cap, n)) return rc2;
  CK(cudaMemcpyAsync(ctx->weights_dev, weights, sizeof(double) * n, cudaMemcpyHostToDevice, s));
  const DataView dv = ctx->scratch.view();
  launch_weighted_absor_moments(dv, ctx->weights_dev, ctx->rb, s);
  launch_reduce_partials(ctx->rb, 16, s);
  launch_solve_weighted_absor(dv, ctx->rb.moments, ctx->small_dev + kSmOut, s); ctx->launches += 3;
  CKL();
  CK(cudaMemcpyAsync(ctx->pin, ctx->small_dev + kSmOut, sizeof(double) * (1 + LSQR_MAX_PARAMS), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const int np = (int)ctx->pin[0];
  for (int j = 0; j < np; j++) out_params[j] = ctx->pin[1 + j];
  *n_params = np;
  ctx->scratch.moments_valid = false;
  return LSQR_OK;
}

int lsqr_microbench_fma(lsqr_ctx* ctx, int kind, int iters, double* out_fma_per_s, double* out_ms) {
  if (!ctx || kind < 0 || kind > 2 || iters <= 0) return LSQR_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  float* sink = reinterpret_cast<float*>(ctx->small_dev + kSmSink);
  const int blocks = ctx->num_sms * 8, threads = 256;
  launch_fma_bench(kind, 16, blocks, threads, sink, s);  // warm-up
  CK(cudaEventRecord(ctx->ev[0], s));
  launch_fma_bench(kind, iters, blocks, threads, sink, s); ctx->launches += 2;
  CK(cudaEventRecord(ctx->ev[1], s));
  CKL();
  CK(cudaStreamSynchronize(s));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  const double per_thread = (kind == 0 ? 16.0 : kind == 1 ? 32.0 : 8.0) * 8.0 * (double)iters;
  if (out_ms) *out_ms = ms;
  if (out_fma_per_s) *out_fma_per_s = per_thread * (double)blocks * threads / (ms * 1e-3);
  return LSQR_OK;
}

int lsqr_last_refine_stats(const lsqr_ctx* ctx, double* kernel_ms, double* algorithmic_bytes, int* lm_iterations) {
  if (!ctx) return LSQR_ERR_ARG;
  if (kernel_ms) *kernel_ms = ctx->refine_kernel_ms;
  if (algorithmic_bytes) *algorithmic_bytes = ctx->refine_bytes;
  if (lm_iterations) *lm_iterations = ctx->lm_iterations;
  return LSQR_OK;
}

}  // extern "C"
