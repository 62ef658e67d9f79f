// fp32 "fast mode" consensus: the hot kernel of the whole path (RANSAC.hxx:94-99 / :239-244).
//
// Formulation (chosen by measurement: tools/ptx_lab, tools/consensus_lab*.cu, profiles/r01_ptx_lab_rf_model.txt):
//   * An sm_100a SM sub-partition issues one warp instruction per cycle; a packed fma.rn.f32x2 (SASS FFMA2) holds both
//     halves of the FMA pipe for two cycles, an ALU instruction the 16-lane ALU pipe for two, and the register file delivers
//     two 32-bit operands per cycle.  The cost of an evaluation is therefore the FFMA2s of its residual plus whatever the
//     threshold test and the count cost in issue slots, ALU cycles and register reads.
//   * Arithmetic is packed over PAIRS OF POINTS (f32x2); the hypothesis constants are hoisted once per hypothesis in fp64
//     (hoist32_kernel) so that the residual is a bare FMA chain.
//   * Large requests (H >= kCbMinHyps) take the CONSTANT-BANK kernel (consensus_cb_kernel): the points of one launch sit in
//     the 64 KB constant bank, a point pair is a UNIFORM-register operand (LDCU -> `FFMA2 R, R.F32, UR.F32x2, R.F32x2`) and an
//     FFMA2 reads 2-3 registers instead of the 5-6 it reads when the pair comes out of shared memory.
//   * Small requests (adaptive RANSAC rounds, tests) take consensus32_kernel: point tiles staged by TMA bulk copies into a
//     2-deep shared-memory ring (as in k_score.cu), pairs through LDS.128.
//   * Counting, three ALU instructions per two residuals in the hot kernel:
//       - hyperplanes, dense systems, hypersphere family: CARRY CHAIN (count_carry) -- the residual arrives shifted (the hoisted
//         constant carries +delta; hypersphere: t' = d^2 - (r - delta)^2), inlier <=> bits(s') < bits(window) unsigned (window 2 delta,
//         or the hypothesis' own 4 r delta); two carry-only IADD3 and one IADD3.X with the two predicates as carry-ins, one
//         register operand each;
//       - 2-D lines (Line2D, Line<2>): two FFMA2 per pair leave the ALU pipe as the limit of any 3-ALU count, so of every four
//         point pairs two are tested as |s| < delta (two FSET.BF and one three-input IADD3 on their raw words,
//         count_abs_lt4_raw, decoded by cb_raw_decode) and two through the sign of s^2 - delta^2 (one more FFMA2, LEA.HI):
//         5 + 5 pipe slots per two pairs instead of 4 + 6;
//       - vector residuals (lines of 3 and more dimensions, absolute orientation, ray, pivot, ultrasound): the FMA chain starts
//         at -delta^2, so the SIGN BIT of the sum is the decision and LEA.HI adds it (count_sign).
//     Both kernels evaluate the same predicate per model, so their counts are bit-identical (tested).
// Tensor cores are deliberately unused: contraction depth <= 8 (BASELINE.json north_star: <= 4 for its estimators), and
// TF32 / bf16 operands would destroy a residual of 0.5 on coordinates of 1000.
#include "engine.h"
#include "fast_forms.cuh"

#include <mutex>

namespace lsqr {

// ---- mbarrier / TMA-bulk helpers ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarrier_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count)); }
__device__ __forceinline__ void mbarrier_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarrier_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src), "r"(bytes),
               "r"(smem_addr(bar))
               : "memory");
}

template <int M>
__global__ void hoist32_kernel(const double* __restrict__ hyp64, size_t hld, uint32_t H, DataView dv, EstCfg cfg, float* __restrict__ hyp32) {
  constexpr int P = Model<M>::P, Q = Model<M>::Q32;
  const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  double prm[P], c[kMaxDim];
  float q[Q];
#pragma unroll
  for (int j = 0; j < P; j++) prm[j] = hyp64[(size_t)j * hld + h];
#pragma unroll
  for (int j = 0; j < kMaxDim; j++) c[j] = dv.center[j];
  hoist32<M>(prm, c, cfg, q);
  const bool ok = prm[0] == prm[0];
  // layout: groups of four constants, [ceil(Q/4)][hld] float4 -> one 16-byte load per group in the consensus kernels
  float4* out = reinterpret_cast<float4*>(hyp32);
#pragma unroll
  for (int g = 0; g < (Q + 3) / 4; g++) {
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; c++) v[c] = (ok && 4 * g + c < Q) ? q[4 * g + c] : ((4 * g + c == thr_slot<M>()) ? 0.0f : __int_as_float(0x7fc00000));   // degenerate: NaN constants, empty window
    out[(size_t)g * hld + h] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// Reads the Q hoisted constants of hypothesis h (NaN for h >= H: never agrees).
template <int Q>
__device__ __forceinline__ void load_hyp32(const float* __restrict__ hyp, size_t hld, uint32_t h, uint32_t H, float* q) {
  const float4* in = reinterpret_cast<const float4*>(hyp);
  const float nan = __int_as_float(0x7fc00000);
#pragma unroll
  for (int g = 0; g < (Q + 3) / 4; g++) {
    const float4 v = (h < H) ? __ldg(in + (size_t)g * hld + h) : make_float4(nan, nan, nan, nan);
    const float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int c = 0; c < 4; c++) if (4 * g + c < Q) q[4 * g + c] = t[c];
  }
}

void launch_hoist32(int model, const double* hyp64, size_t hld, uint32_t H, const DataView& dv, const EstCfg& cfg, float* hyp32, cudaStream_t s) {
  if (H == 0) return;
  const unsigned blocks = (H + 255) / 256;
#define CALL(MM) hoist32_kernel<MM><<<blocks, 256, 0, s>>>(hyp64, hld, H, dv, cfg, hyp32)
  LSQR_DISPATCH_MODEL(model, CALL)
#undef CALL
}

// PPI = point pairs per inner iteration (2 -> one LDS.128 per component, 1 -> one LDS.64)
template <int M, int R, int THREADS, int TILE, int PPI>
__global__ void __launch_bounds__(THREADS) consensus32_kernel(const float* __restrict__ soa, size_t ld, uint32_t tiles_total, uint32_t tiles_per_chunk,
                                                               const float* __restrict__ hyp, size_t hld, uint32_t H, float delta, float delta2,
                                                               uint32_t* __restrict__ counts) {
  constexpr int D = Model<M>::D, Q = Model<M>::Q32;
  constexpr uint32_t kTileBytes = (uint32_t)(TILE * sizeof(float));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tile0 = reinterpret_cast<float*>(smem_raw);
  float* tile1 = tile0 + D * TILE;
  __shared__ __align__(8) uint64_t bars[2];

  const int tid = threadIdx.x;
  const uint32_t hbase = blockIdx.x * (THREADS * R);
  f2 q[R][Q];
  uint32_t cnt[R];
  unsigned long long out[R];   // kShifted models: outliers in the high word (count_carry)
#pragma unroll
  for (int r = 0; r < R; r++) {
    const uint32_t h = hbase + r * THREADS + tid;
    float qs[Q];
    load_hyp32<Q>(hyp, hld, h, H, qs);
#pragma unroll
    for (int j = 0; j < Q; j++) q[r][j] = splat(qs[j]);
    cnt[r] = 0;
    out[r] = 0ull;
  }
  Thr2 thr;
  thr.delta = splat(delta); thr.neg_delta2 = splat(-delta2); thr.fdelta = delta;
  const uint32_t negk = 0u - __float_as_uint(2.0f * delta);   // 2^32 - bits(2 delta)

  const uint32_t t0 = blockIdx.y * tiles_per_chunk;
  const uint32_t t1 = min(t0 + tiles_per_chunk, tiles_total);
  if (tid == 0) { mbarrier_init(&bars[0], 1); mbarrier_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  auto issue = [&](uint32_t t, int buf) {
    float* dst = buf ? tile1 : tile0;
    mbarrier_expect_tx(&bars[buf], kTileBytes * D);
#pragma unroll
    for (int d = 0; d < D; d++) tma_load_1d(dst + d * TILE, soa + (size_t)d * ld + (size_t)t * TILE, kTileBytes, &bars[buf]);
  };
  if (tid == 0 && t0 < t1) issue(t0, 0);
  uint32_t phase0 = 0, phase1 = 0;
  for (uint32_t t = t0; t < t1; t++) {
    const int buf = (t - t0) & 1;
    if (tid == 0 && t + 1 < t1) issue(t + 1, buf ^ 1);
    if (buf) { mbarrier_wait(&bars[1], phase1); phase1 ^= 1; } else { mbarrier_wait(&bars[0], phase0); phase0 ^= 1; }
    const float* tl = buf ? tile1 : tile0;
#pragma unroll 1
    for (int i = 0; i < TILE; i += 2 * PPI) {
      f2 x[PPI][D];
#pragma unroll
      for (int d = 0; d < D; d++) {
        if constexpr (PPI == 4) {
          const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(tl + d * TILE + i), w = *reinterpret_cast<const ulonglong2*>(tl + d * TILE + i + 4);
          x[0][d].v = v.x; x[1][d].v = v.y; x[2][d].v = w.x; x[3][d].v = w.y;
        } else if constexpr (PPI == 2) {
          const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(tl + d * TILE + i);
          x[0][d].v = v.x; x[1][d].v = v.y;
        } else {
          x[0][d].v = *reinterpret_cast<const unsigned long long*>(tl + d * TILE + i);
        }
      }
#pragma unroll
      for (int r = 0; r < R; r++) {
#pragma unroll
        for (int u = 0; u < PPI; u++) {
          // the predicate must be the constant-bank kernel's for every (hypothesis, datum): shifted models count through the
          // carry chain; the 2-D lines take |s| < delta for the first two point pairs of every group of four and the sign of
          // s^2 - delta^2 for the other two (kMixed, see consensus_cb_kernel); vector residuals the sign of the sum
          if constexpr (Eval<M>::kShifted) count_carry(out[r], Eval<M>::dist(q[r], x[u]), hyp_negk<M>(q[r], negk));
          else if (Eval<M>::kHasAbsForm && (u & 2) == 0) count_abs_lt(cnt[r], Eval<M>::dist(q[r], x[u]), hyp_thr<M>(q[r], thr.fdelta));
          else count_sign(cnt[r], Eval<M>::signed_(q[r], x[u], thr));
        }
      }
    }
    __syncthreads();  // everyone is done with this buffer before it is refilled
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    const uint32_t h = hbase + r * THREADS + tid;
    // NaN padding counts as outliers; delta = 0 (bits(2 delta) = 0: no carry ever) admits nothing, like |s| < 0
    if constexpr (Eval<M>::kShifted) cnt[r] = hyp_negk<M>(q[r], negk) ? (t1 - t0) * (uint32_t)TILE - (uint32_t)(out[r] >> 32) : 0u;
    if (h < H && cnt[r]) atomicAdd(&counts[h], cnt[r]);
  }
}

template <int M, int R, int THREADS, int TILE, int PPI>
static int run_consensus32(const DataView& dv, const float* hyp, size_t hld, uint32_t H, const EstCfg& cfg, uint32_t* counts, int num_sms, cudaStream_t s) {
  const uint32_t tiles_total = (uint32_t)(dv.span / TILE);
  const uint32_t hyp_blocks = (H + THREADS * R - 1) / (THREADS * R);
  uint32_t want = (uint32_t)num_sms * 16;  // ~16 work items per SM for load balance
  uint32_t chunks = (want + hyp_blocks - 1) / hyp_blocks;
  uint32_t max_chunks = (tiles_total + 3) / 4;
  if (max_chunks == 0) max_chunks = 1;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks > 65535) chunks = 65535;
  if (chunks == 0) chunks = 1;
  const uint32_t tiles_per_chunk = (tiles_total + chunks - 1) / chunks;
  chunks = (tiles_total + tiles_per_chunk - 1) / tiles_per_chunk;
  const size_t smem = 2 * (size_t)Model<M>::D * TILE * sizeof(float);
  auto kern = consensus32_kernel<M, R, THREADS, TILE, PPI>;
  // the attribute is per device (a multi-device context launches the same kernel on every GPU of the process)
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set[dev & 63] = true; }
  kern<<<dim3(hyp_blocks, chunks), THREADS, smem, s>>>(dv.soa32, dv.ld, tiles_total, tiles_per_chunk, hyp, hld, H, (float)cfg.delta, (float)cfg.delta2, counts);
  return 1;
}


// ---- constant-bank variant ---------------------------------------------------------------------
// One launch scores ALL hypotheses of the request against the `cp` points whose components sit in
// the constant bank ([D][cp] floats, cp a multiple of 16, copied there device-to-device before the
// launch).  grid = (hypothesis blocks) x (point sub-chunks of `sub` points); every thread keeps R
// hypotheses as scalars (broadcast .F32 operands), the point pairs are uniform registers.
constexpr int kCbFloats = 16320;              // 65280 B of the 64 KB bank
constexpr uint32_t kCbMinHyps = 98304;        // below this the per-launch overhead outweighs the gain
__constant__ float4 c_tile[kCbFloats / 4];
static std::mutex g_cb_mutex[16];             // the bank is per device, not per context

#ifndef LSQR_CB_MINBLOCKS
#define LSQR_CB_MINBLOCKS 5
#endif
constexpr uint32_t kCbRawMaxSub = 496;       // <= 511 inliers per raw counter, multiple of 16
template <int M> constexpr uint32_t cb_points() { return (uint32_t)(kCbFloats / Model<M>::D / 16 * 16); }   // points per launch

template <int M, int R, int THREADS, int PPI>
__global__ void __launch_bounds__(THREADS, LSQR_CB_MINBLOCKS) consensus_cb_kernel(uint32_t npts, uint32_t sub, const float* __restrict__ hyp, size_t hld, uint32_t H,
                                                                float delta, float delta2, uint32_t* __restrict__ counts) {
  constexpr int D = Model<M>::D, Q = Model<M>::Q32;
  constexpr uint32_t cp = cb_points<M>();
  const int tid = threadIdx.x;
  const uint32_t hbase = blockIdx.x * (THREADS * R);
  float qf[R][Q];
  uint32_t cnt[R], sgn[R];     // sgn: the sign-counted half of the mixed form
  unsigned long long out[R];   // kShifted models: outliers in the high word (count_carry)
#pragma unroll
  for (int r = 0; r < R; r++) {
    load_hyp32<Q>(hyp, hld, hbase + r * THREADS + tid, H, qf[r]);
    cnt[r] = 0; sgn[r] = 0;
    out[r] = 0ull;
  }
  const uint32_t negk = 0u - __float_as_uint(2.0f * delta);   // 2^32 - bits(2 delta)
  Thr2 thr;
  thr.delta = splat(delta); thr.neg_delta2 = splat(-delta2); thr.fdelta = delta;
  // one induction variable in units of 2*PPI points; the component offsets are immediates (cp is a multiple of 16)
  const uint32_t p0 = blockIdx.y * sub, p1 = min(p0 + sub, npts);
  const int g0 = (int)(p0 / (2 * PPI)), g1 = (int)(p1 / (2 * PPI));
  // raw-sum counting (count_abs_lt4_raw) holds 9 bits: the host keeps sub <= kCbRawMaxSub points per CTA
  constexpr bool kCarry = Eval<M>::kShifted;
  constexpr bool kRaw = Eval<M>::kHasAbsForm && !kCarry;
  static_assert(!kRaw || PPI % 4 == 0, "the mixed count alternates in groups of four point pairs");
  // points of group g (2*PPI points) from the constant bank: uniform loads, the values live in uniform registers
  auto load_group = [&](f2 (&x)[PPI][D], int g) {
#pragma unroll
    for (int d = 0; d < D; d++) {
      if constexpr (PPI % 2 == 0) {
#pragma unroll
        for (int j = 0; j < PPI / 2; j++) {
          const float4 v = c_tile[d * (int)(cp / 4) + (PPI / 2) * g + j];
          x[2 * j][d] = join(v.x, v.y); x[2 * j + 1][d] = join(v.z, v.w);
        }
      } else {
        const float2 v = reinterpret_cast<const float2*>(c_tile)[d * (int)(cp / 2) + g];
        x[0][d] = join(v.x, v.y);
      }
    }
  };
  auto score_group = [&](const f2 (&x)[PPI][D]) {
#pragma unroll
    for (int r = 0; r < R; r++) {
      f2 q[Q];
#pragma unroll
      for (int j = 0; j < Q; j++) q[j] = splat(qf[r][j]);
      if constexpr (kCarry) {
        const uint32_t nk = hyp_negk<M>(qf[r], negk);
#pragma unroll
        for (int u = 0; u < PPI; u++) count_carry(out[r], Eval<M>::dist(q, x[u]), nk);
      } else if constexpr (kRaw) {
        // kMixed: of every four point pairs two are tested as |s| < delta (2 FFMA2 + 3 ALU instructions per pair: ALU-bound on
        // its own) and two through the sign of s^2 - delta^2 (3 FFMA2 + 2 ALU: FMA-bound on its own); together 5 + 5 pipe slots
        // per two pairs instead of 4 + 6
        const float th = hyp_thr<M>(qf[r], thr.fdelta);
#pragma unroll
        for (int u = 0; u < PPI; u += 4) {
          count_abs_lt4_raw(cnt[r], Eval<M>::dist(q, x[u]), Eval<M>::dist(q, x[u + 1]), th);
          count_sign(sgn[r], Eval<M>::signed_(q, x[u + 2], thr));
          count_sign(sgn[r], Eval<M>::signed_(q, x[u + 3], thr));
        }
      } else {
#pragma unroll
        for (int u = 0; u < PPI; u++) count_sign(cnt[r], Eval<M>::signed_(q, x[u], thr));
      }
    }
  };
#pragma unroll 1
  for (int g = g0; g < g1; g++) {
    f2 x[PPI][D];
    load_group(x, g);
    score_group(x);
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    const uint32_t h = hbase + r * THREADS + tid;
    // NaN padding counts as outliers; delta = 0 (bits(2 delta) = 0: no carry ever) admits nothing, like |s| < 0
    if constexpr (kCarry) cnt[r] = hyp_negk<M>(qf[r], negk) ? (uint32_t)(g1 - g0) * (2u * PPI) - (uint32_t)(out[r] >> 32) : 0u;
    if constexpr (kRaw) cnt[r] = cb_raw_decode(cnt[r]) + sgn[r];
    if (h < H && cnt[r]) atomicAdd(&counts[h], cnt[r]);
  }
}

// Hypotheses per thread (R) / point pairs per iteration (PPI) of the constant-bank kernel (128 threads), from the sweeps in
// profiles/r01_tune_cb_sweep_models.txt and profiles/r02_tune_cb_blocking.txt (the plane's 12 x 8 -- sixteen points per iteration --
// is 4.6 % faster than the 10 x 4 it replaced: half the loop overhead per point).  Two things decide them:
//   * few hypotheses per thread on a model with few operations per datum make the uniform constant loads (D per point pair)
//     the limit: 3-8x slower, see the plane4 / dense5 columns of the first sweep;
//   * ptxas fetches the data with LDCU into uniform registers only when a loaded pair has enough consumers in the loop body;
//     below that it falls back to indexed LDC.64 into ordinary registers, which an SM sub-partition sustains at about one
//     per 38 cycles (pivot at R = 6, PPI = 1: 12 LDC.64 per 90 FFMA2, 0.70 T evals/s; at R = 8, PPI = 2 the same body
//     uses LDCU).  tests/test_sass_cpu.py pins "no LDC c[0x3] in the loop" for every model.
#ifndef LSQR_CB_THREADS
#define LSQR_CB_THREADS 128
#endif
#ifndef LSQR_CB_R_PLANE
#define LSQR_CB_R_PLANE 12
#endif
#ifndef LSQR_CB_PPI_PLANE
#define LSQR_CB_PPI_PLANE 8
#endif
constexpr int cb_blocking(int m, bool want_r) {
#ifdef LSQR_CB_OVR_MODEL      // R&D builds (tools/build_variants.sh): override one model
  if (m == LSQR_CB_OVR_MODEL) return want_r ? LSQR_CB_OVR_R : LSQR_CB_OVR_PPI;
#endif
  int r = 8, ppi = 2;
  switch (m) {
    case PLANE3: r = LSQR_CB_R_PLANE; ppi = LSQR_CB_PPI_PLANE; break;
    case LINE2D: case LINE2: r = 12; ppi = 8; break;   // two counters per hypothesis (mixed count): 16 x 4 spills, 12 x 8 measured 2.5 % over 14 x 4
    case PLANE2: r = 16; ppi = 4; break;
    case LINE3: r = 8; ppi = 4; break;
    case DENSE2: r = 10; ppi = 8; break;
    case CIRCLE2: r = 12; ppi = 2; break;
    case SPHERE3: r = 10; ppi = 2; break;
    case ABSOR: r = 4; ppi = 4; break;
    case RAY: r = 8; ppi = 4; break;
    case PIVOT: r = 8; ppi = 2; break;
    case DENSE5: case DENSE6: case SPHERE4: case PLANE4: r = 8; ppi = 4; break;
    case USXW: case USCP: r = 4; ppi = 2; break;
    default:   // the wider template space: hypotheses per thread bounded by the registers their constants take
      switch (model_family(m)) {
        case FAM_PLANE: case FAM_DENSE: r = model_dim(m) <= 4 ? 10 : 8; ppi = model_dim(m) <= 6 ? 4 : 2; break;
        case FAM_SPHERE: r = 8; ppi = model_dim(m) <= 6 ? 4 : 2; break;
        case FAM_LINE: r = model_dim(m) <= 5 ? 6 : 4; ppi = 2; break;
        default: break;
      }
      break;
  }
  return want_r ? r : ppi;
}
template <int M> struct BlockCB { static constexpr int R = cb_blocking(M, true), PPI = cb_blocking(M, false); };

template <int M>
static int run_consensus_cb(const DataView& dv, const float* hyp, size_t hld, uint32_t H, const EstCfg& cfg, uint32_t* counts, int num_sms, cudaStream_t s) {
  constexpr int D = Model<M>::D, R = BlockCB<M>::R, PPI = BlockCB<M>::PPI, THREADS = LSQR_CB_THREADS;
  constexpr uint32_t cp = cb_points<M>();
  auto kern = consensus_cb_kernel<M, R, THREADS, PPI>;
  int dev = 0;
  cudaGetDevice(&dev);
  static int occ[16] = {0};
  if (!occ[dev & 15]) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, THREADS, 0) != cudaSuccess || o < 1) o = 4;
    occ[dev & 15] = o;
  }
  void* bank = nullptr;
  if (cudaGetSymbolAddress(&bank, c_tile) != cudaSuccess) return -1;
  const uint32_t hyp_blocks = (H + THREADS * R - 1) / (THREADS * R);
  const uint32_t slots = (uint32_t)num_sms * (uint32_t)occ[dev & 15];
  // sub-chunks per launch: fill the last wave (the launches of one request serialise on the bank)
  constexpr bool kRaw = Eval<M>::kHasAbsForm && !Eval<M>::kShifted;
  auto pick_sub = [&](uint32_t npts) {
    uint32_t best_sub = kRaw ? std::min(npts, kCbRawMaxSub) : npts; double best_eff = 0.0;
    for (uint32_t nsub = 1; nsub <= 64; nsub++) {
      uint32_t sub = ((npts + nsub - 1) / nsub + 15) / 16 * 16;
      if (kRaw && sub > kCbRawMaxSub) continue;
      if (sub < 128 && nsub > 1) break;
      const uint32_t ny = (npts + sub - 1) / sub;
      const double waves = (double)hyp_blocks * ny / slots;
      const double eff = waves / (double)(uint64_t)(waves + 0.999999);
      if (eff > best_eff + 0.01) { best_eff = eff; best_sub = sub; }
    }
    return best_sub;
  };
  std::lock_guard<std::mutex> lock(g_cb_mutex[dev & 15]);
  int launches = 0;
  for (size_t base = 0; base < dv.span; base += cp) {
    const uint32_t npts = (uint32_t)((dv.span - base < cp) ? (dv.span - base) : cp);   // span is a multiple of 1024, cp of 16
    if (base >= dv.n) break;                                                        // only NaN padding left
    if (cudaMemcpy2DAsync(bank, sizeof(float) * cp, dv.soa32 + base, sizeof(float) * dv.ld, sizeof(float) * npts, D, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return -1;
    const uint32_t sub = pick_sub(npts);
    kern<<<dim3(hyp_blocks, (npts + sub - 1) / sub), THREADS, 0, s>>>(npts, sub, hyp, hld, H, (float)cfg.delta, (float)cfg.delta2, counts);
    launches += 2;
  }
  // the bank is shared by every context on this device: finish before another request may refill it
  if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
  return launches;
}

// Register blocking per model: R hypotheses per thread sized so that 2*Q32*R duplicated constants
// plus the point pairs stay below ~128 registers at 256 threads.
constexpr int block32_r(int m) {
  switch (m) {
    case PLANE3: case LINE2D: return 9;
    case LINE2: case CIRCLE2: case SPHERE3: case RAY: case PLANE4: return 8;
    case ABSOR: case USXW: case USCP: return 3;
    case PIVOT: return 4;
    default: { const int q = model_info(m).Q32; return q <= 8 ? 6 : (q <= 10 ? 5 : (q <= 12 ? 4 : 3)); }
  }
}
template <int M> struct Block32 { static constexpr int R = block32_r(M), PPI = (M == ABSOR || M == RAY || M == PIVOT || M == USXW || M == USCP) ? 1 : ((M == LINE2D || M == LINE2) ? 4 : 2); };

int launch_consensus32(int model, const DataView& dv, const float* hyp32, size_t hld, uint32_t H, const EstCfg& cfg, uint32_t* counts, int num_sms,
                       cudaStream_t s) {
  if (H == 0 || dv.n == 0) return 0;
  if (H >= kCbMinHyps) {
#define CALL(MM) return run_consensus_cb<MM>(dv, hyp32, hld, H, cfg, counts, num_sms, s)
    LSQR_DISPATCH_MODEL(model, CALL)
#undef CALL
  }
#define CALL(MM)                                                                                                        \
  if (H <= 8192) return run_consensus32<MM, 1, 128, 512, Block32<MM>::PPI>(dv, hyp32, hld, H, cfg, counts, num_sms, s);  \
  return run_consensus32<MM, Block32<MM>::R, 256, 512, Block32<MM>::PPI>(dv, hyp32, hld, H, cfg, counts, num_sms, s)
  LSQR_DISPATCH_MODEL(model, CALL)
#undef CALL
  return 0;
}

}  // namespace lsqr
