// Consensus set + least-squares refine (RANSAC.hxx:129-138 and the estimators'
// leastSquaresEstimate bodies).
//
// One streaming pass per refine: every datum is read once from the fp64 SoA arrays, agree() is
// evaluated in reference arithmetic against the winning hypothesis, the consensus bit is written
// (one 32-bit word per warp via ballot) and -- in the same pass -- the estimator's least-squares
// moments are accumulated over the inliers in fp64.  The kernel is HBM-bound: D*8 bytes read per
// datum, 1 bit written.  Per-block partial sums are combined in a fixed order (reproducible), and a
// single-thread kernel turns the moments into parameters with small Jacobi eigen / pseudo-inverse
// solves.  The sphere's geometric fit runs Levenberg-Marquardt with one such pass per function
// evaluation (J^T J, J^T r and the cost in one sweep) and an on-device controller.
//
// All position components are accumulated relative to dv.center (shift-invariant scatter / normal
// equations), which keeps the fp64 sums well conditioned for coordinates far from the origin.
#include "engine.h"
#include "../../include/lsqr_b200.h"
#include "lm_minpack.cuh"

namespace lsqr {

// number of accumulated doubles per model (index 0 is always the inlier count)
__host__ __device__ inline int n_moments(int model, bool lm) {
  switch (model) {
    case PLANE3: case LINE3: return 10;
    case LINE2D: case LINE2: return 6;
    case CIRCLE2: return lm ? 11 : 9;
    case SPHERE3: return lm ? 16 : 14;
    case SPHERE4: return lm ? 22 : 20;
    case PLANE4: return 15;
    case ABSOR: return 16;
    case RAY: return 10;
    case PIVOT: return 22;
    case DENSE5: return 21;   // count, A^T A upper triangle (15), A^T b (5)
    case DENSE6: return 28;   // count, 21, 6
    case USXW: return lm ? 79 : 91;   // LM: count, J^T J (66), J^T e (11), cost; analytic: count, A^T A (78), A^T b (12)
    case USCP: return lm ? 46 : 55;   // LM: count, 36, 8, cost; analytic: count, 45, 9
  }
  return 0;
}
static_assert(kLmStateDoubles >= LM_SIZE, "engine.h: LM state buffer too small");

int moments_count(int model, bool lm) { return n_moments(model, lm); }

template <int M> struct Mom { static constexpr int N = 0, NLM = 0, NPLM = 1; };   // NPLM = parameters of the LM problem
template <> struct Mom<PLANE3>  { static constexpr int N = 10, NLM = 0, NPLM = 1; };
template <> struct Mom<LINE3>   { static constexpr int N = 10, NLM = 0, NPLM = 1; };
template <> struct Mom<LINE2D>  { static constexpr int N = 6,  NLM = 0, NPLM = 1; };
template <> struct Mom<LINE2>   { static constexpr int N = 6,  NLM = 0, NPLM = 1; };
template <> struct Mom<CIRCLE2> { static constexpr int N = 9,  NLM = 11, NPLM = 3; };
template <> struct Mom<SPHERE3> { static constexpr int N = 14, NLM = 16, NPLM = 4; };
template <> struct Mom<SPHERE4> { static constexpr int N = 20, NLM = 22, NPLM = 5; };
template <> struct Mom<PLANE4>  { static constexpr int N = 15, NLM = 0, NPLM = 1; };
template <> struct Mom<ABSOR>   { static constexpr int N = 16, NLM = 0, NPLM = 1; };
template <> struct Mom<RAY>     { static constexpr int N = 10, NLM = 0, NPLM = 1; };
template <> struct Mom<PIVOT>   { static constexpr int N = 22, NLM = 0, NPLM = 1; };
template <> struct Mom<DENSE5>  { static constexpr int N = 21, NLM = 0, NPLM = 1; };
template <> struct Mom<DENSE6>  { static constexpr int N = 28, NLM = 0, NPLM = 1; };
template <> struct Mom<USXW>    { static constexpr int N = 91, NLM = 79, NPLM = 11; };
template <> struct Mom<USCP>    { static constexpr int N = 55, NLM = 46, NPLM = 8; };

// q = centred datum.  acc[0] is the inlier count: the callers add it (the streaming pass feeds zeros for data outside the
// consensus set instead of branching, so only the count must know).
template <int DIM> __device__ __forceinline__ void acc_scatter(const double* q, double* acc) {
  int o = 1;
#pragma unroll
  for (int j = 0; j < DIM; j++) acc[o++] += q[j];
#pragma unroll
  for (int j = 0; j < DIM; j++)
#pragma unroll
    for (int k = j; k < DIM; k++) { acc[o] = fma(q[j], q[k], acc[o]); o++; }   // (this TU is built with -fmad=false for agree(); the sums may fuse)
}
template <int DIM> __device__ __forceinline__ void acc_sphere_alg(const double* q, double* acc) {
  double s = 0;
#pragma unroll
  for (int j = 0; j < DIM; j++) s = fma(q[j], q[j], s);
  acc_scatter<DIM>(q, acc);
  int o = 1 + DIM + DIM * (DIM + 1) / 2;
#pragma unroll
  for (int j = 0; j < DIM; j++) { acc[o] = fma(s, q[j], acc[o]); o++; }
  acc[o] += s;
}
// residual / Jacobian of SphereParametersEstimator.hxx:394-431 at x = (centre', r)
template <int DIM> __device__ __forceinline__ void acc_sphere_lm(const double* q, const double* x, double* acc) {
  double J[DIM + 1], s = 0;
#pragma unroll
  for (int j = 0; j < DIM; j++) s += (q[j] - x[j]) * (q[j] - x[j]);
  const double sv = sqrt(s), r = sv - x[DIM];
#pragma unroll
  for (int j = 0; j < DIM; j++) J[j] = (x[j] - q[j]) / sv;
  J[DIM] = -1.0;
  acc[0] += 1.0;
  int o = 1;
#pragma unroll
  for (int a = 0; a <= DIM; a++)
#pragma unroll
    for (int b = a; b <= DIM; b++) { acc[o] = fma(J[a], J[b], acc[o]); o++; }
#pragma unroll
  for (int a = 0; a <= DIM; a++) { acc[o] = fma(J[a], r, acc[o]); o++; }
  acc[o] = fma(r, r, acc[o]);
}

template <int M> __device__ __forceinline__ void accumulate(const double* q, double* acc);
template <> __device__ __forceinline__ void accumulate<PLANE3>(const double* q, double* acc) { acc_scatter<3>(q, acc); }
template <> __device__ __forceinline__ void accumulate<PLANE4>(const double* q, double* acc) { acc_scatter<4>(q, acc); }
template <> __device__ __forceinline__ void accumulate<LINE3>(const double* q, double* acc) { acc_scatter<3>(q, acc); }
template <> __device__ __forceinline__ void accumulate<LINE2D>(const double* q, double* acc) { acc_scatter<2>(q, acc); }
template <> __device__ __forceinline__ void accumulate<LINE2>(const double* q, double* acc) { acc_scatter<2>(q, acc); }
template <> __device__ __forceinline__ void accumulate<CIRCLE2>(const double* q, double* acc) { acc_sphere_alg<2>(q, acc); }
template <> __device__ __forceinline__ void accumulate<SPHERE3>(const double* q, double* acc) { acc_sphere_alg<3>(q, acc); }
template <> __device__ __forceinline__ void accumulate<SPHERE4>(const double* q, double* acc) { acc_sphere_alg<4>(q, acc); }
// AbsoluteOrientationParametersEstimator.cxx:134-166: sums of both point sets and of p1 p2^T
template <> __device__ __forceinline__ void accumulate<ABSOR>(const double* q, double* acc) {
#pragma unroll
  for (int j = 0; j < 6; j++) acc[1 + j] += q[j];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) acc[7 + r * 3 + c] = fma(q[r], q[3 + c], acc[7 + r * 3 + c]);
}
// RayIntersectionParametersEstimator.cxx:108-123
template <> __device__ __forceinline__ void accumulate<RAY>(const double* q, double* acc) {
  const double* n = q + 3;
  acc[1] += n[0] * n[0]; acc[2] += n[0] * n[1]; acc[3] += n[0] * n[2];
  acc[4] += n[1] * n[1]; acc[5] += n[1] * n[2]; acc[6] += n[2] * n[2];
  const double s = n[0] * q[0] + n[1] * q[1] + n[2] * q[2];
  acc[7] += q[0] - s * n[0]; acc[8] += q[1] - s * n[1]; acc[9] += q[2] - s * n[2];
}
// Normal equations of the rows [R | -I] x = -t (PivotCalibrationParametersEstimator.cxx:77-83)
template <> __device__ __forceinline__ void accumulate<PIVOT>(const double* q, double* acc) {
  int o = 1;
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = a; b < 3; b++) acc[o++] += q[a] * q[b] + q[3 + a] * q[3 + b] + q[6 + a] * q[6 + b];  // (R^T R)_{ab}
#pragma unroll
  for (int j = 0; j < 9; j++) acc[o++] += q[j];                                                          // sum R
#pragma unroll
  for (int a = 0; a < 3; a++) acc[o++] += q[a] * q[9] + q[3 + a] * q[10] + q[6 + a] * q[11];            // R^T t
#pragma unroll
  for (int a = 0; a < 3; a++) acc[o++] += q[9 + a];                                                      // sum t
}

// Normal equations of the rows a.x = b (DenseLinearEquationSystemParametersEstimator.hxx:64-96)
template <int N> __device__ __forceinline__ void acc_dense(const double* q, double* acc) {
  int o = 1;
#pragma unroll
  for (int a = 0; a < N; a++)
#pragma unroll
    for (int b = a; b < N; b++) { acc[o] = fma(q[a], q[b], acc[o]); o++; }
#pragma unroll
  for (int a = 0; a < N; a++) { acc[o] = fma(q[a], q[N], acc[o]); o++; }
}
template <> __device__ __forceinline__ void accumulate<DENSE5>(const double* q, double* acc) { acc_dense<5>(q, acc); }
template <> __device__ __forceinline__ void accumulate<DENSE6>(const double* q, double* acc) { acc_dense<6>(q, acc); }

// Normal equations of the rows [u R2, v R2, R2, -I] x = -t2 (SinglePointTargetUSCalibrationParametersEstimator.cxx:137-189)
template <> __device__ __forceinline__ void accumulate<USXW>(const double* q, double* acc) {
  const double u = q[12], v = q[13];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    double row[12];
#pragma unroll
    for (int c = 0; c < 3; c++) { row[c] = q[3 * r + c] * u; row[3 + c] = q[3 * r + c] * v; row[6 + c] = q[3 * r + c]; row[9 + c] = (c == r) ? -1.0 : 0.0; }
    const double b = -q[9 + r];
    int o = 1;
#pragma unroll
    for (int a = 0; a < 12; a++)
#pragma unroll
      for (int bb = a; bb < 12; bb++) acc[o++] += row[a] * row[bb];
#pragma unroll
    for (int a = 0; a < 12; a++) acc[o++] += row[a] * b;
  }
}
// Levenberg-Marquardt pass of the ultrasound calibrations.  The reference hands MINPACK the SCALAR residuals d_i = |e_i| with
// e_i = R2 (u m_x c1 + v m_y c2 + t3) + t2 - t1 (f, SinglePointTargetUSCalibrationParametersEstimator.cxx:415-507) and the
// Jacobian rows (e_i^T de_i/dx) / d_i (gradf, :510-658); the iteration path depends on that choice (J^T J of the scalar form is
// not the one of the vector form), so the same rows are accumulated here: count, J^T J (upper triangle), J^T d, sum d^2.
// x = [t1 (NT1 = 3, cross-wire only), t3, omega_z, omega_y, omega_x, m_x, m_y]; `tgt` is t1 or the measured pointer tip p.
template <int NT1> __device__ __forceinline__ void acc_us_lm_rows(const double* q, const double* x, const double* tgt, double* acc) {
  constexpr int NP = NT1 + 8;
  const double* y = x + NT1;   // t3, angles, scales
  const double sz = sin(y[3]), cz = cos(y[3]), sy = sin(y[4]), cy = cos(y[4]), sx = sin(y[5]), cx = cos(y[5]);
  const double mx = y[6], my = y[7], u = q[12], v = q[13];
  const double c1[3] = {cz * cy, sz * cy, -sy};
  const double c2[3] = {cz * sy * sx - sz * cx, sz * sy * sx + cz * cx, cy * sx};
  const double dc1[3][3] = {{-sz * cy, cz * cy, 0}, {-cz * sy, -sz * sy, -cy}, {0, 0, 0}};
  const double dc2[3][3] = {{-sz * sy * sx - cz * cx, cz * sy * sx - sz * cx, 0}, {cz * cy * sx, sz * cy * sx, -sy * sx},
                            {cz * sy * cx + sz * sx, sz * sy * cx - cz * sx, cy * cx}};
  double w[3], dw[8][3];   // d w / d (t3 (3), omega (3), m_x, m_y)
#pragma unroll
  for (int k = 0; k < 3; k++) {
    w[k] = u * mx * c1[k] + v * my * c2[k] + y[k];
#pragma unroll
    for (int p = 0; p < 3; p++) { dw[p][k] = (p == k) ? 1.0 : 0.0; dw[3 + p][k] = u * mx * dc1[p][k] + v * my * dc2[p][k]; }
    dw[6][k] = u * c1[k];
    dw[7][k] = v * c2[k];
  }
  double e[3], J[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) J[p] = 0.0;
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const double a = q[3 * r], b = q[3 * r + 1], c = q[3 * r + 2];
    e[r] = a * w[0] + b * w[1] + c * w[2] + q[9 + r] - tgt[r];
    if (NT1) J[r] = -e[r];
#pragma unroll
    for (int p = 0; p < 8; p++) J[NT1 + p] += (a * dw[p][0] + b * dw[p][1] + c * dw[p][2]) * e[r];
  }
  const double d = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
#pragma unroll
  for (int p = 0; p < NP; p++) J[p] /= d;
  acc[0] += 1.0;
  int o = 1;
#pragma unroll
  for (int i = 0; i < NP; i++)
#pragma unroll
    for (int j = i; j < NP; j++) { acc[o] = fma(J[i], J[j], acc[o]); o++; }
#pragma unroll
  for (int i = 0; i < NP; i++) { acc[o] = fma(J[i], d, acc[o]); o++; }
  acc[o] = fma(d, d, acc[o]);
}
__device__ __forceinline__ void acc_us_lm(const double* q, const double* x, double* acc) { acc_us_lm_rows<3>(q, x, x, acc); }

// Normal equations of the rows [u R2, v R2, R2] x = p - t2 (SinglePointTargetUSCalibrationParametersEstimator.cxx:806-846)
template <> __device__ __forceinline__ void accumulate<USCP>(const double* q, double* acc) {
  const double u = q[12], v = q[13];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    double row[9];
#pragma unroll
    for (int c = 0; c < 3; c++) { row[c] = q[3 * r + c] * u; row[3 + c] = q[3 * r + c] * v; row[6 + c] = q[3 * r + c]; }
    const double b = q[14 + r] - q[9 + r];
    int o = 1;
#pragma unroll
    for (int a = 0; a < 9; a++)
#pragma unroll
      for (int bb = a; bb < 9; bb++) acc[o++] += row[a] * row[bb];
#pragma unroll
    for (int a = 0; a < 9; a++) acc[o++] += row[a] * b;
  }
}
// calibrated pointer: the same rows with the measured tip position p (q[14..16]) in the place of t1 and no unknown for it
__device__ __forceinline__ void acc_uscp_lm(const double* q, const double* x, double* acc) { acc_us_lm_rows<0>(q, x, q + 14, acc); }

__host__ __device__ inline bool centred_comp(int model, int d) {
  switch (model) {
    case RAY: return d < 3;
    case PIVOT: return d >= 9;
    case USXW: return d >= 9 && d < 12;
    case USCP: return false;
    case DENSE5: case DENSE6: return false;
    default: return true;
  }
}

// ---- mbarrier / TMA-bulk helpers (SASS: SYNCS / UBLKCP) ------------------------------------
__device__ __forceinline__ uint32_t mm_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mm_bar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mm_smem(bar)), "r"(count)); }
__device__ __forceinline__ void mm_bar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mm_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mm_bar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mm_smem(bar)) : "memory"); }
__device__ __forceinline__ void mm_bar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(mm_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mm_tma_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mm_smem(dst)), "l"(src), "r"(bytes),
               "r"(mm_smem(bar))
               : "memory");
}

// Blocking of the streaming pass: one persistent CTA per SM walks tiles of TILE data; a tile is D rows of TILE doubles in
// shared memory, filled by D bulk copies (cp.async.bulk, one mbarrier per stage), STAGES tiles deep.  The bytes in flight per
// SM are (STAGES - 1) tiles whatever the compute phase of the warps is -- what the register-staged version lacked
// (profiles/r01_ncu_mask_moments_v2.txt: 25 % warps active, wait + long_scoreboard 60 %, 0.62 of HBM).  Tiles of ~48 KB
// (2048 3-D points: four rows of 32 per thread, interleaved) x 4 stages measured best (profiles/r02_tune_mm.txt).
#ifndef LSQR_MM_SMEM_KB
#define LSQR_MM_SMEM_KB 208
#endif
#ifndef LSQR_MM_TILE_KB
#define LSQR_MM_TILE_KB 48
#endif
#ifndef LSQR_MM_THREADS
#define LSQR_MM_THREADS 512
#endif
template <int M, bool LM> struct MMCfg {
  static constexpr int D = Model<M>::D;
  static constexpr int NM = LM ? Mom<M>::NLM : Mom<M>::N;
  // models with more than 40 accumulators per thread (the two ultrasound calibrations: 46-91 doubles) get the whole register
  // file of an SM for 256 threads; everyone else runs 512 threads
  static constexpr int THREADS = NM > 40 ? 256 : LSQR_MM_THREADS;
  static constexpr int kTileCap = LSQR_MM_TILE_KB * 1024 + 8192;
  static constexpr int TILE = (D * 2048 * 8 <= kTileCap) ? 2048 : (D * 1024 * 8 <= kTileCap) ? 1024 : ((D * 512 * 8 <= kTileCap) ? 512 : 256);
  static constexpr int PPT = (TILE + THREADS - 1) / THREADS;                 // data per thread per tile
  static constexpr int TILE_BYTES = D * TILE * 8;
  static constexpr int kFit = LSQR_MM_SMEM_KB * 1024 / TILE_BYTES;
  static constexpr int STAGES = kFit >= 8 ? 8 : (kFit >= 2 ? kFit : 2);      // stage and phase are carried as counters, no division
  static constexpr size_t SMEM = (size_t)STAGES * TILE_BYTES;
  // light moment sets are accumulated without a branch (zeros for data outside the consensus set): the rows of a tile are
  // then independent straight-line chains that the scheduler interleaves; Levenberg-Marquardt rows (sqrt, division) and the
  // 46-91-term calibration rows stay behind a branch
  static constexpr bool BRANCHLESS = !LM && NM <= 32;
};
int mask_moments_ctas_per_sm() { return 1; }

// MODE 0: every datum counts; 1: evaluate agree() and write the consensus bits (and, when maskbytes != nullptr, the
// std::vector<bool>-shaped byte per datum that the caller of compute() receives); 2: read stored bits.
// HBM-bound: D*8 bytes read per datum, 1 bit (or 1 byte + 1 bit) written.
template <int M, int MODE, bool LM>
__global__ void __launch_bounds__(MMCfg<M, LM>::THREADS, 1) mask_moments_kernel(DataView dv, uint32_t begin, uint32_t end, const double* __restrict__ params_dev,
                                                               const double* __restrict__ lm_state, EstCfg cfg, uint32_t* __restrict__ maskbits,
                                                               uint8_t* __restrict__ maskbytes, double* __restrict__ partials) {
  using C = MMCfg<M, LM>;
  constexpr int D = Model<M>::D, P = Model<M>::P, HQ = Model<M>::HQ;
  constexpr int NM = C::NM, TILE = C::TILE, THREADS = C::THREADS, PPT = C::PPT, STAGES = C::STAGES;
  constexpr bool kFullTile = PPT * THREADS == TILE;
  extern __shared__ __align__(128) unsigned char mm_smem_raw[];
  double* ring = reinterpret_cast<double*>(mm_smem_raw);                     // [STAGES][D][TILE]
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
  __shared__ double sh[THREADS / 32][NM > 0 ? NM : 1];
  const uint32_t tid = threadIdx.x, lane = tid & 31;

  // LM passes are enqueued a few evaluations ahead of the controller's status word: once the iteration has stopped, the
  // remaining passes are no-ops (the controller ignores the stale moments)
  const bool active = !LM || lm_state[LM_STATUS] == 0.0;
  // tiles [t0, t1) in units of TILE data cover [begin, end); this CTA takes t0 + blockIdx.x, + gridDim.x, ...
  const uint32_t t0 = begin / TILE, t1 = (uint32_t)(((uint64_t)end + TILE - 1) / TILE);
  const uint32_t first = t0 + blockIdx.x;
  const uint32_t n_mine = (active && first < t1) ? (t1 - first + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](uint32_t k, uint32_t s) {   // tile number k of this CTA into stage s
    const size_t base = (size_t)(first + k * gridDim.x) * TILE;
    double* dst = ring + (size_t)s * D * TILE;
    mm_bar_expect_tx(&full_bar[s], (uint32_t)C::TILE_BYTES);
#pragma unroll
    for (int d = 0; d < D; d++) mm_tma_g2s(dst + d * TILE, dv.soa64 + (size_t)d * dv.ld + base, TILE * 8, &full_bar[s]);
  };
  // the first tiles are on their way before anything else is set up
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) { mm_bar_init(&full_bar[s], 1); mm_bar_init(&empty_bar[s], THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (uint32_t k = 0; k < (uint32_t)STAGES && k < n_mine; k++) issue(k, k);
  }

  double acc[NM > 0 ? NM : 1];
#pragma unroll
  for (int j = 0; j < NM; j++) acc[j] = 0.0;
  double hq[HQ];
  if (MODE == 1) {
    double prm[P];
#pragma unroll
    for (int j = 0; j < P; j++) prm[j] = params_dev[j];
    prepare<M>(prm, hq);
  }
  // the shift c of the centred components, once per thread (a load inside the loop would put a global-memory latency on every
  // row's critical path)
  double ctr[D];
#pragma unroll
  for (int d = 0; d < D; d++) ctr[d] = centred_comp(M, d) ? dv.center[d] : 0.0;
  double lmx[LM ? Mom<M>::NPLM : 1] = {0};
  if (LM) {
    // the phase selects the evaluation point: x or the trial point
    const int off = (lm_state[LM_PHASE] != 0.0) ? LM_TRIAL : LM_X;
#pragma unroll
    for (int j = 0; j < Mom<M>::NPLM; j++) lmx[j] = lm_state[off + j];
  }
  __syncthreads();   // barriers initialised before anyone waits on them

  // one tile: `checked` tiles straddle begin / end (at most two per launch) and test every row against the range
  auto process = [&](const double (&x)[PPT][D], uint32_t base, bool checked) {
    bool in[PPT], rowok[PPT];
#pragma unroll
    for (int u = 0; u < PPT; u++) {
      const uint32_t p = u * THREADS + tid, i = base + p, row = i - lane;      // rows of 32 data start on multiples of 32
      rowok[u] = (kFullTile || p < (uint32_t)TILE) && (!checked || (row >= begin && row < end));   // warp-uniform
      const bool valid = rowok[u] && (!checked || i < end);
      if (MODE == 0) in[u] = valid;
      else if (MODE == 1) in[u] = valid && agree<M>(hq, x[u], cfg);   // NaN padding beyond n never agrees
      else in[u] = valid && ((maskbits[row >> 5] >> lane) & 1u);
    }
    if (MODE == 1) {
#pragma unroll
      for (int u = 0; u < PPT; u++) {
        const uint32_t p = u * THREADS + tid, i = base + p;
        const unsigned bits = __ballot_sync(0xffffffffu, in[u]);
        if (rowok[u]) {
          if (lane == 0) maskbits[(i - lane) >> 5] = bits;
          if (maskbytes != nullptr && i < end) maskbytes[i] = in[u] ? 1 : 0;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PPT; u++) {
      if constexpr (C::BRANCHLESS) {
        double q[D];
#pragma unroll
        for (int d = 0; d < D; d++) q[d] = in[u] ? x[u][d] - ctr[d] : 0.0;
        acc[0] += in[u] ? 1.0 : 0.0;
        accumulate<M>(q, acc);
      } else if (in[u]) {
        double q[D];
#pragma unroll
        for (int d = 0; d < D; d++) q[d] = x[u][d] - ctr[d];
        if (LM) {
          if constexpr (M == CIRCLE2) acc_sphere_lm<2>(q, lmx, acc);
          if constexpr (M == SPHERE3) acc_sphere_lm<3>(q, lmx, acc);
          if constexpr (M == SPHERE4) acc_sphere_lm<4>(q, lmx, acc);
          if constexpr (M == USXW) acc_us_lm(q, lmx, acc);
          if constexpr (M == USCP) acc_uscp_lm(q, lmx, acc);
        } else { acc[0] += 1.0; accumulate<M>(q, acc); }
      }
    }
  };

  uint32_t s = 0, parity = 0;
  for (uint32_t k = 0; k < n_mine; k++) {
    const uint32_t base = (first + k * gridDim.x) * (uint32_t)TILE;
    const double* tile = ring + (size_t)s * D * TILE;
    mm_bar_wait(&full_bar[s], parity);
    double x[PPT][D];
#pragma unroll
    for (int u = 0; u < PPT; u++) {
      const uint32_t p = u * THREADS + tid;
      if (kFullTile || p < (uint32_t)TILE) {
#pragma unroll
        for (int d = 0; d < D; d++) x[u][d] = tile[d * TILE + p];
      }
    }
    // the stage is free as soon as every warp holds its data in registers; thread 0 refills it STAGES tiles ahead
    __syncwarp();
    if (lane == 0) mm_bar_arrive(&empty_bar[s]);
    if (tid == 0 && k + STAGES < n_mine) { mm_bar_wait(&empty_bar[s], parity); issue(k + STAGES, s); }
    const uint32_t s_next = (s + 1 == (uint32_t)STAGES) ? 0u : s + 1;
    parity ^= (s_next == 0u) ? 1u : 0u;
    s = s_next;
    const bool checked = base < begin || (uint64_t)base + TILE > end;   // CTA-uniform
    if (checked) process(x, base, true); else process(x, base, false);
  }
  if (LM && !active) return;   // partials, moments and the controller state keep their values
  // block reduction: shuffle within warps, shared memory across the warps
#pragma unroll
  for (int j = 0; j < NM; j++) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[tid >> 5][j] = v;
  }
  __syncthreads();
  if ((int)tid < NM) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) v += sh[w][tid];
    partials[(size_t)blockIdx.x * kMaxMoments + tid] = v;
  }
}

template <int M, int MODE, bool LM>
static void run_mask_moments(const DataView& dv, uint32_t begin, uint32_t end, const double* params_dev, const double* lm_state, const EstCfg& cfg,
                             const RefineBuffers& rb, cudaStream_t s) {
  using C = MMCfg<M, LM>;
  auto kern = mask_moments_kernel<M, MODE, LM>;
  static bool attr_set[64] = {false};   // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM); attr_set[dev & 63] = true; }
  kern<<<rb.blocks, C::THREADS, C::SMEM, s>>>(dv, begin, end, params_dev, lm_state, cfg, rb.maskbits, rb.maskbytes, rb.partials);
}

void launch_mask_moments(int model, const DataView& dv, uint32_t begin, uint32_t end, const double* params_dev, int mask_mode,
                         const double* lm_state, const EstCfg& cfg, const RefineBuffers& rb, cudaStream_t s) {
#define BYMODE(MM, LMV)                                                                           \
  if (mask_mode == 0) run_mask_moments<MM, 0, LMV>(dv, begin, end, params_dev, lm_state, cfg, rb, s);      \
  else if (mask_mode == 1) run_mask_moments<MM, 1, LMV>(dv, begin, end, params_dev, lm_state, cfg, rb, s); \
  else run_mask_moments<MM, 2, LMV>(dv, begin, end, params_dev, lm_state, cfg, rb, s)
  if (lm_state) {
    if (model == CIRCLE2) { BYMODE(CIRCLE2, true); }
    else if (model == SPHERE3) { BYMODE(SPHERE3, true); }
    else if (model == SPHERE4) { BYMODE(SPHERE4, true); }
    else if (model == USXW) { BYMODE(USXW, true); }
    else if (model == USCP) { BYMODE(USCP, true); }
    return;
  }
  switch (model) {
    case PLANE3: { BYMODE(PLANE3, false); break; }
    case LINE2D: { BYMODE(LINE2D, false); break; }
    case LINE2: { BYMODE(LINE2, false); break; }
    case LINE3: { BYMODE(LINE3, false); break; }
    case CIRCLE2: { BYMODE(CIRCLE2, false); break; }
    case SPHERE3: { BYMODE(SPHERE3, false); break; }
    case ABSOR: { BYMODE(ABSOR, false); break; }
    case RAY: { BYMODE(RAY, false); break; }
    case PIVOT: { BYMODE(PIVOT, false); break; }
    case DENSE5: { BYMODE(DENSE5, false); break; }
    case DENSE6: { BYMODE(DENSE6, false); break; }
    case USXW: { BYMODE(USXW, false); break; }
    case USCP: { BYMODE(USCP, false); break; }
    case SPHERE4: { BYMODE(SPHERE4, false); break; }
    case PLANE4: { BYMODE(PLANE4, false); break; }
  }
#undef BYMODE
}

// Weighted Horn moments (AbsoluteOrientationParametersEstimator.cxx:217-261): sum w, sum w p1, sum w p2,
// sum w p1 p2^T over all n pairs, relative to dv.center.  Same partials layout as mask_moments_kernel.
__global__ void __launch_bounds__(256) weighted_absor_moments_kernel(DataView dv, const double* __restrict__ w, double* __restrict__ partials) {
  constexpr int NM = 16;
  double acc[NM];
#pragma unroll
  for (int j = 0; j < NM; j++) acc[j] = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < dv.n; i += (size_t)gridDim.x * blockDim.x) {
    double q[6];
#pragma unroll
    for (int d = 0; d < 6; d++) q[d] = dv.soa64[(size_t)d * dv.ld + i] - dv.center[d];
    const double wi = w[i];
    acc[0] += wi;
#pragma unroll
    for (int j = 0; j < 6; j++) acc[1 + j] += q[j] * wi;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) acc[7 + r * 3 + c] += (q[r] * q[3 + c]) * wi;
  }
  __shared__ double sh[8][kMaxMoments];
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NM; j++) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[threadIdx.x >> 5][j] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x < NM) {
    double v = 0.0;
#pragma unroll
    for (int wp = 0; wp < 8; wp++) v += sh[wp][threadIdx.x];
    partials[(size_t)blockIdx.x * kMaxMoments + threadIdx.x] = v;
  }
}
void launch_weighted_absor_moments(const DataView& dv, const double* weights_dev, const RefineBuffers& rb, cudaStream_t s) {
  weighted_absor_moments_kernel<<<rb.blocks, 256, 0, s>>>(dv, weights_dev, rb.partials);
}

// One warp per moment; lanes stride over the blocks, then a fixed-order shuffle tree: the summation
// order depends only on the launch geometry, so results are reproducible run to run.
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int blocks, int nm, double* __restrict__ moments) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= nm) return;
  double v = 0.0;
  for (int b = lane; b < blocks; b += 32) v += partials[(size_t)b * kMaxMoments + j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) moments[j] = v;
}
void launch_reduce_partials(const RefineBuffers& rb, int nm, cudaStream_t s) {
  reduce_partials_kernel<<<(nm + 7) / 8, 256, 0, s>>>(rb.partials, rb.blocks, nm, rb.moments);
}

// ---------------------------------------------------------------------------------------
// moments -> parameters
// ---------------------------------------------------------------------------------------

// Symmetric positive semi-definite solve through the eigen-decomposition of the diagonally
// scaled matrix; eigenvalues below 1e-13 of the largest are dropped.  Returns the rank.
// Stands in for vnl_matrix_inverse on the normal equations of the reference's tall systems.
template <int N>
__device__ int sym_pinv_solve(const double* A, const double* b, double* x) {
  double S[N * N], V[N * N], ev[N], sc[N], y[N];
  for (int i = 0; i < N; i++) sc[i] = A[i * N + i] > 0 ? 1.0 / sqrt(A[i * N + i]) : 1.0;
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) S[i * N + j] = A[i * N + j] * sc[i] * sc[j];
  sym_eig<N>(S, V, ev);
  const double tol = 1e-13 * fabs(ev[N - 1]);
  int rank = 0;
  for (int k = 0; k < N; k++) {
    double d = 0;
    for (int i = 0; i < N; i++) d += V[i * N + k] * (b[i] * sc[i]);
    if (ev[k] > tol) { y[k] = d / ev[k]; rank++; } else y[k] = 0.0;
  }
  for (int i = 0; i < N; i++) { double s = 0; for (int k = 0; k < N; k++) s += V[i * N + k] * y[k]; x[i] = s * sc[i]; }
  return rank;
}

// PlaneParametersEstimator.hxx:141-171 (col 0) / LineParametersEstimator.hxx:80-110 (col DIM-1)
template <int DIM> __device__ int solve_scatter(const double* m, const double* c, int col, double* out) {
  const double n = m[0];
  if (n < 1.0) return 0;
  double C[DIM * DIM], V[DIM * DIM], ev[DIM];
  int o = 1 + DIM;
  for (int j = 0; j < DIM; j++) for (int k = j; k < DIM; k++) { const double v = m[o++] - m[1 + j] * m[1 + k] / n; C[j * DIM + k] = v; C[k * DIM + j] = v; }
  sym_eig<DIM>(C, V, ev);
  for (int j = 0; j < DIM; j++) out[j] = V[j * DIM + col];
  for (int j = 0; j < DIM; j++) out[DIM + j] = m[1 + j] / n + c[j];
  return 2 * DIM;
}
// Line2DParametersEstimator.cxx:50-100
__device__ int solve_line2d(const double* m, const double* c, double* out) {
  const double n = m[0];
  if (n < 1.0) return 0;
  const double mx = m[1] / n, my = m[2] / n;
  const double c11 = m[3] - n * mx * mx, c12 = m[4] - n * mx * my, c22 = m[5] - n * my * my;
  double nx, ny;
  if (c11 < 1e-12) {
    nx = 1.0; ny = 0.0;
    if (c22 < 1e-12) return 0;
  } else {
    const double lambda1 = (c11 + c22 + sqrt((c11 - c22) * (c11 - c22) + 4 * c12 * c12)) / 2.0;
    nx = -c12; ny = lambda1 - c22;
    const double norm = sqrt(nx * nx + ny * ny);
    nx /= norm; ny /= norm;
  }
  out[0] = nx; out[1] = ny; out[2] = mx + c[0]; out[3] = my + c[1];
  return 4;
}
// SphereParametersEstimator.hxx:267-307 through the normal equations of [-2p, 1] x = -|p|^2.
// Result in centred coordinates: out = (centre', r).
template <int DIM> __device__ int solve_sphere_alg(const double* m, double* out) {
  constexpr int NP = DIM + 1;
  const double n = m[0];
  if (n < (double)NP) return 0;
  double A[NP * NP], b[NP], x[NP];
  int o = 1 + DIM;
  for (int j = 0; j < DIM; j++) for (int k = j; k < DIM; k++) { const double v = 4.0 * m[o++]; A[j * NP + k] = v; A[k * NP + j] = v; }
  for (int j = 0; j < DIM; j++) { A[j * NP + DIM] = -2.0 * m[1 + j]; A[DIM * NP + j] = -2.0 * m[1 + j]; }
  A[DIM * NP + DIM] = n;
  for (int j = 0; j < DIM; j++) b[j] = 2.0 * m[o++];
  b[DIM] = -m[o];
  if (sym_pinv_solve<NP>(A, b, x) < NP) return 0;
  double r2 = -x[DIM];
  for (int j = 0; j < DIM; j++) { out[j] = x[j]; r2 += x[j] * x[j]; }
  if (!(r2 > 0)) return 0;
  out[DIM] = sqrt(r2);
  return NP;
}
// AbsoluteOrientationParametersEstimator.cxx:134-205 (Horn): M = sum p1 p2^T - N mu1 mu2^T, 4x4 N matrix,
// eigenvector of the largest eigenvalue, t = mu2 - R mu1 with the normalised quaternion.
// m[0] is the number of pairs, or the sum of the weights for the weighted variant (:208-297), whose
// size guard is on the number of pairs and is applied by the caller.
__device__ int solve_absor(const double* m, const double* c, double* out, bool weighted = false) {
  const double n = m[0];
  if (!weighted && n < 3.0) return 0;
  if (weighted && !(n == n && n != 0.0)) return 0;
  double mu1[3], mu2[3], Mm[9], Nm[16], V[16], ev[4], R[9];
  for (int j = 0; j < 3; j++) { mu1[j] = m[1 + j] / n; mu2[j] = m[4 + j] / n; }
  for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) Mm[r * 3 + cc] = m[7 + r * 3 + cc] - n * mu1[r] * mu2[cc];
  const double tr = Mm[0] + Mm[4] + Mm[8];
  const double A12 = Mm[5] - Mm[7], A20 = Mm[6] - Mm[2], A01 = Mm[1] - Mm[3];
  for (int i = 0; i < 16; i++) Nm[i] = 0.0;
  Nm[0] = tr; Nm[1] = A12; Nm[2] = A20; Nm[3] = A01; Nm[4] = A12; Nm[8] = A20; Nm[12] = A01;
  for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) Nm[(r + 1) * 4 + cc + 1] = ((r == cc) ? -tr : 0.0) + (Mm[r * 3 + cc] + Mm[cc * 3 + r]);
  sym_eig<4>(Nm, V, ev);
  double q[4];
  for (int r = 0; r < 4; r++) { q[r] = V[r * 4 + 3]; out[r] = q[r]; }
  const double norm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  quat_to_rot(q[0] / norm, q[1] / norm, q[2] / norm, q[3] / norm, R);
  for (int r = 0; r < 3; r++) {
    const double f = R[3 * r] * (mu1[0] + c[0]) + R[3 * r + 1] * (mu1[1] + c[1]) + R[3 * r + 2] * (mu1[2] + c[2]);
    out[4 + r] = (mu2[r] + c[3 + r]) - f;
  }
  return 7;
}
// RayIntersectionParametersEstimator.cxx:124-143
__device__ int solve_ray(const double* m, const double* c, double* out) {
  const double n = m[0];
  double A[9], x[3];
  A[0] = n - m[1]; A[1] = -m[2]; A[2] = -m[3];
  A[3] = A[1]; A[4] = n - m[4]; A[5] = -m[5];
  A[6] = A[2]; A[7] = A[5]; A[8] = n - m[6];
  if (n < 1.0 || sym_pinv_solve<3>(A, m + 7, x) < 3) return 0;
  for (int j = 0; j < 3; j++) out[j] = x[j] + c[j];
  return 3;
}
// PivotCalibrationParametersEstimator.cxx:63-96 through the 6x6 normal equations
__device__ int solve_pivot(const double* m, const double* c, double* out) {
  const double n = m[0];
  if (n < 3.0) return 0;
  double A[36], b[6], x[6];
  int o = 1;
  for (int a = 0; a < 3; a++) for (int bb = a; bb < 3; bb++) { const double v = m[o++]; A[a * 6 + bb] = v; A[bb * 6 + a] = v; }
  const double* SR = m + 7;  // sum R, row-major
  for (int a = 0; a < 3; a++) for (int bb = 0; bb < 3; bb++) {
    A[a * 6 + 3 + bb] = -SR[bb * 3 + a];  // -(sum R)^T
    A[(3 + bb) * 6 + a] = -SR[bb * 3 + a];
    A[(3 + a) * 6 + 3 + bb] = (a == bb) ? n : 0.0;
  }
  for (int a = 0; a < 3; a++) { b[a] = -m[16 + a]; b[3 + a] = m[19 + a]; }
  if (sym_pinv_solve<6>(A, b, x) < 6) return 0;
  for (int j = 0; j < 3; j++) { out[j] = x[j]; out[3 + j] = x[3 + j] + c[9 + j]; }
  return 6;
}

// DenseLinearEquationSystemParametersEstimator.hxx:64-96 through the n x n normal equations; rank < n -> no solution
template <int N> __device__ int solve_dense(const double* m, double* out) {
  if (m[0] < (double)N) return 0;
  double A[N * N], b[N];
  int o = 1;
  for (int a = 0; a < N; a++) for (int bb = a; bb < N; bb++) { const double v = m[o++]; A[a * N + bb] = v; A[bb * N + a] = v; }
  for (int a = 0; a < N; a++) b[a] = m[o++];
  return sym_pinv_solve<N>(A, b, out) < N ? 0 : N;
}

// Analytic cross-wire calibration (SinglePointTargetUSCalibrationParametersEstimator.cxx:120-270) through the
// 12 x 12 normal equations; t1 comes back relative to the centre of the t2 components.
__device__ int solve_us(const double* m, const double* c, double* out) {
  if (m[0] < 4.0) return 0;
  // diagonally scaled Cholesky of the 12 x 12 normal equations; a pivot below 1e-13 of the unit diagonal means
  // rank < 12 (the reference's "points do not yield a solution", .cxx:195-196)
  double S[144], sc[12], y[12], x[12];
  {
    int o = 1;
    for (int a = 0; a < 12; a++) for (int bb = a; bb < 12; bb++) { const double v = m[o++]; S[a * 12 + bb] = v; S[bb * 12 + a] = v; }
    for (int a = 0; a < 12; a++) y[a] = m[o++];
  }
  for (int i = 0; i < 12; i++) { if (!(S[i * 12 + i] > 0)) return 0; sc[i] = 1.0 / sqrt(S[i * 12 + i]); }
  for (int i = 0; i < 12; i++) { for (int j = 0; j < 12; j++) S[i * 12 + j] *= sc[i] * sc[j]; y[i] *= sc[i]; }
  for (int j = 0; j < 12; j++) {
    double d = S[j * 12 + j];
    for (int k = 0; k < j; k++) d -= S[j * 12 + k] * S[j * 12 + k];
    if (!(d > 1e-13)) return 0;
    S[j * 12 + j] = sqrt(d);
    for (int i = j + 1; i < 12; i++) { double t = S[i * 12 + j]; for (int k = 0; k < j; k++) t -= S[i * 12 + k] * S[j * 12 + k]; S[i * 12 + j] = t / S[j * 12 + j]; }
  }
  for (int i = 0; i < 12; i++) { double t = y[i]; for (int k = 0; k < i; k++) t -= S[i * 12 + k] * x[k]; x[i] = t / S[i * 12 + i]; }
  for (int i = 11; i >= 0; i--) { double t = x[i]; for (int k = i + 1; k < 12; k++) t -= S[k * 12 + i] * x[k]; x[i] = t / S[i * 12 + i]; }
  for (int i = 0; i < 12; i++) x[i] *= sc[i];
  if (c) for (int j = 0; j < 3; j++) x[9 + j] += c[9 + j];
  return us_post(x, out) ? 20 : 0;
}

// Analytic calibrated-pointer calibration (.cxx:789-920) through the 9 x 9 normal equations (scaled Cholesky as above)
__device__ int solve_uscp(const double* m, double* out) {
  if (m[0] < 3.0) return 0;
  double S[81], sc[9], y[9], x[9];
  {
    int o = 1;
    for (int a = 0; a < 9; a++) for (int bb = a; bb < 9; bb++) { const double v = m[o++]; S[a * 9 + bb] = v; S[bb * 9 + a] = v; }
    for (int a = 0; a < 9; a++) y[a] = m[o++];
  }
  for (int i = 0; i < 9; i++) { if (!(S[i * 9 + i] > 0)) return 0; sc[i] = 1.0 / sqrt(S[i * 9 + i]); }
  for (int i = 0; i < 9; i++) { for (int j = 0; j < 9; j++) S[i * 9 + j] *= sc[i] * sc[j]; y[i] *= sc[i]; }
  for (int j = 0; j < 9; j++) {
    double d = S[j * 9 + j];
    for (int k = 0; k < j; k++) d -= S[j * 9 + k] * S[j * 9 + k];
    if (!(d > 1e-13)) return 0;
    S[j * 9 + j] = sqrt(d);
    for (int i = j + 1; i < 9; i++) { double t = S[i * 9 + j]; for (int k = 0; k < j; k++) t -= S[i * 9 + k] * S[j * 9 + k]; S[i * 9 + j] = t / S[j * 9 + j]; }
  }
  for (int i = 0; i < 9; i++) { double t = y[i]; for (int k = 0; k < i; k++) t -= S[i * 9 + k] * x[k]; x[i] = t / S[i * 9 + i]; }
  for (int i = 8; i >= 0; i--) { double t = x[i]; for (int k = i + 1; k < 9; k++) t -= S[k * 9 + i] * x[k]; x[i] = t / S[i * 9 + i]; }
  for (int i = 0; i < 9; i++) x[i] *= sc[i];
  return uscp_post(x, out) ? 17 : 0;
}

// out[0] = number of parameters (0 = the reference's empty vector), out[1..] = parameters.
// For CIRCLE2/SPHERE3 the parameters stay in centred coordinates when keep_centred != 0 (LM start).
__global__ void solve_moments_kernel(int model, DataView dv, const double* __restrict__ m, int keep_centred, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double p[LSQR_MAX_PARAMS];
  int np = 0;
  const double* c = dv.center;
  switch (model) {
    case PLANE3: np = solve_scatter<3>(m, c, 0, p); break;
    case PLANE4: np = solve_scatter<4>(m, c, 0, p); break;
    case LINE3: np = solve_scatter<3>(m, c, 2, p); break;
    case LINE2: np = solve_scatter<2>(m, c, 1, p); break;
    case LINE2D: np = solve_line2d(m, c, p); break;
    case CIRCLE2: np = solve_sphere_alg<2>(m, p); if (np && !keep_centred) { p[0] += c[0]; p[1] += c[1]; } break;
    case SPHERE3: np = solve_sphere_alg<3>(m, p); if (np && !keep_centred) { p[0] += c[0]; p[1] += c[1]; p[2] += c[2]; } break;
    case SPHERE4: np = solve_sphere_alg<4>(m, p); if (np && !keep_centred) { p[0] += c[0]; p[1] += c[1]; p[2] += c[2]; p[3] += c[3]; } break;
    case ABSOR: np = solve_absor(m, c, p); break;
    case RAY: np = solve_ray(m, c, p); break;
    case PIVOT: np = solve_pivot(m, c, p); break;
    case DENSE5: np = solve_dense<5>(m, p); break;
    case DENSE6: np = solve_dense<6>(m, p); break;
    case USXW: np = solve_us(m, keep_centred ? nullptr : c, p); break;   // LM start: t1 stays relative to the centre
    case USCP: np = solve_uscp(m, p); break;
  }
  out[0] = (double)np;
  for (int j = 0; j < np; j++) out[1 + j] = p[j];
}
__global__ void solve_weighted_absor_kernel(DataView dv, const double* __restrict__ m, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double p[8];
  const int np = solve_absor(m, dv.center, p, true);
  out[0] = (double)np;
  for (int j = 0; j < np; j++) out[1 + j] = p[j];
}
void launch_solve_weighted_absor(const DataView& dv, const double* moments, double* out_dev, cudaStream_t s) {
  solve_weighted_absor_kernel<<<1, 32, 0, s>>>(dv, moments, out_dev);
}
void launch_solve_moments(int model, const DataView& dv, const double* moments, int keep_centred, double* out_dev, cudaStream_t s) {
  solve_moments_kernel<<<1, 32, 0, s>>>(model, dv, moments, keep_centred, out_dev);
}

// ---------------------------------------------------------------------------------------
// Levenberg-Marquardt controller: lm_minpack.cuh (MINPACK's lmder restated on the normal equations, the reference's
// tolerances per estimator).  Result only if MINPACK would report info 1..4 (vnl_levenberg_marquardt::minimize returns
// true), else empty parameters.  One pass of mask_moments_kernel delivers J^T J, J^T f and |f|^2 at the point the state
// asks for.
// ---------------------------------------------------------------------------------------
__host__ __device__ inline void lm_update_model(int model, const double* m, double* st) {
  if (model == CIRCLE2) lm_update<3>(m, st, lm_tolerances(0));
  else if (model == SPHERE3) lm_update<4>(m, st, lm_tolerances(0));
  else if (model == SPHERE4) lm_update<5>(m, st, lm_tolerances(0));
  else if (model == USXW) lm_update<11>(m, st, lm_tolerances(1));
  else if (model == USCP) lm_update<8>(m, st, lm_tolerances(2));
}

__global__ void lm_init_kernel(const double* __restrict__ alg_out, double* __restrict__ st) {
  if (threadIdx.x != 0) return;
  for (int i = 0; i < LM_SIZE; i++) st[i] = 0.0;
  const int np = (int)alg_out[0];
  if (np == 0) { st[LM_STATUS] = 2.0; return; }
  for (int j = 0; j < np && j < kLmMaxP; j++) st[LM_X + j] = alg_out[1 + j];   // cross-wire: the first 11 of its 20 parameters
}
__global__ void lm_update_kernel(int model, const double* __restrict__ m, double* __restrict__ st) {
  if (threadIdx.x != 0) return;
  lm_update_model(model, m, st);
}
// .cxx:305-327: entries 11..19 rebuilt from the optimised angles and scales
__device__ void us_expand(const double* x, double* prm) {
  for (int j = 0; j < 11; j++) prm[j] = x[j];
  const double cz = cos(x[6]), sz = sin(x[6]), cy = cos(x[7]), sy = sin(x[7]), cx = cos(x[8]), sx = sin(x[8]);
  prm[11] = x[9] * cz * cy; prm[12] = x[9] * sz * cy; prm[13] = -x[9] * sy;
  prm[14] = x[10] * (cz * sy * sx - sz * cx); prm[15] = x[10] * (sz * sy * sx + cz * cx); prm[16] = x[10] * cy * sx;
  prm[17] = cz * sy * cx + sz * sx; prm[18] = sz * sy * cx - cz * sx; prm[19] = cy * cx;
}
__global__ void lm_finish_kernel(int model, DataView dv, const double* __restrict__ st, double* __restrict__ out) {
  if (threadIdx.x != 0) return;
  if (st[LM_STATUS] != 1.0) { out[0] = 0.0; return; }
  if (model == USXW) {
    double x[11];
    for (int j = 0; j < 11; j++) x[j] = st[LM_X + j];
    for (int j = 0; j < 3; j++) x[j] += dv.center[9 + j];   // t1 was estimated relative to the centre of the t2 components
    out[0] = 20;
    us_expand(x, out + 1);
    return;
  }
  if (model == USCP) {   // .cxx:958-983: entries 8..16 rebuilt from the optimised angles and scales
    double x[11], full[20];
    x[0] = x[1] = x[2] = 0.0;
    for (int j = 0; j < 8; j++) x[3 + j] = st[LM_X + j];
    us_expand(x, full);
    out[0] = 17;
    for (int j = 0; j < 17; j++) out[1 + j] = full[3 + j];
    return;
  }
  const int dim = (model == CIRCLE2) ? 2 : (model == SPHERE3 ? 3 : 4);
  out[0] = dim + 1;
  for (int j = 0; j < dim; j++) out[1 + j] = st[LM_X + j] + dv.center[j];
  out[1 + dim] = st[LM_X + dim];
}
void launch_lm_init(const double* alg_out_dev, double* state, cudaStream_t s) { lm_init_kernel<<<1, 32, 0, s>>>(alg_out_dev, state); }
void launch_lm_update(int model, const double* moments, double* state, cudaStream_t s) { lm_update_kernel<<<1, 32, 0, s>>>(model, moments, state); }
void launch_lm_finish(int model, const DataView& dv, const double* state, double* out_dev, cudaStream_t s) { lm_finish_kernel<<<1, 32, 0, s>>>(model, dv, state, out_dev); }
int lm_status_offset() { return LM_STATUS; }

// ---------------------------------------------------------------------------------------
__global__ void expand_mask_kernel(const uint32_t* __restrict__ bits, uint32_t first, uint32_t count, uint8_t* __restrict__ bytes) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) { const uint32_t i = first + t; bytes[i] = (bits[i >> 5] >> (i & 31)) & 1u; }
}
void launch_expand_mask(const uint32_t* bits, uint32_t first, uint32_t count, uint8_t* bytes, cudaStream_t s) {
  if (count) expand_mask_kernel<<<(count + 255) / 256, 256, 0, s>>>(bits, first, count, bytes);
}


// ---------------------------------------------------------------------------------------
// Batched small problems: one thread block per problem (BASELINE.json config 5; in the
// reference this is a host loop of RANSAC<T,S>::compute calls).  Everything -- subset
// generation, minimal solve, consensus, arg-max, consensus set, least-squares refine -- happens
// inside the block with the problem's points resident in shared memory, in fp64 reference
// arithmetic.  Exhaustive mode enumerates all C(n,k) subsets (RANSAC.hxx:150-249); otherwise
// rounds of blockDim.x Philox hypotheses with the stop rule of RANSAC.hxx:107-110 between rounds.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long r = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) r = sh[w] > r ? sh[w] : r;
  return r;
}

template <int M>
__device__ void block_moments(const double* pts, uint32_t n, uint32_t ldp, const double* hq, const EstCfg& cfg, const double* lmx, bool lm,
                              int use_mask, uint8_t* mask_out, double* sh_part, double* sh_mom) {
  constexpr int D = Model<M>::D;
  constexpr int NMA = Mom<M>::N, NML = Mom<M>::NLM;
  double acc[kMaxMoments];
  const int nm = lm ? NML : NMA;
  for (int j = 0; j < kMaxMoments; j++) acc[j] = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    double x[D];
#pragma unroll
    for (int d = 0; d < D; d++) x[d] = pts[d * ldp + i];
    const bool in = use_mask ? agree<M>(hq, x, cfg) : true;
    if (mask_out) mask_out[i] = in ? 1 : 0;
    if (in) {
      if (lm) {
        if constexpr (M == CIRCLE2) acc_sphere_lm<2>(x, lmx, acc);
        if constexpr (M == SPHERE3) acc_sphere_lm<3>(x, lmx, acc);
        if constexpr (M == SPHERE4) acc_sphere_lm<4>(x, lmx, acc);
      } else { acc[0] += 1.0; accumulate<M>(x, acc); }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  for (int j = 0; j < nm; j++) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh_part[warp * kMaxMoments + j] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x < nm) { double v = 0; for (int w = 0; w < nw; w++) v += sh_part[w * kMaxMoments + threadIdx.x]; sh_mom[threadIdx.x] = v; }
  __syncthreads();
}

template <int M>
__global__ void __launch_bounds__(256) batch_kernel(BatchArgs a, EstCfg cfg, int ls_type, uint32_t group) {
  constexpr int D = Model<M>::D, P = Model<M>::P, K = Model<M>::K, HQ = Model<M>::HQ;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* pts = reinterpret_cast<double*>(smem_raw);  // [D][ldp]
  const uint32_t ldp = a.max_n;
  __shared__ unsigned long long sh_key[8];
  __shared__ double sh_part[8 * kMaxMoments];
  __shared__ double sh_mom[kMaxMoments];
  __shared__ double sh_prm[LSQR_MAX_PARAMS + 4];
  __shared__ double sh_state[LM_SIZE];
  __shared__ unsigned long long sh_best;
  __shared__ unsigned long long sh_tries;
  __shared__ int sh_ok;

  const uint32_t b = blockIdx.x;
  const uint64_t off = a.offsets[b] - a.base;                 // record offset inside this launch's data
  const uint32_t n = (uint32_t)(a.offsets[b + 1] - a.offsets[b]);
  const uint64_t gb = a.first_problem + b;                    // global problem index: the sampler's counter does not depend on how problems are split over GPUs
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (uint32_t i = threadIdx.x; i < n * D; i += blockDim.x) pts[(i % D) * ldp + (i / D)] = a.data[off * D + i];
  if (threadIdx.x == 0) {
    sh_best = 0ull;
    unsigned long long all = 0xFFFFFFFFull;  // RANSAC::choose saturates at UINT_MAX (RANSAC.hxx:254-280)
    if (n >= (uint32_t)K) { const uint64_t c = binom(n, K); all = c < all ? c : all; } else all = 0;
    sh_tries = a.exhaustive ? all : (all < a.tries ? all : (unsigned long long)a.tries);
  }
  __syncthreads();

  unsigned long long best = 0ull;
  const double log1mp = log(1.0 - a.prob);
  // G threads share one hypothesis (each counts every G-th datum).  Randomized mode re-evaluates the stop rule after every
  // round and ends in a single-thread refit, so it runs small blocks (many resident per SM) with short rounds; exhaustive
  // mode is throughput-bound and scores one hypothesis per thread (launch_batch picks block size and G).
  const uint32_t G = group;
  const uint32_t per_round = blockDim.x / G, sub_lane = threadIdx.x % G;
  for (unsigned long long done = 0; done < sh_tries; done += per_round) {
    const unsigned long long h = done + threadIdx.x / G;
    unsigned long long key = 0ull;
    if (h < sh_tries) {
      int32_t sub[K];
      if (a.exhaustive) unrank_lex<K>(h, n, sub); else sample_subset<K>((gb << 32) | h, a.seed, n, sub);
      double sp[K * D], prm[P], hq[HQ];
#pragma unroll
      for (int j = 0; j < K; j++)
#pragma unroll
        for (int d = 0; d < D; d++) sp[j * D + d] = pts[d * ldp + sub[j]];
      const bool ok = estimate<M>(sp, cfg, prm);   // the same for all G threads of a hypothesis
      uint32_t c = 0;
      if (ok) {
        prepare<M>(prm, hq);
        for (uint32_t i = sub_lane; i < n; i += G) {
          double x[D];
#pragma unroll
          for (int d = 0; d < D; d++) x[d] = pts[d * ldp + i];
          c += agree<M>(hq, x, cfg) ? 1u : 0u;
        }
      }
      if (G > 1) {   // uniform over the block; the G threads of a hypothesis are adjacent lanes of one warp
        const unsigned grp = (0xFFFFFFFFu >> (32u - G)) << ((threadIdx.x & 31u) & ~(G - 1u));
        for (uint32_t o = 1; o < G; o <<= 1) c += __shfl_xor_sync(grp, c, o);
      }
      if (ok) key = ((unsigned long long)c << 32) | (0xFFFFFFFFull - h);
    }
    const unsigned long long round_best = block_max_u64(key, sh_key);
    if (round_best > best) {
      best = round_best;
      if (!a.exhaustive && threadIdx.x == 0) {  // stop rule, RANSAC.hxx:104-110
        const uint32_t c = (uint32_t)(best >> 32);
        unsigned long long cap = sh_tries;
        if (c == n) cap = 0;
        else if (a.prob > 0.0 && a.prob < 1.0) {
          const double den = log(1.0 - pow((double)c / (double)n, (double)K));
          const double t = log1mp / den + 0.5;
          const unsigned long long nt = t >= 4294967295.0 ? 0xFFFFFFFFull : (unsigned long long)(long long)t;
          cap = nt < cap ? nt : cap;
        }
        sh_tries = cap;
      }
    }
    __syncthreads();
  }

  // winner -> consensus set -> least squares (RANSAC.hxx:129-138)
  const uint32_t best_count = (uint32_t)(best >> 32);
  if (threadIdx.x == 0) {
    sh_ok = 0;
    if (best_count > 0) {
      const unsigned long long h = 0xFFFFFFFFull - (best & 0xFFFFFFFFull);
      int32_t sub[K];
      if (a.exhaustive) unrank_lex<K>(h, n, sub); else sample_subset<K>((gb << 32) | h, a.seed, n, sub);
      double sp[K * D], prm[P];
      for (int j = 0; j < K; j++) for (int d = 0; d < D; d++) sp[j * D + d] = pts[d * ldp + sub[j]];
      if (estimate<M>(sp, cfg, prm)) { prepare<M>(prm, sh_prm); sh_ok = 1; }
    }
  }
  __syncthreads();
  uint8_t* mask_out = a.out_masks ? a.out_masks + off : nullptr;
  if (threadIdx.x == 0) a.out_counts[b] = best_count;
  if (!sh_ok) {
    for (uint32_t i = threadIdx.x; i < (uint32_t)P; i += blockDim.x) a.out_params[(size_t)b * P + i] = nan;
    if (mask_out) for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) mask_out[i] = 0;
    return;
  }
  double hq[HQ];
#pragma unroll
  for (int j = 0; j < HQ; j++) hq[j] = sh_prm[j];
  block_moments<M>(pts, n, ldp, hq, cfg, nullptr, false, 1, mask_out, sh_part, sh_mom);
  double zero_center[kMaxDim];
  for (int j = 0; j < kMaxDim; j++) zero_center[j] = 0.0;
  __shared__ double sh_out[LSQR_MAX_PARAMS + 4];
  if (threadIdx.x == 0) {
    double p[LSQR_MAX_PARAMS];
    int np = 0;
    const double* m = sh_mom;
    const double* c = zero_center;
    switch (M) {
      case PLANE3: np = solve_scatter<3>(m, c, 0, p); break;
      case PLANE4: np = solve_scatter<4>(m, c, 0, p); break;
      case LINE3: np = solve_scatter<3>(m, c, 2, p); break;
      case LINE2: np = solve_scatter<2>(m, c, 1, p); break;
      case LINE2D: np = solve_line2d(m, c, p); break;
      case CIRCLE2: np = solve_sphere_alg<2>(m, p); break;
      case SPHERE3: np = solve_sphere_alg<3>(m, p); break;
      case SPHERE4: np = solve_sphere_alg<4>(m, p); break;
      case ABSOR: np = solve_absor(m, c, p); break;
      case RAY: np = solve_ray(m, c, p); break;
      case PIVOT: np = solve_pivot(m, c, p); break;
      case DENSE5: np = solve_dense<5>(m, p); break;
      case DENSE6: np = solve_dense<6>(m, p); break;
    }
    sh_out[0] = np;
    for (int j = 0; j < np; j++) sh_out[1 + j] = p[j];
  }
  __syncthreads();
  if constexpr (M == CIRCLE2 || M == SPHERE3 || M == SPHERE4) {
    if (ls_type == 1) {  // geometric: Levenberg-Marquardt from the algebraic fit
      if (threadIdx.x == 0) {
        for (int i = 0; i < LM_SIZE; i++) sh_state[i] = 0.0;
        const int np = (int)sh_out[0];
        if (np == 0) sh_state[LM_STATUS] = 2.0;
        for (int j = 0; j < np; j++) sh_state[LM_X + j] = sh_out[1 + j];
      }
      __syncthreads();
      // The controller state is touched by thread 0 only; what the block needs per pass (status, evaluation
      // point) goes through sh_bcast, with a barrier on either side of every read.
      __shared__ double sh_bcast[6];
      for (;;) {
        if (threadIdx.x == 0) {
          sh_bcast[0] = sh_state[LM_STATUS];
          const int o = (sh_state[LM_PHASE] != 0.0) ? LM_TRIAL : LM_X;
          for (int j = 0; j < 5; j++) sh_bcast[1 + j] = sh_state[o + j];
        }
        __syncthreads();
        double lmx[5];
        const double status = sh_bcast[0];
        for (int j = 0; j < 5; j++) lmx[j] = sh_bcast[1 + j];
        __syncthreads();
        if (status != 0.0) break;
        block_moments<M>(pts, n, ldp, hq, cfg, lmx, true, 1, nullptr, sh_part, sh_mom);
        if (threadIdx.x == 0) lm_update_model(M, sh_mom, sh_state);
      }
      if (threadIdx.x == 0) {
        if (sh_state[LM_STATUS] != 1.0) sh_out[0] = 0.0;
        else for (int j = 0; j < P; j++) sh_out[1 + j] = sh_state[LM_X + j];
      }
      __syncthreads();
    }
  }
  const int np = (int)sh_out[0];
  for (uint32_t i = threadIdx.x; i < (uint32_t)P; i += blockDim.x) a.out_params[(size_t)b * P + i] = np ? sh_out[1 + i] : nan;
}

int launch_batch(const BatchArgs& a, const EstCfg& cfg, int ls_type, cudaStream_t s) {
  if (a.n_problems == 0) return 0;
  const int D = model_info(a.model).D;
  const size_t smem = (size_t)D * a.max_n * sizeof(double);
  if (smem > 200 * 1024) return -1;
  const int threads = a.exhaustive ? 256 : 64;
  const uint32_t group = a.exhaustive ? 1u : 2u;
#define CALL(MM)                                                                                  \
  {                                                                                               \
    auto kern = batch_kernel<MM>;                                                                 \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    kern<<<a.n_problems, threads, smem, s>>>(a, cfg, ls_type, group);                            \
  }
  switch (a.model) {
    case PLANE3: CALL(PLANE3) break;
    case LINE2D: CALL(LINE2D) break;
    case LINE2: CALL(LINE2) break;
    case LINE3: CALL(LINE3) break;
    case CIRCLE2: CALL(CIRCLE2) break;
    case SPHERE3: CALL(SPHERE3) break;
    case SPHERE4: CALL(SPHERE4) break;
    case PLANE4: CALL(PLANE4) break;
    case ABSOR: CALL(ABSOR) break;
    case RAY: CALL(RAY) break;
    case PIVOT: CALL(PIVOT) break;
    case DENSE5: CALL(DENSE5) break;
    case DENSE6: CALL(DENSE6) break;
    default: return -1;
  }
#undef CALL
  return 1;
}

}  // namespace lsqr
