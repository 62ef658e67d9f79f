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

// ---- packed pair arithmetic (sm_100 f32x2) -------------------------------------------------
struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 splat(float a) { f2 r; asm("mov.b64 %0, {%1, %1};" : "=l"(r.v) : "f"(a)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ void halves(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 join(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }

// ---- hoisted fp32 hypotheses ---------------------------------------------------------------
// Q32 floats per hypothesis, computed once in fp64 from the raw parameters, the data centre c
// (positions are stored as x - c in fp32) and the thresholds:
//   PLANE3   (nx,ny,nz, delta - n.(a-c))      the kernels form s' = s + delta and test 0 <= s' < 2 delta (see count_carry)
//   LINE2D   (nx,ny, -n.(a-c))                 |s| < delta (two FFMA2 per pair leave the ALU pipe as the limit either way; measured: the carry form is 4 % slower here)
//   LINE2    (n_perp, -n_perp.(a-c))            the 2-D line form on the perpendicular of the unit direction
//   LINE3    3 x (n_k, -n_j, -((a-c) x n)_i)    Pluecker form
//   CIRCLE/SPHERE (ctr-c, -(r-delta)^2, 2^32 - bits(4 r delta)): t' = d^2 - (r-delta)^2, inlier <=> 0 <= t' < 4 r delta  (<=> |d - r| < delta)
//   ABSOR    (R[9], R c1 + t - c2)
//   RAY      (x - c)
//   PIVOT    (tDRF, -(tW - c))
//   DENSE n  (x[n], -1, delta): a.x - b + delta as n+1 FMAs over the (uncentred) augmented row
//   USXW     (m_x R3(:,1), m_y R3(:,2), t3, -(t1 - c))
//   USCP     (m_x R3(:,1), m_y R3(:,2), t3)
template <int M> __device__ __forceinline__ void hoist32(const double* p, const double* c, const EstCfg& cfg, float* q);
template <> __device__ __forceinline__ void hoist32<PLANE3>(const double* p, const double* c, const EstCfg& cfg, float* q) {
  q[0] = (float)p[0]; q[1] = (float)p[1]; q[2] = (float)p[2];
  q[3] = (float)(cfg.delta - (p[0] * (p[3] - c[0]) + p[1] * (p[4] - c[1]) + p[2] * (p[5] - c[2])));   // shifted residual s + delta
}
template <int DIM> __device__ __forceinline__ void hoist_plane_nd(const double* p, const double* c, const EstCfg& cfg, float* q) {
  double s = 0;
  for (int i = 0; i < DIM; i++) { q[i] = (float)p[i]; s += p[i] * (p[DIM + i] - c[i]); }
  q[DIM] = (float)(cfg.delta - s);
}
#define LSQR_DEF_(ID, DIM) template <> __device__ __forceinline__ void hoist32<ID>(const double* p, const double* c, const EstCfg& cfg, float* q) { hoist_plane_nd<DIM>(p, c, cfg, q); }
LSQR_PLANE_ND_LIST(LSQR_DEF_)
#undef LSQR_DEF_
template <> __device__ __forceinline__ void hoist32<LINE2D>(const double* p, const double* c, const EstCfg&, float* q) {
  q[0] = (float)p[0]; q[1] = (float)p[1];
  q[2] = (float)(-(p[0] * (p[2] - c[0]) + p[1] * (p[3] - c[1])));
}
// kD lines (LineParametersEstimator.hxx:135-150: |v - (v.n) n|^2 < delta^2 with v = x - a).  The fast mode takes the direction as
// the unit vector it is for every hypothesis estimate() produces (it is re-normalised here, a no-op then), which makes the
// rejection |v x n| in 3-D and |n_perp . v| in 2-D:
//   d = 2: the 2-D line form, s = n_perp . x - n_perp . (a - c)                                        2 FMA per datum
//   d = 3: Pluecker form, w = x X n - (a - c) X n, every component two FMAs on a hoisted constant:      6 + 3 FMA per datum
// instead of the 9 / 12 operations of the literal form.  The validation mode keeps the literal form (models.cuh).
template <> __device__ __forceinline__ void hoist32<LINE2>(const double* p, const double* c, const EstCfg&, float* q) {
  const double inv = 1.0 / sqrt(p[0] * p[0] + p[1] * p[1]);
  const double px = -p[1] * inv, py = p[0] * inv;
  q[0] = (float)px; q[1] = (float)py;
  q[2] = (float)(-(px * (p[2] - c[0]) + py * (p[3] - c[1])));
}
template <> __device__ __forceinline__ void hoist32<LINE3>(const double* p, const double* c, const EstCfg&, float* q) {
  const double inv = 1.0 / sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  const double n[3] = {p[0] * inv, p[1] * inv, p[2] * inv}, a[3] = {p[3] - c[0], p[4] - c[1], p[5] - c[2]};
  // w_i = x_j n_k - x_k n_j - (a_j n_k - a_k n_j)   with (i, j, k) cyclic: constants (n_k, -n_j, -m_i) per component
  for (int i = 0; i < 3; i++) {
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    q[3 * i] = (float)n[k]; q[3 * i + 1] = (float)(-n[j]); q[3 * i + 2] = (float)(-(a[j] * n[k] - a[k] * n[j]));
  }
}
// d >= 4: the literal form, constants (n, a - c)
template <int DIM> __device__ __forceinline__ void hoist_line_nd(const double* p, const double* c, float* q) {
  for (int i = 0; i < DIM; i++) { q[i] = (float)p[i]; q[DIM + i] = (float)(p[DIM + i] - c[i]); }
}
#define LSQR_DEF_(ID, DIM) template <> __device__ __forceinline__ void hoist32<ID>(const double* p, const double* c, const EstCfg&, float* q) { hoist_line_nd<DIM>(p, c, q); }
LSQR_DEF_(LINE4, 4) LSQR_DEF_(LINE5, 5) LSQR_DEF_(LINE6, 6) LSQR_DEF_(LINE7, 7) LSQR_DEF_(LINE8, 8)
#undef LSQR_DEF_
// |d - r| < delta  <=>  (r - delta)^2 <= d^2 < (r + delta)^2  <=>  0 <= t' < 4 r delta  with  t' = d^2 - (r - delta)^2: the chain
// starts at -(r - delta)^2 and the datum is an inlier iff bits(t') < bits(4 r delta) as unsigned integers (count_carry; the
// threshold is per hypothesis).  r < delta: no lower bound, t' = d^2 + 1 against (r + delta)^2 + 1.  A threshold that is not
// positive (negative delta: SphereParametersEstimator.hxx:20 keeps its sign) is stored as 0: nothing agrees.
template <int DIM> __device__ __forceinline__ void hoist_sphere(const double* p, const double* c, const EstCfg& cfg, float* q) {
  for (int i = 0; i < DIM; i++) q[i] = (float)(p[i] - c[i]);
  const double r = p[DIM], dl = cfg.delta;
  double start, width;
  if (r >= dl) { start = -(r - dl) * (r - dl); width = 4.0 * r * dl; }
  else { start = 1.0; width = (r + dl) * (r + dl) + 1.0; }
  q[DIM] = (float)start;
  const float wf = (float)width;
  // stored ready for count_carry: 2^32 - bits(window); 0 = nothing agrees (window not positive, or NaN)
  q[DIM + 1] = __uint_as_float((wf > 0.0f) ? 0u - __float_as_uint(wf) : 0u);
}
#define LSQR_DEF_(ID, DIM) template <> __device__ __forceinline__ void hoist32<ID>(const double* p, const double* c, const EstCfg& cfg, float* q) { hoist_sphere<DIM>(p, c, cfg, q); }
LSQR_SPHERE_ALL_LIST(LSQR_DEF_)
#undef LSQR_DEF_
template <> __device__ __forceinline__ void hoist32<ABSOR>(const double* p, const double* c, const EstCfg&, float* q) {
  double R[9];
  quat_to_rot(p[0], p[1], p[2], p[3], R);
  for (int i = 0; i < 9; i++) q[i] = (float)R[i];
  for (int i = 0; i < 3; i++) q[9 + i] = (float)(R[3 * i] * c[0] + R[3 * i + 1] * c[1] + R[3 * i + 2] * c[2] + p[4 + i] - c[3 + i]);
}
template <> __device__ __forceinline__ void hoist32<RAY>(const double* p, const double* c, const EstCfg&, float* q) {
  for (int i = 0; i < 3; i++) q[i] = (float)(p[i] - c[i]);
}
template <> __device__ __forceinline__ void hoist32<PIVOT>(const double* p, const double* c, const EstCfg&, float* q) {
  for (int i = 0; i < 3; i++) { q[i] = (float)p[i]; q[3 + i] = (float)(-(p[3 + i] - c[9 + i])); }   // the chain starts at -(tW - c)
}

template <int N> __device__ __forceinline__ void hoist_dense(const double* p, const EstCfg& cfg, float* q) {
  for (int i = 0; i < N; i++) q[i] = (float)p[i];
  q[N] = -1.0f;
  q[N + 1] = (float)cfg.delta;   // the chain starts at +delta: s' = a.x - b + delta, inlier <=> bits(s') < bits(2 delta)
}
#define LSQR_DEF_(ID, N) template <> __device__ __forceinline__ void hoist32<ID>(const double* p, const double*, const EstCfg& cfg, float* q) { hoist_dense<N>(p, cfg, q); }
LSQR_DENSE_N_LIST(LSQR_DEF_)
#undef LSQR_DEF_

template <> __device__ __forceinline__ void hoist32<USXW>(const double* p, const double* c, const EstCfg&, float* q) {
  for (int i = 0; i < 6; i++) q[i] = (float)p[11 + i];
  for (int i = 0; i < 3; i++) { q[6 + i] = (float)p[3 + i]; q[9 + i] = (float)(-(p[i] - c[9 + i])); }   // the chain starts at -(t1 - c)
}

template <> __device__ __forceinline__ void hoist32<USCP>(const double* p, const double*, const EstCfg&, float* q) {
  for (int i = 0; i < 6; i++) q[i] = (float)p[8 + i];
  for (int i = 0; i < 3; i++) q[6 + i] = (float)p[i];
}

// slot of the per-hypothesis counting window among the hoisted constants (-1: the window is 2 delta for every hypothesis)
template <int M> constexpr int thr_slot() { return model_family(M) == FAM_SPHERE ? model_dim(M) + 1 : -1; }

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

// ---- per-model residual forms on point pairs ----------------------------------------------
struct Thr2 { f2 delta, neg_delta2; float fdelta; };

// signed(): g with inlier <=> g < 0.   For PLANE3 / LINE2D, dist(): s with inlier <=> |s| < delta.
template <int M> struct Eval;
template <> struct Eval<PLANE3> {
  static constexpr bool kHasAbsForm = true;
  static constexpr int kThr = -1;
  static constexpr bool kShifted = true;   // dist() is s + delta (the hoisted constant carries the shift)
  __device__ static __forceinline__ f2 dist(const f2* q, const f2* x) { return fma2(q[0], x[0], fma2(q[1], x[1], fma2(q[2], x[2], q[3]))); }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2& t) { const f2 s = sub2(dist(q, x), t.delta); return fma2(s, s, t.neg_delta2); }
};
template <int DIM> struct EvalPlaneND {
  static constexpr bool kHasAbsForm = true;
  static constexpr int kThr = -1;
  static constexpr bool kShifted = true;   // dist() is s + delta (the hoisted constant carries the shift)
  __device__ static __forceinline__ f2 dist(const f2* q, const f2* x) {
    f2 s = q[DIM];
#pragma unroll
    for (int i = DIM - 1; i >= 0; i--) s = fma2(q[i], x[i], s);
    return s;
  }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2& t) { const f2 s = sub2(dist(q, x), t.delta); return fma2(s, s, t.neg_delta2); }
};
#define LSQR_DEF_(ID, DIM) template <> struct Eval<ID> : EvalPlaneND<DIM> {};
LSQR_PLANE_ND_LIST(LSQR_DEF_)
#undef LSQR_DEF_
template <> struct Eval<LINE2D> {
  static constexpr bool kHasAbsForm = true;
  static constexpr int kThr = -1;
  static constexpr bool kShifted = false;
  __device__ static __forceinline__ f2 dist(const f2* q, const f2* x) { return fma2(q[0], x[0], fma2(q[1], x[1], q[2])); }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2& t) { const f2 s = dist(q, x); return fma2(s, s, t.neg_delta2); }
};
// Every instruction below takes at most ONE datum operand: in the constant-bank kernel the data are uniform registers and an
// FFMA2 / FADD2 encodes a single uniform source -- a second one forces ptxas to fetch the datum with LDC into ordinary
// registers (indexed constant loads, a quarter of the throughput: measured on the pivot estimator, 0.7 -> 2.3 T evals/s).
template <> struct Eval<LINE2> {     // the 2-D line form on the perpendicular (see hoist32<LINE2>)
  static constexpr bool kHasAbsForm = true;
  static constexpr int kThr = -1;
  static constexpr bool kShifted = false;
  __device__ static __forceinline__ f2 dist(const f2* q, const f2* x) { return fma2(q[0], x[0], fma2(q[1], x[1], q[2])); }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2& t) { const f2 s = dist(q, x); return fma2(s, s, t.neg_delta2); }
};
template <> struct Eval<LINE3> {     // Pluecker form: |x X n - m|^2 - delta^2 (see hoist32<LINE3>)
  static constexpr bool kHasAbsForm = false;
  static constexpr int kThr = -1;
  static constexpr bool kShifted = false;
  __device__ static __forceinline__ f2 dist(const f2*, const f2*) { return splat(0.f); }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2& t) {
    f2 g = t.neg_delta2;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const f2 w = fma2(x[(i + 1) % 3], q[3 * i], fma2(x[(i + 2) % 3], q[3 * i + 1], q[3 * i + 2]));
      g = fma2(w, w, g);
    }
    return g;
  }
};
// d >= 4: the literal form of LineParametersEstimator.hxx:135-150, v = x - a, w = v - (v.n) n, |w|^2 - delta^2: 4d operations
template <int DIM> struct EvalLineND {
  static constexpr bool kHasAbsForm = false;
  static constexpr int kThr = -1;
  static constexpr bool kShifted = false;
  __device__ static __forceinline__ f2 dist(const f2*, const f2*) { return splat(0.f); }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2& t) {
    f2 v[DIM];
#pragma unroll
    for (int i = 0; i < DIM; i++) v[i] = sub2(x[i], q[DIM + i]);
    f2 vn = mul2(v[0], q[0]);
#pragma unroll
    for (int i = 1; i < DIM; i++) vn = fma2(v[i], q[i], vn);
    const f2 nvn = sub2(splat(0.f), vn);
    f2 g = t.neg_delta2;
#pragma unroll
    for (int i = 0; i < DIM; i++) { const f2 w = fma2(nvn, q[i], v[i]); g = fma2(w, w, g); }
    return g;
  }
};
template <> struct Eval<LINE4> : EvalLineND<4> {};
template <> struct Eval<LINE5> : EvalLineND<5> {};
template <> struct Eval<LINE6> : EvalLineND<6> {};
template <> struct Eval<LINE7> : EvalLineND<7> {};
template <> struct Eval<LINE8> : EvalLineND<8> {};
// t' = d^2 - (r - delta)^2 (chain start q[DIM]); per-hypothesis window q[DIM + 1] = 4 r delta (see hoist_sphere)
template <int DIM> __device__ __forceinline__ f2 sphere_t(const f2* q, const f2* x) {
  f2 t = q[DIM];
#pragma unroll
  for (int i = 0; i < DIM; i++) { const f2 w = sub2(x[i], q[i]); t = fma2(w, w, t); }
  return t;
}
template <int DIM> struct EvalSphere {
  static constexpr bool kHasAbsForm = true;
  static constexpr int kThr = DIM + 1;      // per-hypothesis window, stored as 2^32 - bits(window)
  static constexpr bool kShifted = true;    // dist() is already the shifted residual
  __device__ static __forceinline__ f2 dist(const f2* q, const f2* x) { return sphere_t<DIM>(q, x); }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2&) { return sphere_t<DIM>(q, x); }
};
#define LSQR_DEF_(ID, DIM) template <> struct Eval<ID> : EvalSphere<DIM> {};
LSQR_SPHERE_ALL_LIST(LSQR_DEF_)
#undef LSQR_DEF_
template <> struct Eval<ABSOR> {
  static constexpr bool kHasAbsForm = false;
  static constexpr int kThr = -1;
  static constexpr bool kShifted = false;
  __device__ static __forceinline__ f2 dist(const f2*, const f2*) { return splat(0.f); }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2& t) {
    f2 g = t.neg_delta2;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const f2 d = sub2(fma2(q[3 * i], x[0], fma2(q[3 * i + 1], x[1], fma2(q[3 * i + 2], x[2], q[9 + i]))), x[3 + i]);
      g = fma2(d, d, g);
    }
    return g;
  }
};
template <> struct Eval<RAY> {
  static constexpr bool kHasAbsForm = false;
  static constexpr int kThr = -1;
  static constexpr bool kShifted = false;
  __device__ static __forceinline__ f2 dist(const f2*, const f2*) { return splat(0.f); }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2& t) {
    const f2 vx = sub2(q[0], x[0]), vy = sub2(q[1], x[1]), vz = sub2(q[2], x[2]);
    const f2 tt = fma2(x[3], vx, fma2(x[4], vy, mul2(x[5], vz)));
    const f2 ntt = sub2(splat(0.f), tt);
    const f2 dx = fma2(ntt, x[3], vx), dy = fma2(ntt, x[4], vy), dz = fma2(ntt, x[5], vz);
    const f2 g = fma2(dx, dx, fma2(dy, dy, fma2(dz, dz, t.neg_delta2)));
    // t >= 0 && dist^2 < delta^2  <=>  max(g, -t) < 0; with NaN padding the compare fails and the (positive) NaN -t is kept
    float g0, g1, n0, n1;
    halves(g, g0, g1); halves(ntt, n0, n1);
    return join(g0 > n0 ? g0 : n0, g1 > n1 ? g1 : n1);
  }
};
template <> struct Eval<PIVOT> {
  static constexpr bool kHasAbsForm = false;
  static constexpr int kThr = -1;
  static constexpr bool kShifted = false;
  __device__ static __forceinline__ f2 dist(const f2*, const f2*) { return splat(0.f); }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2& t) {
    f2 g = t.neg_delta2;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const f2 d = add2(fma2(q[0], x[3 * i], fma2(q[1], x[3 * i + 1], fma2(q[2], x[3 * i + 2], q[3 + i]))), x[9 + i]);
      g = fma2(d, d, g);
    }
    return g;
  }
};

// |a.x - b| < delta in the shifted form: n+1 FMAs, q[N] = -1 multiplies the right-hand side, the chain starts at q[N+1] = delta
// (DenseLinearEquationSystemParametersEstimator.hxx:111-119)
template <int N> __device__ __forceinline__ f2 dense_dist(const f2* q, const f2* x) {
  f2 s = fma2(q[N], x[N], q[N + 1]);
#pragma unroll
  for (int i = N - 1; i >= 0; i--) s = fma2(q[i], x[i], s);
  return s;
}
template <int N> struct EvalDense {
  static constexpr bool kHasAbsForm = true;
  static constexpr int kThr = -1;   // threshold = delta for every hypothesis
  static constexpr bool kShifted = true;
  __device__ static __forceinline__ f2 dist(const f2* q, const f2* x) { return dense_dist<N>(q, x); }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2& t) { const f2 s = sub2(dist(q, x), t.delta); return fma2(s, s, t.neg_delta2); }
};
#define LSQR_DEF_(ID, N) template <> struct Eval<ID> : EvalDense<N> {};
LSQR_DENSE_N_LIST(LSQR_DEF_)
#undef LSQR_DEF_

// cross-wire: e = R2 (u c1 + v c2 + t3) + t2 - t1, |e|^2 - delta^2   (21 FMA-pipe operations)
template <> struct Eval<USXW> {
  static constexpr bool kHasAbsForm = false;
  static constexpr int kThr = -1;
  static constexpr bool kShifted = false;
  __device__ static __forceinline__ f2 dist(const f2*, const f2*) { return splat(0.f); }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2& t) {
    f2 w[3];
#pragma unroll
    for (int k = 0; k < 3; k++) w[k] = fma2(x[12], q[k], fma2(x[13], q[3 + k], q[6 + k]));
    f2 g = t.neg_delta2;
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const f2 e = add2(fma2(x[3 * r], w[0], fma2(x[3 * r + 1], w[1], fma2(x[3 * r + 2], w[2], q[9 + r]))), x[9 + r]);
      g = fma2(e, e, g);
    }
    return g;
  }
};

// calibrated pointer: e = R2 (u c1 + v c2 + t3) + (t2 - p); the difference t2 - p is formed in fp64 when the fp32 copy is written
template <> struct Eval<USCP> {
  static constexpr bool kHasAbsForm = false;
  static constexpr int kThr = -1;
  static constexpr bool kShifted = false;
  __device__ static __forceinline__ f2 dist(const f2*, const f2*) { return splat(0.f); }
  __device__ static __forceinline__ f2 signed_(const f2* q, const f2* x, const Thr2& t) {
    f2 w[3];
#pragma unroll
    for (int k = 0; k < 3; k++) w[k] = fma2(x[12], q[k], fma2(x[13], q[3 + k], q[6 + k]));
    f2 g = t.neg_delta2;
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const f2 e = add2(fma2(x[3 * r], w[0], fma2(x[3 * r + 1], w[1], mul2(x[3 * r + 2], w[2]))), x[9 + r]);   // rows 9..11 of the fp32 copy hold t2 - p (ingest_kernel)
      g = fma2(e, e, g);
    }
    return g;
  }
};

template <int M> __device__ __forceinline__ float hyp_thr(const f2* q, float delta) {
  if constexpr (Eval<M>::kThr >= 0) { float lo, hi; halves(q[Eval<M>::kThr], lo, hi); return lo; }
  else return delta;
}
template <int M> __device__ __forceinline__ float hyp_thr(const float* q, float delta) {
  if constexpr (Eval<M>::kThr >= 0) return q[Eval<M>::kThr];
  else return delta;
}
// 2^32 - bits(window) for count_carry: the window is 2 delta for every hypothesis, or the hypothesis' own (hypersphere family)
template <int M, class Q> __device__ __forceinline__ uint32_t hyp_negk(const Q* q, uint32_t negk_delta) {
  if constexpr (Eval<M>::kThr >= 0) return __float_as_uint(hyp_thr<M>(q, 0.f));
  else return negk_delta;
}
// FSETP + predicated IADD on two residuals, written in PTX so that ptxas keeps the 2-instruction form.
__device__ __forceinline__ void count_abs_lt(uint32_t& cnt, f2 s, float delta) {
  float a, b;
  halves(s, a, b);
  asm("{\n\t.reg .pred p0, p1;\n\t"
      "setp.lt.f32 p0, %1, %3;\n\tsetp.lt.f32 p1, %2, %3;\n\t"
      "@p0 add.u32 %0, %0, 1;\n\t@p1 add.u32 %0, %0, 1;\n\t}"
      : "+r"(cnt) : "f"(fabsf(a)), "f"(fabsf(b)), "f"(delta));
}
// Outlier counting through the carry chain, for residuals in the shifted form s' = s + delta (inlier <=> 0 <= s' < 2 delta
// <=> bits(s') < bits(2 delta) as unsigned integers: negative floats, NaN and everything >= 2 delta have larger bit patterns).
// The counter is the HIGH word of a 64-bit value whose low word is replaced by bits(s') before a 64-bit add of
// 2^32 - bits(2 delta): the carry out of the low word is the outlier flag.  ptxas turns two such steps into
//   IADD3 RZ, P0, PT, s0, -K, RZ ;  IADD3 RZ, P1, PT, s1, -K, RZ ;  IADD3.X cnt, PT, PT, RZ, RZ, cnt, P1, P0
// i.e. three ALU instructions per two residuals with ONE register operand each (no value but the counter is written).
__device__ __forceinline__ void count_carry(unsigned long long& acc, f2 s, uint32_t negk) {   // negk = 2^32 - bits(2 delta)
  float a, b;
  halves(s, a, b);
  asm("{\n\t.reg .b64 w, k;\n\t.reg .b32 lo, hi;\n\t"
      "cvt.u64.u32 k, %3;\n\t"
      "mov.b64 {lo, hi}, %0;\n\tmov.b64 w, {%1, hi};\n\tadd.u64 %0, w, k;\n\t"
      "mov.b64 {lo, hi}, %0;\n\tmov.b64 w, {%2, hi};\n\tadd.u64 %0, w, k;\n\t}"
      : "+l"(acc) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(negk));
}
// Three ALU instructions per two residuals: the two FSET.BF results are summed as raw words by one IADD3
// (raw += 0x3F800000 per inlier).  0x3F800000 = 127 << 23, so raw holds (127 k mod 512) << 23 and k is recovered
// by cb_raw_decode() as long as no more than 511 inliers were summed since the last decode.
__device__ __forceinline__ void count_abs_lt4_raw(uint32_t& raw, f2 s01, f2 s23, float delta) {
  float a, b, c, d;
  halves(s01, a, b);
  halves(s23, c, d);
  asm("{\n\t.reg .f32 f0, f1, f2, f3;\n\t.reg .b32 t0, t1, t2, t3;\n\t"
      "set.lt.f32.f32 f0, %1, %5;\n\tset.lt.f32.f32 f1, %2, %5;\n\tset.lt.f32.f32 f2, %3, %5;\n\tset.lt.f32.f32 f3, %4, %5;\n\t"
      "mov.b32 t0, f0;\n\tmov.b32 t1, f1;\n\tmov.b32 t2, f2;\n\tmov.b32 t3, f3;\n\t"
      "add.u32 t0, t0, t1;\n\tadd.u32 %0, %0, t0;\n\tadd.u32 t2, t2, t3;\n\tadd.u32 %0, %0, t2;\n\t}"
      : "+r"(raw) : "f"(fabsf(a)), "f"(fabsf(b)), "f"(fabsf(c)), "f"(fabsf(d)), "f"(delta));
}
__device__ __forceinline__ uint32_t cb_raw_decode(uint32_t raw) { return ((raw >> 23) * 383u) & 511u; }   // 127 * 383 = 1 (mod 512)
__device__ __forceinline__ void count_sign(uint32_t& cnt, f2 g) {
  float a, b;
  halves(g, a, b);
  cnt += __float_as_uint(a) >> 31;
  cnt += __float_as_uint(b) >> 31;
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
// profiles/r01_tune_cb_sweep_models.txt and profiles/r02_tune_cb_blocking.txt.  Two things decide them:
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
#define LSQR_CB_R_PLANE 10
#endif
#ifndef LSQR_CB_PPI_PLANE
#define LSQR_CB_PPI_PLANE 4
#endif
constexpr int cb_blocking(int m, bool want_r) {
#ifdef LSQR_CB_OVR_MODEL      // R&D builds (tools/build_variants.sh): override one model
  if (m == LSQR_CB_OVR_MODEL) return want_r ? LSQR_CB_OVR_R : LSQR_CB_OVR_PPI;
#endif
  int r = 8, ppi = 2;
  switch (m) {
    case PLANE3: r = LSQR_CB_R_PLANE; ppi = LSQR_CB_PPI_PLANE; break;
    case LINE2D: case LINE2: r = 14; ppi = 4; break;   // two counters per hypothesis (mixed count): 16 would spill
    case PLANE2: r = 16; ppi = 4; break;
    case LINE3: r = 8; ppi = 2; break;
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
