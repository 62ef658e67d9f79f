// Levenberg-Marquardt controller of the refine path, on the device.
//
// The reference runs every nonlinear refit through vnl_levenberg_marquardt::minimize
// (SphereParametersEstimator.hxx:321-331, SinglePointTargetUSCalibrationParametersEstimator.cxx:284-297, :928-941), which is
// MINPACK's lmder (J. J. More', "The Levenberg-Marquardt algorithm: implementation and theory", LNM 630, 1978; More', Garbow,
// Hillstrom, ANL-80-74): mode 1 (internal scaling by the column norms of the Jacobian), factor 100, a trust region on the
// scaled step, the lmpar iteration for the damping parameter, the actual/predicted reduction ratio and four stopping
// tests; minimize() succeeds iff MINPACK's info is 1..4.  Where the reference's tolerances stop the iteration short of the
// minimiser (the calibrated-pointer calibration uses 1e-7), its result is whatever that iteration has reached, so the
// controller here follows the same algorithm step for step instead of "some LM": same trust-region updates, same tests,
// same evaluation count.
//
// What differs is the data flow.  MINPACK factors the m x n Jacobian (m = number of data) by Householder QR with column
// pivoting; here one streaming pass over the data (mask_moments_kernel) delivers A = J^T J, g = J^T f and |f|^2 at the
// evaluation point, and the controller takes R from a Cholesky factorisation of P^T A P with the same pivoting rule
// (largest remaining column norm) and q = R^-T P^T g, which are the R and the first n components of Q^T f of that QR.
// The factorisation works on the diagonally scaled matrix, so its accuracy does not suffer from the spread of parameter
// units (millimetres, radians, pixel scales).  State lives in a small device buffer; the host never sees an iterate.
#pragma once

#include <math.h>
#ifndef __CUDACC__   // the controller is plain C++: tests/lm_controller_check.cxx runs it on the host against MINPACK's lmder
#define __host__
#define __device__
#endif

namespace lsqr {

constexpr int kLmMaxP = 11;
// layout of the controller state (doubles)
enum : int {
  LM_X = 0,                      // [kLmMaxP] current point
  LM_TRIAL = 11,                 // [kLmMaxP] point whose evaluation is pending
  LM_STATUS = 22,                // 0 run, 1 converged (MINPACK info 1..4), 2 failed
  LM_PHASE = 23,                 // 0 = the pending evaluation is at x, 1 = at the trial point
  LM_EVALS = 24,                 // nfev
  LM_ITER = 25, LM_PAR = 26, LM_DELTA = 27, LM_PNORM = 28, LM_FNORM = 29, LM_XNORM = 30, LM_GNORM = 31,
  LM_INFO = 32,                  // MINPACK's info once stopped
  LM_DIAG = 33,                  // [kLmMaxP] scaling
  LM_QTF = 44,                   // [kLmMaxP] first n components of Q^T f
  LM_STEP = 55,                  // [kLmMaxP] step to the trial point
  LM_IPVT = 66,                  // [kLmMaxP] pivot order
  LM_ACN = 77,                   // [kLmMaxP] column norms of J at x
  LM_R = 88,                     // [kLmMaxP * kLmMaxP] R of J P = Q R at x (row-major, upper triangle; the lower one is scratch)
  LM_SIZE = 209
};

struct LmTol { double ftol, xtol, gtol; int maxfev; };
// The reference's settings (vnl_levenberg_marquardt setters at the cited lines; what is not set keeps VNL's default):
//   kind 0  circle / sphere geometric fit (SphereParametersEstimator.hxx:323-329): xtol = gtol = 10e-16, ftol = VNL's default
//           xtol_default * 0.01 = 1e-10, at most 500 function evaluations;
//   kind 1  cross-wire calibration (SinglePointTargetUSCalibrationParametersEstimator.cxx:287-295): all three 10e-16, 5000;
//   kind 2  calibrated-pointer calibration (:931-939): all three 10e-8, 5000.
__host__ __device__ inline LmTol lm_tolerances(int kind) {
  if (kind == 1) return LmTol{10e-16, 10e-16, 10e-16, 5000};
  if (kind == 2) return LmTol{10e-8, 10e-8, 10e-8, 5000};
  return LmTol{1e-8 * 0.01, 10e-16, 10e-16, 500};
}

template <int N> __host__ __device__ inline double lm_norm(const double* v) {
  double s = 0.0;
  for (int i = 0; i < N; i++) s += v[i] * v[i];
  return sqrt(s);
}

// R, ipvt, acnorm, qtf from the normal equations A (row-major, symmetric) and g.
template <int N> __host__ __device__ inline void lm_factor(const double* A, const double* g, double* R, double* ipvt, double* acn, double* qtf) {
  double S[N * N], sc[N], rem[N];
  int piv[N];
  for (int j = 0; j < N; j++) { acn[j] = sqrt(A[j * N + j] > 0.0 ? A[j * N + j] : 0.0); sc[j] = acn[j] > 0.0 ? 1.0 / acn[j] : 0.0; piv[j] = j; }
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) S[i * N + j] = A[i * N + j] * sc[i] * sc[j];
  for (int j = 0; j < N; j++) rem[j] = sc[j] > 0.0 ? 1.0 : 0.0;          // remaining squared norm of scaled column j
  for (int i = 0; i < N * N; i++) R[i] = 0.0;
  // Rs (upper triangular, of the scaled, permuted matrix) is built row by row in the rows of R, in pivot positions
  for (int j = 0; j < N; j++) {
    int kmax = j;
    for (int k = j; k < N; k++) if (rem[piv[k]] * acn[piv[k]] * acn[piv[k]] > rem[piv[kmax]] * acn[piv[kmax]] * acn[piv[kmax]]) kmax = k;
    { const int t = piv[j]; piv[j] = piv[kmax]; piv[kmax] = t; }
    for (int c = 0; c < j; c++) { const double t = R[c * N + j]; R[c * N + j] = R[c * N + kmax]; R[c * N + kmax] = t; }   // swap the columns already built
    const int pj = piv[j];
    double d = S[pj * N + pj];
    for (int c = 0; c < j; c++) d -= R[c * N + j] * R[c * N + j];
    if (!(d > 0.0)) break;                                                // the remaining columns are (numerically) dependent: R rows stay zero
    const double rjj = sqrt(d);
    R[j * N + j] = rjj;
    for (int k = j + 1; k < N; k++) {
      const int pk = piv[k];
      double t = S[pj * N + pk];
      for (int c = 0; c < j; c++) t -= R[c * N + j] * R[c * N + k];
      t /= rjj;
      R[j * N + k] = t;
      rem[pk] -= t * t;
      if (rem[pk] < 0.0) rem[pk] = 0.0;
    }
  }
  // unscale the columns: R = Rs * diag(acn[piv]);  qtf = R^-T (P^T g)
  for (int i = 0; i < N; i++) for (int k = i; k < N; k++) R[i * N + k] *= acn[piv[k]];
  for (int j = 0; j < N; j++) {
    double t = g[piv[j]];
    for (int c = 0; c < j; c++) t -= R[c * N + j] * qtf[c];
    qtf[j] = R[j * N + j] != 0.0 ? t / R[j * N + j] : 0.0;
    ipvt[j] = (double)piv[j];
  }
}

// MINPACK qrsolv: least-squares solution of [R P^T; D] x = [qtb; 0].  r(i,j) = R[i*N+j]; the strict lower triangle of R is scratch.
template <int N> __host__ __device__ inline void lm_qrsolv(double* R, const int* ipvt, const double* diag, const double* qtb, double* x, double* sdiag) {
  double wa[N];
  for (int j = 0; j < N; j++) {
    for (int i = j; i < N; i++) R[i * N + j] = R[j * N + i];
    x[j] = R[j * N + j];
    wa[j] = qtb[j];
  }
  for (int j = 0; j < N; j++) {
    const int l = ipvt[j];
    if (diag[l] != 0.0) {
      for (int k = j; k < N; k++) sdiag[k] = 0.0;
      sdiag[j] = diag[l];
      double qtbpj = 0.0;
      for (int k = j; k < N; k++) {
        if (sdiag[k] == 0.0) continue;
        double c, s;
        if (fabs(R[k * N + k]) < fabs(sdiag[k])) { const double ct = R[k * N + k] / sdiag[k]; s = 0.5 / sqrt(0.25 + 0.25 * (ct * ct)); c = s * ct; }
        else { const double tn = sdiag[k] / R[k * N + k]; c = 0.5 / sqrt(0.25 + 0.25 * (tn * tn)); s = c * tn; }
        R[k * N + k] = c * R[k * N + k] + s * sdiag[k];
        const double t = c * wa[k] + s * qtbpj;
        qtbpj = -s * wa[k] + c * qtbpj;
        wa[k] = t;
        for (int i = k + 1; i < N; i++) {
          const double u = c * R[i * N + k] + s * sdiag[i];
          sdiag[i] = -s * R[i * N + k] + c * sdiag[i];
          R[i * N + k] = u;
        }
      }
    }
    sdiag[j] = R[j * N + j];
    R[j * N + j] = x[j];
  }
  int nsing = N;
  for (int j = 0; j < N; j++) {
    if (sdiag[j] == 0.0 && nsing == N) nsing = j;
    if (nsing < N) wa[j] = 0.0;
  }
  for (int k = 1; k <= nsing; k++) {
    const int j = nsing - k;
    double sum = 0.0;
    for (int i = j + 1; i < nsing; i++) sum += R[i * N + j] * wa[i];
    wa[j] = (wa[j] - sum) / sdiag[j];
  }
  for (int j = 0; j < N; j++) x[ipvt[j]] = wa[j];
}

// MINPACK lmpar: the damping parameter with | |D x| - delta | <= 0.1 delta, and the step x.
template <int N> __host__ __device__ inline void lm_lmpar(double* R, const int* ipvt, const double* diag, const double* qtb, double delta, double* par, double* x) {
  const double dwarf = 2.2250738585072014e-308;
  double wa1[N], wa2[N], sdiag[N];
  int nsing = N;
  for (int j = 0; j < N; j++) {
    wa1[j] = qtb[j];
    if (R[j * N + j] == 0.0 && nsing == N) nsing = j;
    if (nsing < N) wa1[j] = 0.0;
  }
  for (int k = 1; k <= nsing; k++) {
    const int j = nsing - k;
    wa1[j] /= R[j * N + j];
    const double t = wa1[j];
    for (int i = 0; i < j; i++) wa1[i] -= R[i * N + j] * t;
  }
  for (int j = 0; j < N; j++) x[ipvt[j]] = wa1[j];
  for (int j = 0; j < N; j++) wa2[j] = diag[j] * x[j];
  double dxnorm = lm_norm<N>(wa2), fp = dxnorm - delta;
  if (fp <= 0.1 * delta) { *par = 0.0; return; }
  double parl = 0.0;
  if (nsing >= N) {
    for (int j = 0; j < N; j++) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
    for (int j = 0; j < N; j++) {
      double sum = 0.0;
      for (int i = 0; i < j; i++) sum += R[i * N + j] * wa1[i];
      wa1[j] = (wa1[j] - sum) / R[j * N + j];
    }
    const double t = lm_norm<N>(wa1);
    parl = ((fp / delta) / t) / t;
  }
  for (int j = 0; j < N; j++) {
    double sum = 0.0;
    for (int i = 0; i <= j; i++) sum += R[i * N + j] * qtb[i];
    wa1[j] = sum / diag[ipvt[j]];
  }
  const double gnorm = lm_norm<N>(wa1);
  double paru = gnorm / delta;
  if (paru == 0.0) paru = dwarf / (delta < 0.1 ? delta : 0.1);
  if (*par < parl) *par = parl;
  if (*par > paru) *par = paru;
  if (*par == 0.0) *par = gnorm / dxnorm;
  for (int iter = 1;; iter++) {
    if (*par == 0.0) { const double t = 0.001 * paru; *par = dwarf > t ? dwarf : t; }
    double t = sqrt(*par);
    for (int j = 0; j < N; j++) wa1[j] = t * diag[j];
    lm_qrsolv<N>(R, ipvt, wa1, qtb, x, sdiag);
    for (int j = 0; j < N; j++) wa2[j] = diag[j] * x[j];
    dxnorm = lm_norm<N>(wa2);
    t = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= t && t < 0.0) || iter == 10) break;
    for (int j = 0; j < N; j++) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
    for (int j = 0; j < N; j++) {
      wa1[j] /= sdiag[j];
      const double u = wa1[j];
      for (int i = j + 1; i < N; i++) wa1[i] -= R[i * N + j] * u;
    }
    t = lm_norm<N>(wa1);
    const double parc = ((fp / delta) / t) / t;
    if (fp > 0.0 && *par > parl) parl = *par;
    if (fp < 0.0 && *par < paru) paru = *par;
    t = *par + parc;
    *par = parl > t ? parl : t;
  }
}

// moments of one evaluation pass: [0] count, J^T J upper triangle (row-major), J^T f, |f|^2
template <int N> __host__ __device__ inline void lm_load_normal(const double* m, double* A, double* g, double* cost) {
  int o = 1;
  for (int a = 0; a < N; a++) for (int b = a; b < N; b++) { const double v = m[o++]; A[a * N + b] = v; A[b * N + a] = v; }
  for (int a = 0; a < N; a++) g[a] = m[o++];
  *cost = m[o];
}

// start of an outer iteration of lmder at st[LM_X] with the normal equations (A, g): factor, scale, gradient test.
// Returns false when the iteration has stopped (status set).
template <int N> __host__ __device__ inline bool lm_outer(const double* A, const double* g, double* st, const LmTol& tol) {
  lm_factor<N>(A, g, st + LM_R, st + LM_IPVT, st + LM_ACN, st + LM_QTF);
  const double* acn = st + LM_ACN;
  double* diag = st + LM_DIAG;
  if (st[LM_ITER] == 1.0) {
    double wa[N];
    for (int j = 0; j < N; j++) { diag[j] = acn[j] == 0.0 ? 1.0 : acn[j]; wa[j] = diag[j] * st[LM_X + j]; }
    st[LM_XNORM] = lm_norm<N>(wa);
    st[LM_DELTA] = 100.0 * st[LM_XNORM];                                  // factor = 100 (VNL)
    if (st[LM_DELTA] == 0.0) st[LM_DELTA] = 100.0;
  }
  double gnorm = 0.0;
  const double fnorm = st[LM_FNORM];
  if (fnorm != 0.0) {
    for (int j = 0; j < N; j++) {
      const int l = (int)st[LM_IPVT + j];
      if (acn[l] != 0.0) {
        double sum = 0.0;
        for (int i = 0; i <= j; i++) sum += st[LM_R + i * N + j] * (st[LM_QTF + i] / fnorm);
        const double v = fabs(sum / acn[l]);
        if (v > gnorm) gnorm = v;
      }
    }
  }
  st[LM_GNORM] = gnorm;
  if (gnorm <= tol.gtol) { st[LM_INFO] = 4.0; st[LM_STATUS] = 1.0; return false; }
  for (int j = 0; j < N; j++) if (acn[j] > diag[j]) diag[j] = acn[j];
  return true;
}

// inner iteration: damping parameter, step, trial point; asks for the evaluation of the trial point
template <int N> __host__ __device__ inline void lm_inner(double* st) {
  int ipvt[N];
  double p[N], wa3[N];
  for (int j = 0; j < N; j++) ipvt[j] = (int)st[LM_IPVT + j];
  double par = st[LM_PAR];
  lm_lmpar<N>(st + LM_R, ipvt, st + LM_DIAG, st + LM_QTF, st[LM_DELTA], &par, p);
  st[LM_PAR] = par;
  for (int j = 0; j < N; j++) { st[LM_STEP + j] = -p[j]; st[LM_TRIAL + j] = st[LM_X + j] - p[j]; wa3[j] = st[LM_DIAG + j] * (-p[j]); }
  st[LM_PNORM] = lm_norm<N>(wa3);
  if (st[LM_ITER] == 1.0 && st[LM_PNORM] < st[LM_DELTA]) st[LM_DELTA] = st[LM_PNORM];
  st[LM_PHASE] = 1.0;
}

// One controller step: consumes the moments of the evaluation that was pending and either stops or leaves the next
// evaluation point in the state.
template <int N> __host__ __device__ inline void lm_update(const double* m, double* st, const LmTol& tol) {
  const double epsmch = 2.220446049250313e-16;
  if (st[LM_STATUS] != 0.0) return;
  double A[N * N], g[N], cost;
  lm_load_normal<N>(m, A, g, &cost);
  if (st[LM_PHASE] == 0.0) {                                              // the initial evaluation
    if (m[0] < (double)N) { st[LM_STATUS] = 2.0; return; }                // vnl_levenberg_marquardt: fewer residuals than unknowns -> failure
    st[LM_EVALS] = 1.0; st[LM_ITER] = 1.0; st[LM_PAR] = 0.0;
    st[LM_FNORM] = sqrt(cost);
    if (!lm_outer<N>(A, g, st, tol)) return;
    lm_inner<N>(st);
    return;
  }
  st[LM_EVALS] += 1.0;
  const double fnorm = st[LM_FNORM], fnorm1 = sqrt(cost), pnorm = st[LM_PNORM];
  double actred = -1.0;
  if (0.1 * fnorm1 < fnorm) { const double t = fnorm1 / fnorm; actred = 1.0 - t * t; }
  double wa3[N];
  for (int j = 0; j < N; j++) wa3[j] = 0.0;
  for (int j = 0; j < N; j++) {
    const double t = st[LM_STEP + (int)st[LM_IPVT + j]];
    for (int i = 0; i <= j; i++) wa3[i] += st[LM_R + i * N + j] * t;
  }
  const double temp1 = lm_norm<N>(wa3) / fnorm, temp2 = (sqrt(st[LM_PAR]) * pnorm) / fnorm;
  const double prered = temp1 * temp1 + temp2 * temp2 / 0.5, dirder = -(temp1 * temp1 + temp2 * temp2);
  const double ratio = prered != 0.0 ? actred / prered : 0.0;
  if (ratio <= 0.25) {
    double t = actred >= 0.0 ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
    if (0.1 * fnorm1 >= fnorm || t < 0.1) t = 0.1;
    st[LM_DELTA] = t * (st[LM_DELTA] < pnorm / 0.1 ? st[LM_DELTA] : pnorm / 0.1);
    st[LM_PAR] /= t;
  } else if (st[LM_PAR] == 0.0 || ratio >= 0.75) {
    st[LM_DELTA] = pnorm / 0.5;
    st[LM_PAR] *= 0.5;
  }
  const bool success = ratio >= 1e-4;
  if (success) {
    double wa2[N];
    for (int j = 0; j < N; j++) { st[LM_X + j] = st[LM_TRIAL + j]; wa2[j] = st[LM_DIAG + j] * st[LM_X + j]; }
    st[LM_XNORM] = lm_norm<N>(wa2);
    st[LM_FNORM] = fnorm1;
    st[LM_ITER] += 1.0;
  }
  int info = 0;
  const bool fsmall = fabs(actred) <= tol.ftol && prered <= tol.ftol && 0.5 * ratio <= 1.0;
  if (fsmall) info = 1;
  if (st[LM_DELTA] <= tol.xtol * st[LM_XNORM]) info = 2;
  if (fsmall && info == 2) info = 3;
  if (info == 0) {
    if (st[LM_EVALS] >= (double)tol.maxfev) info = 5;
    if (fabs(actred) <= epsmch && prered <= epsmch && 0.5 * ratio <= 1.0) info = 6;
    if (st[LM_DELTA] <= epsmch * st[LM_XNORM]) info = 7;
    if (st[LM_GNORM] <= epsmch) info = 8;
  }
  if (info != 0) { st[LM_INFO] = (double)info; st[LM_STATUS] = info <= 4 ? 1.0 : 2.0; return; }
  if (success && !lm_outer<N>(A, g, st, tol)) return;                     // new Jacobian at the accepted point (same pass delivered it)
  lm_inner<N>(st);
}

}  // namespace lsqr
