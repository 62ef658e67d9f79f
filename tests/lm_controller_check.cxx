// TEST TOOL (host only, built by tests/test_lm_cpu.py with g++): runs the engine's Levenberg-Marquardt controller
// (lsqrrecipes_b200/csrc/lm_minpack.cuh -- plain C++ when not compiled by nvcc) on the CPU, feeding it the moments
// J^T J, J^T f, |f|^2 that the streaming pass would deliver, next to MINPACK's lmder on the full m x n Jacobian
// (oracle/minpack_lm.h) for the same functor, and reports both end points, evaluation counts and MINPACK info codes.
// The functors are the oracle's restatements of the reference's f / gradf.
#include <cstdio>
#include <cstring>
#include <vector>

extern "C" {
#include "../oracle/lsqr_oracle.c"
}
#include "../lsqrrecipes_b200/csrc/lm_minpack.cuh"

namespace {

template <int N> struct Run {
  // moments layout of one pass: count, J^T J upper triangle, J^T f, |f|^2
  static void moments(mpk_fcn fcn, void* user, int m, const double* x, std::vector<double>& f, std::vector<double>& J, double* mom) {
    fcn(user, m, N, x, f.data(), J.data(), m, 1);
    fcn(user, m, N, x, f.data(), J.data(), m, 2);
    int o = 0;
    mom[o++] = (double)m;
    for (int a = 0; a < N; a++) for (int b = a; b < N; b++) { double s = 0; for (int i = 0; i < m; i++) s += J[i + (size_t)m * a] * J[i + (size_t)m * b]; mom[o++] = s; }
    for (int a = 0; a < N; a++) { double s = 0; for (int i = 0; i < m; i++) s += J[i + (size_t)m * a] * f[i]; mom[o++] = s; }
    double c = 0; for (int i = 0; i < m; i++) c += f[i] * f[i];
    mom[o] = c;
  }
  static void go(mpk_fcn fcn, void* user, int m, int kind, const double* init, double* out) {
    const lsqr::LmTol tol = lsqr::lm_tolerances(kind);
    // (a) MINPACK lmder on the full Jacobian
    std::vector<double> x(init, init + N), fvec(m);
    int nfev = 0, njev = 0;
    const int info = mpk_lmder(fcn, user, m, N, x.data(), fvec.data(), tol.ftol, tol.xtol, tol.gtol, tol.maxfev, 100.0, &nfev, &njev);
    // (b) the engine's controller on the moments
    double st[lsqr::LM_SIZE];
    for (int i = 0; i < lsqr::LM_SIZE; i++) st[i] = 0.0;
    for (int j = 0; j < N; j++) st[lsqr::LM_X + j] = init[j];
    std::vector<double> f(m), J((size_t)m * N);
    double mom[1 + N * (N + 1) / 2 + N + 1];
    int passes = 0;
    while (st[lsqr::LM_STATUS] == 0.0 && passes < 6000) {
      const double* at = st + (st[lsqr::LM_PHASE] != 0.0 ? lsqr::LM_TRIAL : lsqr::LM_X);
      moments(fcn, user, m, at, f, J, mom);
      lsqr::lm_update<N>(mom, st, tol);
      passes++;
    }
    int o = 0;
    out[o++] = info; out[o++] = nfev; out[o++] = st[lsqr::LM_INFO]; out[o++] = st[lsqr::LM_EVALS]; out[o++] = st[lsqr::LM_STATUS]; out[o++] = passes;
    for (int j = 0; j < N; j++) out[o++] = x[j];
    for (int j = 0; j < N; j++) out[o++] = st[lsqr::LM_X + j];
  }
};

}  // namespace

// model: oracle model id; data: packed inliers; init: start (the analytic estimate's leading entries).
// out: [info_minpack, nfev_minpack, info_ctrl, nfev_ctrl, status_ctrl, passes, x_minpack[N], x_ctrl[N]]
extern "C" int lmchk_run(int model, const double* data, int n, const double* init, double* out) {
  if (model == M_CIRCLE2 || model == M_SPHERE3 || model == M_SPHERE4) {
    sphere_lm_ctx c; c.d = data; c.dim = model == M_CIRCLE2 ? 2 : (model == M_SPHERE3 ? 3 : 4);
    if (c.dim == 2) Run<3>::go(sphere_lm_fcn, &c, n, 0, init, out);
    else if (c.dim == 3) Run<4>::go(sphere_lm_fcn, &c, n, 0, init, out);
    else Run<5>::go(sphere_lm_fcn, &c, n, 0, init, out);
    return c.dim + 1;
  }
  us_lm_ctx c; c.d = data; c.n = (size_t)n;
  if (model == M_USXW) { Run<11>::go(usxw_lm_fcn, &c, n, 1, init, out); return 11; }
  if (model == M_USCP) { Run<8>::go(uscp_lm_fcn, &c, n, 2, init, out); return 8; }
  return -1;
}

// The functor alone (for running the REAL MINPACK -- scipy.optimize.leastsq -- on bit-identical residuals and Jacobians):
// iflag 1 -> fvec[n], iflag 2 -> fjac[n][N] row-major.
extern "C" int lmchk_eval(int model, const double* data, int n, const double* x, int iflag, double* out) {
  int N = 0;
  mpk_fcn fcn = 0;
  sphere_lm_ctx sc; us_lm_ctx uc;
  void* user = 0;
  if (model == M_CIRCLE2 || model == M_SPHERE3 || model == M_SPHERE4) { sc.d = data; sc.dim = model == M_CIRCLE2 ? 2 : (model == M_SPHERE3 ? 3 : 4); N = sc.dim + 1; fcn = sphere_lm_fcn; user = &sc; }
  else if (model == M_USXW) { uc.d = data; uc.n = (size_t)n; N = 11; fcn = usxw_lm_fcn; user = &uc; }
  else if (model == M_USCP) { uc.d = data; uc.n = (size_t)n; N = 8; fcn = uscp_lm_fcn; user = &uc; }
  else return -1;
  if (iflag == 1) { std::vector<double> J(1); fcn(user, n, N, x, out, J.data(), n, 1); return N; }
  std::vector<double> f(n), J((size_t)n * N);
  fcn(user, n, N, x, f.data(), J.data(), n, 2);
  for (int i = 0; i < n; i++) for (int j = 0; j < N; j++) out[(size_t)i * N + j] = J[i + (size_t)n * j];
  return N;
}
