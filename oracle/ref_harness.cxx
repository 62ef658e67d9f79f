// TEST INFRASTRUCTURE ONLY -- not part of the shipped engine, never on the product path.
//
// C-callable harness around the UNMODIFIED reference sources (zivy/LSQRRecipes), compiled
// from where they lie under /root/reference against oracle/vnl_shim (VNL is absent from the
// image; see vnl_shim_core.h for what that implies).  Built by oracle/Makefile into
// oracle/_ref/libref_oracle.so.  No reference source is copied into this repository: this
// file only #includes the reference headers and calls their public API.
//
// Used (a) to pin the plain-C restatement in oracle/lsqr_oracle.c, (b) to mint the golden
// fixtures in tests/golden/, (c) as the "reference" CPU baseline in bench.py.
//
// Data crosses this boundary as packed doubles, D per datum (see ref_model_info).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "Point2D.h"
#include "Point3D.h"
#include "Frame.h"
#include "Ray3D.h"
#include "RANSAC.h"
#include "PlaneParametersEstimator.h"
#include "LineParametersEstimator.h"
#include "Line2DParametersEstimator.h"
#include "SphereParametersEstimator.h"
#include "AbsoluteOrientationParametersEstimator.h"
#include "RayIntersectionParametersEstimator.h"
#include "PivotCalibrationParametersEstimator.h"
#include "DenseLinearEquationSystemParametersEstimator.h"
#include "SinglePointTargetUSCalibrationParametersEstimator.h"

using namespace lsqrRecipes;

enum { M_PLANE3 = 0, M_LINE2D = 1, M_LINE2 = 2, M_LINE3 = 3, M_CIRCLE2 = 4, M_SPHERE3 = 5, M_ABSOR = 6, M_RAY = 7, M_PIVOT = 8, M_DENSE5 = 9, M_DENSE6 = 10, M_USXW = 11, M_USCP = 12, M_SPHERE4 = 13, M_PLANE4 = 14,
       // further instantiations of the reference's dimension-templated estimators (ids of lsqr_oracle.c / include/lsqr_b200.h)
       M_PLANE2 = 15, M_PLANE5 = 16, M_PLANE6 = 17, M_PLANE7 = 18, M_PLANE8 = 19, M_SPHERE5 = 20, M_SPHERE6 = 21, M_SPHERE7 = 22, M_SPHERE8 = 23,
       M_LINE4 = 24, M_LINE5 = 25, M_LINE6 = 26, M_LINE7 = 27, M_LINE8 = 28, M_DENSE2 = 29, M_DENSE3 = 30, M_DENSE4 = 31, M_DENSE7 = 32, M_DENSE8 = 33, M_COUNT = 34 };

namespace {

typedef std::pair<Point3D, Point3D> PointPair;

template <class T> struct Marshal;
template <unsigned n> struct Marshal<Point<double, n> > {
  enum { D = n };
  static Point<double, n> get(const double* p) { Point<double, n> r; for (unsigned i = 0; i < n; i++) r[i] = p[i]; return r; }
};
template <> struct Marshal<PointPair> {
  enum { D = 6 };
  static PointPair get(const double* p) { PointPair r; for (int i = 0; i < 3; i++) { r.first[i] = p[i]; r.second[i] = p[3 + i]; } return r; }
};
template <> struct Marshal<Ray3D> {
  enum { D = 6 };
  static Ray3D get(const double* p) { Ray3D r; for (int i = 0; i < 3; i++) { r.p[i] = p[i]; r.n[i] = p[3 + i]; } return r; }
};
template <> struct Marshal<Frame> {
  enum { D = 12 };
  static Frame get(const double* p) {
    double R[3][3], t[3];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[i][j] = p[3 * i + j]; t[i] = p[9 + i]; }
    return Frame(R, t);
  }
};

template <unsigned n> struct Marshal<AugmentedRow<double, n> > {
  enum { D = n + 1 };
  static AugmentedRow<double, n> get(const double* p) { double tmp[n + 1]; for (unsigned i = 0; i <= n; i++) tmp[i] = p[i]; return AugmentedRow<double, n>(tmp); }
};

typedef SingleUnknownPointTargetUSCalibrationParametersEstimator::DataType CrossWireDatum;
template <> struct Marshal<CrossWireDatum> {
  enum { D = 14 };
  static CrossWireDatum get(const double* p) {
    double R[3][3], t[3];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[i][j] = p[3 * i + j]; t[i] = p[9 + i]; }
    CrossWireDatum d;
    d.T2 = Frame(R, t);
    d.q[0] = p[12]; d.q[1] = p[13];
    return d;
  }
};

typedef CalibratedPointerTargetUSCalibrationParametersEstimator::DataType PointerDatum;
template <> struct Marshal<PointerDatum> {
  enum { D = 17 };
  static PointerDatum get(const double* p) {
    double R[3][3], t[3];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[i][j] = p[3 * i + j]; t[i] = p[9 + i]; }
    PointerDatum d;
    d.T2 = Frame(R, t);
    d.q[0] = p[12]; d.q[1] = p[13];
    for (int i = 0; i < 3; i++) d.p[i] = p[14 + i];
    return d;
  }
};

template <class T> void unpack(const double* data, size_t n, std::vector<T>& out) {
  out.clear(); out.reserve(n);
  for (size_t i = 0; i < n; i++) out.push_back(Marshal<T>::get(data + i * Marshal<T>::D));
}

struct Cfg { int model; double delta; double aux; int ls_type; };

// Calls f(estimator*, (T*)0) with the concrete reference estimator for cfg.model.
template <class F> int dispatch(const Cfg& c, F& f) {
  switch (c.model) {
    case M_PLANE3: { PlaneParametersEstimator<3> e(c.delta); return f(&e, (Point3D*)0); }
    case M_PLANE4: { typedef Point<double, 4> Point4D; PlaneParametersEstimator<4> e(c.delta); return f(&e, (Point4D*)0); }
    case M_LINE2D: { Line2DParametersEstimator e(c.delta); return f(&e, (Point2D*)0); }
    case M_LINE2: { LineParametersEstimator<2> e(c.delta); return f(&e, (Point2D*)0); }
    case M_LINE3: { LineParametersEstimator<3> e(c.delta); return f(&e, (Point3D*)0); }
    case M_CIRCLE2: { SphereParametersEstimator<2> e(c.delta, c.ls_type == 0 ? SphereParametersEstimator<2>::ALGEBRAIC : SphereParametersEstimator<2>::GEOMETRIC); return f(&e, (Point2D*)0); }
    case M_SPHERE3: { SphereParametersEstimator<3> e(c.delta, c.ls_type == 0 ? SphereParametersEstimator<3>::ALGEBRAIC : SphereParametersEstimator<3>::GEOMETRIC); return f(&e, (Point3D*)0); }
    case M_SPHERE4: { typedef Point<double, 4> Point4D; SphereParametersEstimator<4> e(c.delta, c.ls_type == 0 ? SphereParametersEstimator<4>::ALGEBRAIC : SphereParametersEstimator<4>::GEOMETRIC); return f(&e, (Point4D*)0); }
    case M_ABSOR: { AbsoluteOrientationParametersEstimator e(c.delta); return f(&e, (PointPair*)0); }
    case M_RAY: { if (c.aux > 0) { RayIntersectionParametersEstimator e(c.delta, c.aux); return f(&e, (Ray3D*)0); } RayIntersectionParametersEstimator e(c.delta); return f(&e, (Ray3D*)0); }
    case M_PIVOT: { PivotCalibrationEstimator e(c.delta); return f(&e, (Frame*)0); }
    case M_DENSE5: { DenseLinearEquationSystemParametersEstimator<double, 5> e(c.delta); return f(&e, (AugmentedRow<double, 5>*)0); }
    case M_DENSE6: { DenseLinearEquationSystemParametersEstimator<double, 6> e(c.delta); return f(&e, (AugmentedRow<double, 6>*)0); }
#define REF_PLANE(ID, DIM) case ID: { PlaneParametersEstimator<DIM> e(c.delta); return f(&e, (Point<double, DIM>*)0); }
    REF_PLANE(M_PLANE2, 2) REF_PLANE(M_PLANE5, 5) REF_PLANE(M_PLANE6, 6) REF_PLANE(M_PLANE7, 7) REF_PLANE(M_PLANE8, 8)
#define REF_SPHERE(ID, DIM) case ID: { SphereParametersEstimator<DIM> e(c.delta, c.ls_type == 0 ? SphereParametersEstimator<DIM>::ALGEBRAIC : SphereParametersEstimator<DIM>::GEOMETRIC); return f(&e, (Point<double, DIM>*)0); }
    REF_SPHERE(M_SPHERE5, 5) REF_SPHERE(M_SPHERE6, 6) REF_SPHERE(M_SPHERE7, 7) REF_SPHERE(M_SPHERE8, 8)
#define REF_LINE(ID, DIM) case ID: { LineParametersEstimator<DIM> e(c.delta); return f(&e, (Point<double, DIM>*)0); }
    REF_LINE(M_LINE4, 4) REF_LINE(M_LINE5, 5) REF_LINE(M_LINE6, 6) REF_LINE(M_LINE7, 7) REF_LINE(M_LINE8, 8)
#define REF_DENSE(ID, N) case ID: { DenseLinearEquationSystemParametersEstimator<double, N> e(c.delta); return f(&e, (AugmentedRow<double, N>*)0); }
    REF_DENSE(M_DENSE2, 2) REF_DENSE(M_DENSE3, 3) REF_DENSE(M_DENSE4, 4) REF_DENSE(M_DENSE7, 7) REF_DENSE(M_DENSE8, 8)
    case M_USXW: {
      SingleUnknownPointTargetUSCalibrationParametersEstimator e(c.delta, c.ls_type == 0 ? SingleUnknownPointTargetUSCalibrationParametersEstimator::ANALYTIC
                                                                                         : SingleUnknownPointTargetUSCalibrationParametersEstimator::ITERATIVE);
      return f(&e, (CrossWireDatum*)0);
    }
    case M_USCP: {
      CalibratedPointerTargetUSCalibrationParametersEstimator e(c.delta, c.ls_type == 0 ? CalibratedPointerTargetUSCalibrationParametersEstimator::ANALYTIC
                                                                                        : CalibratedPointerTargetUSCalibrationParametersEstimator::ITERATIVE);
      return f(&e, (PointerDatum*)0);
    }
  }
  return -1;
}

struct EstimateOp {
  const double* data; size_t n; double* params;
  template <class T> int operator()(ParametersEstimator<T, double>* e, T*) {
    std::vector<T> d; unpack(data, n, d);
    std::vector<double> p; e->estimate(d, p);
    for (size_t i = 0; i < p.size(); i++) params[i] = p[i];
    return (int)p.size();
  }
};
struct LsqOp {
  const double* data; size_t n; double* params;
  template <class T> int operator()(ParametersEstimator<T, double>* e, T*) {
    std::vector<T> d; unpack(data, n, d);
    std::vector<double> p; e->leastSquaresEstimate(d, p);
    for (size_t i = 0; i < p.size(); i++) params[i] = p[i];
    return (int)p.size();
  }
};
struct AgreeOp {
  const double* params; int np; const double* data; size_t n; uint8_t* out;
  template <class T> int operator()(ParametersEstimator<T, double>* e, T*) {
    std::vector<T> d; unpack(data, n, d);
    std::vector<double> p(params, params + np);
    int c = 0;
    for (size_t i = 0; i < n; i++) { bool a = e->agree(p, d[i]); if (out) out[i] = a; c += a; }
    return c;
  }
};
// Full (no early exit) scoring of an ordered subset list: the per-subset body of the
// reference's exhaustive driver (RANSAC.hxx:217-249), hypotheses parallelised with OpenMP.
struct ScoreOp {
  const double* data; size_t n; const int32_t* subsets; size_t H; int k; int P;
  uint32_t* counts; double* params_out; int nthreads;
  template <class T> int operator()(ParametersEstimator<T, double>* e, T*) {
    std::vector<T> d; unpack(data, n, d);
    const double nan = std::numeric_limits<double>::quiet_NaN();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())
#endif
    for (long long h = 0; h < (long long)H; h++) {
      std::vector<T*> sub(k);
      for (int j = 0; j < k; j++) sub[j] = &d[subsets[h * k + j]];
      std::vector<double> p;
      e->estimate(sub, p);
      uint32_t c = 0;
      if (!p.empty()) for (size_t m = 0; m < n; m++) if (e->agree(p, d[m])) c++;
      if (counts) counts[h] = c;
      if (params_out) for (int j = 0; j < P; j++) params_out[h * P + j] = p.empty() ? nan : p[j];
    }
    return 0;
  }
};
struct RansacOp {
  const double* data; size_t n; double prob; bool exhaustive; double* params; uint8_t* mask; double* fraction;
  template <class T> int operator()(ParametersEstimator<T, double>* e, T*) {
    std::vector<T> d; unpack(data, n, d);
    std::vector<double> p; std::vector<bool> cs;
    double f = exhaustive ? RANSAC<T, double>::compute(p, e, d, &cs) : RANSAC<T, double>::compute(p, e, d, prob, &cs);
    *fraction = f;
    for (size_t i = 0; i < p.size(); i++) params[i] = p[i];
    if (mask) for (size_t i = 0; i < n; i++) mask[i] = (i < cs.size() && cs[i]) ? 1 : 0;
    return (int)p.size();
  }
};

}  // namespace

extern "C" {

int ref_model_info(int model, int* D, int* P, int* k) {
  static const int tab[15][3] = {{3, 6, 3}, {2, 4, 2}, {2, 4, 2}, {3, 6, 2}, {2, 3, 3}, {3, 4, 4}, {6, 7, 3}, {6, 3, 2}, {12, 6, 3}, {6, 5, 5}, {7, 6, 6}, {14, 20, 4}, {17, 17, 3}, {4, 5, 5}, {4, 8, 4}};
  if (model < 0 || model >= M_COUNT) return -1;
  if (model >= M_PLANE2 && model <= M_PLANE8) { const int d = model == M_PLANE2 ? 2 : model - M_PLANE5 + 5; *D = d; *P = 2 * d; *k = d; return 0; }
  if (model >= M_SPHERE5 && model <= M_SPHERE8) { const int d = model - M_SPHERE5 + 5; *D = d; *P = d + 1; *k = d + 1; return 0; }
  if (model >= M_LINE4 && model <= M_LINE8) { const int d = model - M_LINE4 + 4; *D = d; *P = 2 * d; *k = 2; return 0; }
  if (model >= M_DENSE2 && model <= M_DENSE8) { const int n = model <= M_DENSE4 ? model - M_DENSE2 + 2 : model - M_DENSE7 + 7; *D = n + 1; *P = n; *k = n; return 0; }
  *D = tab[model][0]; *P = tab[model][1]; *k = tab[model][2];
  return 0;
}

int ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// estimate() on n data (n >= k); returns number of parameters written (0 = degenerate).
int ref_estimate(int model, double delta, double aux, const double* data, size_t n, double* params) {
  Cfg c = {model, delta, aux, 1}; EstimateOp op = {data, n, params}; return dispatch(c, op);
}
int ref_least_squares(int model, double delta, double aux, int ls_type, const double* data, size_t n, double* params) {
  Cfg c = {model, delta, aux, ls_type}; LsqOp op = {data, n, params}; return dispatch(c, op);
}
// agree() of one parameter vector against n data; returns the inlier count, fills out[n] if non-null.
int ref_agree(int model, double delta, double aux, const double* params, int np, const double* data, size_t n, uint8_t* out) {
  Cfg c = {model, delta, aux, 1}; AgreeOp op = {params, np, data, n, out}; return dispatch(c, op);
}
int ref_score_subsets(int model, double delta, double aux, const double* data, size_t n, const int32_t* subsets, size_t H,
                      uint32_t* counts, double* params_out, int nthreads) {
  int D, P, k; if (ref_model_info(model, &D, &P, &k)) return -1;
  Cfg c = {model, delta, aux, 1}; ScoreOp op = {data, n, subsets, H, k, P, counts, params_out, nthreads}; return dispatch(c, op);
}
// AbsoluteOrientationParametersEstimator::weightedLeastSquaresEstimate (AbsoluteOrientationParametersEstimator.cxx:208-297)
int ref_weighted_absor(const double* data, size_t n, const double* weights, double* params) {
  std::vector<PointPair> d; unpack(data, n, d);
  std::vector<double> w(weights, weights + n), p;
  AbsoluteOrientationParametersEstimator e(1.0);
  std::vector<PointPair*> ptrs(d.size());
  for (size_t i = 0; i < d.size(); i++) ptrs[i] = &d[i];
  e.weightedLeastSquaresEstimate(ptrs, w, p);   // the reference has the pointer overload only (.h:86-89)
  for (size_t i = 0; i < p.size(); i++) params[i] = p[i];
  return (int)p.size();
}
// The reference's own RANSAC<T,S>::compute: exhaustive (RANSAC.hxx:150-192) or randomized (:4-145).
int ref_ransac(int model, double delta, double aux, int ls_type, const double* data, size_t n, int exhaustive, double prob,
               double* params, uint8_t* mask, double* fraction) {
  Cfg c = {model, delta, aux, ls_type}; RansacOp op = {data, n, prob, exhaustive != 0, params, mask, fraction}; return dispatch(c, op);
}

}  // extern "C"
