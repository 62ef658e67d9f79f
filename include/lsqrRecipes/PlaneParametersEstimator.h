// Drop-in counterpart of parametersEstimators/PlaneParametersEstimator.{h,hxx} (re-authored).
// Hyperplane dot(n, p - a) = 0, parameters [n_0..n_k, a_0..a_k] with |n| = 1.
// GPU path for dimension 3 (3-point cross-product solver, PlaneParametersEstimator.hxx:48-69) and dimensions 2 and 4..8 (the
// generic SVD null-vector branch, :70-104); agree :196-203; covariance + smallest eigenvector least squares :129-172.
// Dimensions above 8 compile and report "no GPU path" at run time (b200Describe returns false).
#ifndef LSQR_B200_PLANE_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_PLANE_PARAMETERS_ESTIMATOR_H
#include "ParametersEstimator.h"
#include "Point.h"

namespace lsqrRecipes {

template <unsigned int dimension>
class PlaneParametersEstimator : public B200Estimator<Point<double, dimension> > {
  static_assert(dimension >= 2, "a hyperplane needs at least two dimensions");

 public:
  // delta: a point is on the plane if its distance from it is less than delta
  PlaneParametersEstimator(double delta) : B200Estimator<Point<double, dimension> >(dimension), deltaSquared(delta * delta), delta_(delta) {}
  void setDelta(double delta) { deltaSquared = delta * delta; delta_ = delta; }
  virtual bool b200Describe(B200EstimatorDesc& d) const { const int m = lsqr_model_plane(dimension); if (m < 0) return false; d.model = m; d.delta = delta_; return true; }

 private:
  double deltaSquared;
  double delta_;
};

}  // namespace lsqrRecipes
#endif
