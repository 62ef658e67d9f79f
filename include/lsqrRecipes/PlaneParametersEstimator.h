// Drop-in counterpart of parametersEstimators/PlaneParametersEstimator.{h,hxx} (re-authored).
// Hyperplane dot(n, p - a) = 0, parameters [n_0..n_k, a_0..a_k] with |n| = 1.
// Only dimension == 3 is on the GPU path (3-point cross-product solver, PlaneParametersEstimator.hxx:48-69;
// agree :196-203; covariance + smallest eigenvector least squares :129-172).  Other dimensions
// (SVD null-vector branch, :70-104) are a "next" row of SURVEY.md section 8f and fail to compile.
#ifndef LSQR_B200_PLANE_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_PLANE_PARAMETERS_ESTIMATOR_H
#include "ParametersEstimator.h"
#include "Point.h"

namespace lsqrRecipes {

template <unsigned int dimension>
class PlaneParametersEstimator : public B200Estimator<Point<double, dimension> > {
  static_assert(dimension == 3, "lsqr_b200 accelerates PlaneParametersEstimator<3>; other dimensions are not on the GPU path");

 public:
  // delta: a point is on the plane if its distance from it is less than delta
  PlaneParametersEstimator(double delta) : B200Estimator<Point<double, dimension> >(dimension), deltaSquared(delta * delta), delta_(delta) {}
  void setDelta(double delta) { deltaSquared = delta * delta; delta_ = delta; }
  virtual bool b200Describe(B200EstimatorDesc& d) const { d.model = LSQR_PLANE3; d.delta = delta_; return true; }

 private:
  double deltaSquared;
  double delta_;
};

}  // namespace lsqrRecipes
#endif
