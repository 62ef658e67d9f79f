// Drop-in counterpart of parametersEstimators/PlaneParametersEstimator.{h,hxx} (re-authored).
// Hyperplane dot(n, p - a) = 0, parameters [n_0..n_k, a_0..a_k] with |n| = 1.
// GPU path for dimension 3 (3-point cross-product solver, PlaneParametersEstimator.hxx:48-69) and dimension 4 (the
// generic SVD null-vector branch, :70-104); agree :196-203; covariance + smallest eigenvector least squares :129-172.
// Other dimensions fail to compile.
#ifndef LSQR_B200_PLANE_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_PLANE_PARAMETERS_ESTIMATOR_H
#include "ParametersEstimator.h"
#include "Point.h"

namespace lsqrRecipes {

template <unsigned int dimension>
class PlaneParametersEstimator : public B200Estimator<Point<double, dimension> > {
  static_assert(dimension == 3 || dimension == 4, "lsqr_b200 accelerates PlaneParametersEstimator<3> and <4>; other dimensions are not on the GPU path");

 public:
  // delta: a point is on the plane if its distance from it is less than delta
  PlaneParametersEstimator(double delta) : B200Estimator<Point<double, dimension> >(dimension), deltaSquared(delta * delta), delta_(delta) {}
  void setDelta(double delta) { deltaSquared = delta * delta; delta_ = delta; }
  virtual bool b200Describe(B200EstimatorDesc& d) const { d.model = (dimension == 3) ? LSQR_PLANE3 : LSQR_PLANE4; d.delta = delta_; return true; }

 private:
  double deltaSquared;
  double delta_;
};

}  // namespace lsqrRecipes
#endif
