// Drop-in counterpart of parametersEstimators/RayIntersectionParametersEstimator.{h,cxx}
// (re-authored).  Common intersection point [x, y, z] of a set of rays.  Minimal solver: Goldman's
// two-line intersection with parallel / behind-the-origin rejection (.cxx:23-70); least squares:
// 3x3 normal equations (:100-144); agree: closest-point parameter t >= 0 and distance^2 < delta^2
// (:164-179; ray directions are assumed to be unit vectors, as in the reference).
#ifndef LSQR_B200_RAY_INTERSECTION_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_RAY_INTERSECTION_PARAMETERS_ESTIMATOR_H
#include "ParametersEstimator.h"
#include "Point3D.h"
#include "Ray3D.h"

namespace lsqrRecipes {

class RayIntersectionParametersEstimator : public B200Estimator<Ray3D> {
 public:
  RayIntersectionParametersEstimator(double delta, double minimalAngularDeviation = 0.017453292519943295769236907684886)
      : B200Estimator<Ray3D>(2), deltaSquared(delta * delta), delta_(delta), angle(minimalAngularDeviation) {
    static_assert(sizeof(Ray3D) == 48, "Ray3D must be p[3], n[3] packed");
  }
  void setDelta(double delta) { deltaSquared = delta * delta; delta_ = delta; }
  virtual bool b200Describe(B200EstimatorDesc& d) const { d.model = LSQR_RAY; d.delta = delta_; d.aux = angle; return true; }

 protected:
  virtual bool clearsBeforeLeastSquares() const { return false; }
  virtual bool guardsLeastSquaresSize() const { return false; }

 private:
  double deltaSquared;
  double delta_;
  double angle;
};

}  // namespace lsqrRecipes
#endif
