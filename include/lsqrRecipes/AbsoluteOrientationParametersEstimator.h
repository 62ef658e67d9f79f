// Drop-in counterpart of parametersEstimators/AbsoluteOrientationParametersEstimator.{h,cxx}
// (re-authored).  Rigid transformation p2 = R(q) p1 + t from 3D-3D correspondences, parameters
// [s, q_x, q_y, q_z, t_x, t_y, t_z].  Minimal solver: orthonormal-triad construction (.cxx:14-101);
// least squares: Horn's quaternion method (:120-206); agree: |T p1 - p2|^2 < delta^2 (:316-327).
// The weighted variant (:208-297) is a "next" row of SURVEY.md section 8f.
#ifndef LSQR_B200_ABSOLUTE_ORIENTATION_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_ABSOLUTE_ORIENTATION_PARAMETERS_ESTIMATOR_H
#include <utility>

#include "ParametersEstimator.h"
#include "Point3D.h"

namespace lsqrRecipes {

class AbsoluteOrientationParametersEstimator : public B200Estimator<std::pair<Point3D, Point3D> > {
 public:
  AbsoluteOrientationParametersEstimator(double delta) : B200Estimator<std::pair<Point3D, Point3D> >(3), deltaSquared(delta * delta), delta_(delta) {
    static_assert(sizeof(std::pair<Point3D, Point3D>) == 48, "pair<Point3D,Point3D> must be 6 packed doubles");
  }
  void setDelta(double delta) { deltaSquared = delta * delta; delta_ = delta; }
  virtual bool b200Describe(B200EstimatorDesc& d) const { d.model = LSQR_ABSOR; d.delta = delta_; return true; }

 private:
  double deltaSquared;
  double delta_;
};

}  // namespace lsqrRecipes
#endif
