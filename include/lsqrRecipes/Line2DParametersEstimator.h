// Drop-in counterpart of parametersEstimators/Line2DParametersEstimator.{h,cxx} (re-authored).
// 2D line in normal form, parameters [n_x, n_y, a_x, a_y]; estimate .cxx:11-32, closed-form 2x2
// eigen least squares :50-100, agree :119-123.
#ifndef LSQR_B200_LINE2D_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_LINE2D_PARAMETERS_ESTIMATOR_H
#include "ParametersEstimator.h"
#include "Point2D.h"

namespace lsqrRecipes {

class Line2DParametersEstimator : public B200Estimator<Point2D> {
 public:
  Line2DParametersEstimator(double delta) : B200Estimator<Point2D>(2), deltaSquared(delta * delta), delta_(delta) {}
  void setDelta(double delta) { deltaSquared = delta * delta; delta_ = delta; }
  virtual bool b200Describe(B200EstimatorDesc& d) const { d.model = LSQR_LINE2D; d.delta = delta_; return true; }

 private:
  double deltaSquared;
  double delta_;
};

}  // namespace lsqrRecipes
#endif
