// Drop-in counterpart of parametersEstimators/LineParametersEstimator.{h,hxx} (re-authored).
// Line in R^d through a with unit direction n, parameters [n, a]; estimate :23-48 (rejects point
// pairs closer than delta), agree :135-150, covariance + largest eigenvector least squares :68-111.
// GPU path for d = 2 and d = 3.
#ifndef LSQR_B200_LINE_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_LINE_PARAMETERS_ESTIMATOR_H
#include "ParametersEstimator.h"
#include "Point.h"

namespace lsqrRecipes {

template <unsigned int dimension>
class LineParametersEstimator : public B200Estimator<Point<double, dimension> > {
  static_assert(dimension == 2 || dimension == 3, "lsqr_b200 accelerates LineParametersEstimator<2> and <3>");

 public:
  LineParametersEstimator(double delta) : B200Estimator<Point<double, dimension> >(2), deltaSquared(delta * delta), delta_(delta) {}
  void setDelta(double delta) { deltaSquared = delta * delta; delta_ = delta; }
  virtual bool b200Describe(B200EstimatorDesc& d) const { d.model = (dimension == 2) ? LSQR_LINE2 : LSQR_LINE3; d.delta = delta_; return true; }

 private:
  double deltaSquared;
  double delta_;
};

}  // namespace lsqrRecipes
#endif
