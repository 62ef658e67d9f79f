// Drop-in counterpart of parametersEstimators/LineParametersEstimator.{h,hxx} (re-authored).
// Line in R^d through a with unit direction n, parameters [n, a]; estimate :23-48 (rejects point
// pairs closer than delta), agree :135-150, covariance + largest eigenvector least squares :68-111.
// GPU path for d = 2..8; higher dimensions compile and report "no GPU path" at run time (b200Describe returns false).
#ifndef LSQR_B200_LINE_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_LINE_PARAMETERS_ESTIMATOR_H
#include "ParametersEstimator.h"
#include "Point.h"

namespace lsqrRecipes {

template <unsigned int dimension>
class LineParametersEstimator : public B200Estimator<Point<double, dimension> > {
  static_assert(dimension >= 2, "a line needs at least two dimensions");

 public:
  LineParametersEstimator(double delta) : B200Estimator<Point<double, dimension> >(2), deltaSquared(delta * delta), delta_(delta) {}
  void setDelta(double delta) { deltaSquared = delta * delta; delta_ = delta; }
  virtual bool b200Describe(B200EstimatorDesc& d) const { const int m = lsqr_model_line(dimension); if (m < 0) return false; d.model = m; d.delta = delta_; return true; }

 private:
  double deltaSquared;
  double delta_;
};

}  // namespace lsqrRecipes
#endif
