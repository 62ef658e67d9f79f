// Drop-in counterpart of parametersEstimators/PivotCalibrationParametersEstimator.{h,cxx}
// (re-authored).  Pivot calibration: translations [t_DRF, t_W] such that R_i t_DRF + t_i = t_W for
// every tracked pose.  Minimal solver: 9x6 pseudo-inverse (.cxx:9-51); least squares: 3n x 6
// (:63-96); agree: |R t_DRF + t - t_W| < delta (:108-123).
#ifndef LSQR_B200_PIVOT_CALIBRATION_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_PIVOT_CALIBRATION_PARAMETERS_ESTIMATOR_H
#include "Frame.h"
#include "ParametersEstimator.h"

namespace lsqrRecipes {

class PivotCalibrationEstimator : public B200Estimator<Frame> {
 public:
  PivotCalibrationEstimator(double delta) : B200Estimator<Frame>(3) { this->delta = delta; }
  void setDelta(double delta) { this->delta = delta; }
  virtual bool b200Describe(B200EstimatorDesc& d) const { d.model = LSQR_PIVOT; d.delta = delta; return true; }

 private:
  double delta;
};

}  // namespace lsqrRecipes
#endif
