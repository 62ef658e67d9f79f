// Drop-in counterpart of parametersEstimators/SinglePointTargetUSCalibrationParametersEstimator.{h,cxx}
// (re-authored): both single-point-target classes.  Ultrasound calibration with a single unknown point
// target: every datum is a tracked image (T2 = US reference frame -> tracker, q = pixel of the target),
// the 20 parameters are [t1, t3, omega_z, omega_y, omega_x, m_x, m_y, m_x R3(:,1), m_y R3(:,2), R3(:,3)]
// with T2 T3 (u, v, 0, 1)^T = t1.  estimate(): exactly four data, 12 x 12 pseudo-inverse + closest
// rotation + Euler angles (.cxx:17-25, 120-270); leastSquaresEstimate(): ANALYTIC (the same over all
// data) or ITERATIVE (Levenberg-Marquardt over 11 parameters from the analytic start, :272-329);
// agree(): |T2 T3 q - t1|^2 < delta^2 (:71-107).
//
// CalibratedPointerTargetUSCalibrationParametersEstimator (.cxx:663-985) is the same model with the target position
// measured by a tracked pointer (datum adds p, no unknown t1; three data, 9 x 9 system, LM over 8 parameters).
// The plane-phantom variant is not built.
#ifndef LSQR_B200_SINGLE_POINT_TARGET_US_CALIBRATION_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_SINGLE_POINT_TARGET_US_CALIBRATION_PARAMETERS_ESTIMATOR_H
#include "Frame.h"
#include "ParametersEstimator.h"
#include "Point2D.h"
#include "Point3D.h"

namespace lsqrRecipes {

struct SingleUnknownPointTargetUSCalibrationParametersEstimatorDataType {
  Frame T2;
  Point2D q;
};

class SingleUnknownPointTargetUSCalibrationParametersEstimator
    : public B200Estimator<SingleUnknownPointTargetUSCalibrationParametersEstimatorDataType> {
 public:
  enum LeastSquaresType { ANALYTIC = 0, ITERATIVE };
  typedef SingleUnknownPointTargetUSCalibrationParametersEstimatorDataType DataType;

  SingleUnknownPointTargetUSCalibrationParametersEstimator(double delta, LeastSquaresType lsType = ITERATIVE)
      : B200Estimator<DataType>(4), delta_(delta), lsType_(lsType) {}

  void setLeastSquaresType(LeastSquaresType lsType) { lsType_ = lsType; }
  void setDelta(double delta) { delta_ = delta; }

  // exactly four data elements (.cxx:21-22), unlike the other estimators' "at least k"
  virtual void estimate(std::vector<DataType*>& data, std::vector<double>& parameters) {
    parameters.clear();
    if (data.size() != this->minForEstimate) return;
    B200Estimator<DataType>::estimate(data, parameters);
  }
  virtual void estimate(std::vector<DataType>& data, std::vector<double>& parameters) {
    std::vector<DataType*> ptrs(data.size());
    for (size_t i = 0; i < data.size(); i++) ptrs[i] = &data[i];
    estimate(ptrs, parameters);
  }

  virtual bool b200Describe(B200EstimatorDesc& d) const {
    d.model = LSQR_USXW; d.delta = delta_; d.lsType = (lsType_ == ANALYTIC) ? LSQR_LS_ALGEBRAIC : LSQR_LS_GEOMETRIC;
    return true;
  }
  // {Frame, Point2D} is not 14 consecutive doubles (Frame carries an int): [R2 row-major, t2, u, v]
  virtual bool b200PackDatum(const DataType& datum, double* out) const {
    double R[3][3], t[3];
    DataType& d = const_cast<DataType&>(datum);
    d.T2.getRotationMatrix(R);
    d.T2.getTranslation(t);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) out[3 * i + j] = R[i][j]; out[9 + i] = t[i]; }
    out[12] = d.q[0]; out[13] = d.q[1];
    return true;
  }

 private:
  double delta_;
  LeastSquaresType lsType_;
};

struct CalibratedPointerTargetUSCalibrationParametersEstimatorDataType {
  Frame T2;
  Point2D q;
  Point3D p;
};

class CalibratedPointerTargetUSCalibrationParametersEstimator
    : public B200Estimator<CalibratedPointerTargetUSCalibrationParametersEstimatorDataType> {
 public:
  enum LeastSquaresType { ANALYTIC = 0, ITERATIVE };
  typedef CalibratedPointerTargetUSCalibrationParametersEstimatorDataType DataType;

  CalibratedPointerTargetUSCalibrationParametersEstimator(double delta, LeastSquaresType lsType = ITERATIVE)
      : B200Estimator<DataType>(3), delta_(delta), lsType_(lsType) {}

  void setLeastSquaresType(LeastSquaresType lsType) { lsType_ = lsType; }
  void setDelta(double delta) { delta_ = delta; }

  // exactly three data elements (.cxx:674-675)
  virtual void estimate(std::vector<DataType*>& data, std::vector<double>& parameters) {
    parameters.clear();
    if (data.size() != this->minForEstimate) return;
    B200Estimator<DataType>::estimate(data, parameters);
  }
  virtual void estimate(std::vector<DataType>& data, std::vector<double>& parameters) {
    std::vector<DataType*> ptrs(data.size());
    for (size_t i = 0; i < data.size(); i++) ptrs[i] = &data[i];
    estimate(ptrs, parameters);
  }

  virtual bool b200Describe(B200EstimatorDesc& d) const {
    d.model = LSQR_USCP; d.delta = delta_; d.lsType = (lsType_ == ANALYTIC) ? LSQR_LS_ALGEBRAIC : LSQR_LS_GEOMETRIC;
    return true;
  }
  virtual bool b200PackDatum(const DataType& datum, double* out) const {
    double R[3][3], t[3];
    DataType& d = const_cast<DataType&>(datum);
    d.T2.getRotationMatrix(R);
    d.T2.getTranslation(t);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) out[3 * i + j] = R[i][j]; out[9 + i] = t[i]; out[14 + i] = d.p[i]; }
    out[12] = d.q[0]; out[13] = d.q[1];
    return true;
  }

 private:
  double delta_;
  LeastSquaresType lsType_;
};

}  // namespace lsqrRecipes
#endif
