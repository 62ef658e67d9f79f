// Drop-in counterpart of parametersEstimators/AbsoluteOrientationParametersEstimator.{h,cxx}
// (re-authored).  Rigid transformation p2 = R(q) p1 + t from 3D-3D correspondences, parameters
// [s, q_x, q_y, q_z, t_x, t_y, t_z].  Minimal solver: orthonormal-triad construction (.cxx:14-101);
// least squares: Horn's quaternion method (:120-206); agree: |T p1 - p2|^2 < delta^2 (:316-327).
// weightedLeastSquaresEstimate (:208-297) is Horn's method with one weight per pair.
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

  // Weighted least squares, AbsoluteOrientationParametersEstimator.cxx:208-297 (same signatures as the reference's .h:86-89).
  void weightedLeastSquaresEstimate(std::vector<std::pair<Point3D, Point3D>*>& data, std::vector<double>& weights, std::vector<double>& parameters) {
    parameters.clear();
    if (data.size() < this->minForEstimate) return;
    B200EstimatorDesc d;
    b200Describe(d);
    lsqr_ctx* ctx = b200::configured(d);
    if (!ctx) return;
    std::vector<double> packed;
    b200::gather(data, 6, packed);
    double prm[LSQR_MAX_PARAMS];
    int np = 0;
    weights.at(data.size() - 1);   // the reference reads weights[i] for every pair; fail loudly on a short vector
    if (!b200::check(ctx, lsqr_weighted_least_squares(ctx, packed.data(), data.size(), weights.data(), prm, &np))) return;
    parameters.assign(prm, prm + np);
  }
  void weightedLeastSquaresEstimate(std::vector<std::pair<Point3D, Point3D> >& data, std::vector<double>& weights, std::vector<double>& parameters) {
    std::vector<std::pair<Point3D, Point3D>*> ptrs(data.size());
    for (size_t i = 0; i < data.size(); i++) ptrs[i] = &data[i];
    weightedLeastSquaresEstimate(ptrs, weights, parameters);
  }

 private:
  double deltaSquared;
  double delta_;
};

}  // namespace lsqrRecipes
#endif
