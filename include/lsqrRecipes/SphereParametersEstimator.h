// Drop-in counterpart of parametersEstimators/SphereParametersEstimator.{h,hxx} (re-authored).
// Hypersphere (p-c)^T(p-c) = r^2, parameters [c, r]; Cramer minimal solvers for the circle
// (.hxx:80-109) and the sphere (:115-163), agree :255-264 (a distance against delta), algebraic
// least squares :267-307, geometric least squares by Levenberg-Marquardt :310-338.
// GPU path for dimensions 2..8 (4 and above go through the N-D pseudo-inverse solver, :169-202); higher dimensions compile
// and report "no GPU path" at run time (b200Describe returns false).
#ifndef LSQR_B200_SPHERE_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_SPHERE_PARAMETERS_ESTIMATOR_H
#include <exception>

#include "ParametersEstimator.h"
#include "Point.h"

namespace lsqrRecipes {

template <unsigned int dimension>
class SphereParametersEstimator : public B200Estimator<Point<double, dimension> > {
  static_assert(dimension >= 2, "a hypersphere needs at least two dimensions");
  typedef Point<double, dimension> PointT;

 public:
  enum LeastSquaresType { ALGEBRAIC = 0, GEOMETRIC };

  SphereParametersEstimator(double delta, LeastSquaresType lsType = GEOMETRIC) : B200Estimator<PointT>(dimension + 1) {
    if (lsType != ALGEBRAIC && lsType != GEOMETRIC) throw std::exception();  // as SphereParametersEstimator.hxx:17-18
    this->delta = delta;
    this->lsType = lsType;
  }
  void setDelta(double delta) { this->delta = delta; }
  void setLeastSquaresType(LeastSquaresType lsType) { this->lsType = lsType; }

  void algebraicLeastSquaresEstimate(std::vector<PointT*>& data, std::vector<double>& parameters) {
    const LeastSquaresType keep = lsType;
    lsType = ALGEBRAIC;
    this->leastSquaresEstimate(data, parameters);
    lsType = keep;
  }
  // The reference starts LM from caller-supplied parameters; the GPU path always starts from its own
  // algebraic fit (which is what leastSquaresEstimate does, SphereParametersEstimator.hxx:224-230).
  void geometricLeastSquaresEstimate(std::vector<PointT*>& data, std::vector<double>& /*initialParameters*/, std::vector<double>& finalParameters) {
    const LeastSquaresType keep = lsType;
    lsType = GEOMETRIC;
    this->leastSquaresEstimate(data, finalParameters);
    lsType = keep;
  }
  virtual bool b200Describe(B200EstimatorDesc& d) const {
    const int m = lsqr_model_sphere(dimension);
    if (m < 0) return false;
    d.model = m;
    d.delta = delta;
    d.lsType = (lsType == ALGEBRAIC) ? LSQR_LS_ALGEBRAIC : LSQR_LS_GEOMETRIC;
    return true;
  }

 private:
  double delta;
  LeastSquaresType lsType;
};

}  // namespace lsqrRecipes
#endif
