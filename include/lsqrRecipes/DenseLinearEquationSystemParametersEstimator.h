// Drop-in counterpart of parametersEstimators/DenseLinearEquationSystemParametersEstimator.{h,hxx}
// (re-authored).  Rows of a linear system A x = b as data: AugmentedRow<T,n> holds one row
// [a_0 .. a_{n-1} | b] (.h:20-135); estimate() solves n rows through the pseudo-inverse with singular
// values <= EPS dropped and fails on rank < n (.hxx:17-49); leastSquaresEstimate() does the same over
// all rows (:64-96); agree() is |a.x - b| < delta (:111-119).
//
// The engine instantiates n = 2..8 (the reference's test and example use 5 and 6,
// testing/DenseLinearEquationSystemParametersEstimatorTest.cxx:44,157, examples/linearEquationSystemSolver.cxx).
// Larger n have no GPU path: b200Describe() returns false and RANSAC::compute reports failure the
// reference's way (empty parameters, 0).
#ifndef LSQR_B200_DENSE_LINEAR_EQUATION_SYSTEM_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_DENSE_LINEAR_EQUATION_SYSTEM_PARAMETERS_ESTIMATOR_H
#include <cstring>
#include <ostream>
#include <vector>

#include "ParametersEstimator.h"

namespace lsqrRecipes {

template <class T, unsigned int n>
class AugmentedRow {
 public:
  enum { dimension = n };

  AugmentedRow() { for (unsigned int i = 0; i < n; i++) aValues[i] = T(0); bValue = T(0); }
  // fillData has at least n+1 entries: [a | b]
  AugmentedRow(T* fillData) { set(fillData); }
  AugmentedRow(T* fillData, T bData) { set(fillData, bData); }
  AugmentedRow(const AugmentedRow<T, n>& other) { *this = other; }
  AugmentedRow<T, n>& operator=(const AugmentedRow<T, n>& other) {
    for (unsigned int i = 0; i < n; i++) aValues[i] = other.aValues[i];
    bValue = other.bValue;
    return *this;
  }

  T& operator[](unsigned int index) { return index == n ? bValue : aValues[index]; }
  const T& operator[](unsigned int index) const { return index == n ? bValue : aValues[index]; }

  void set(T* fillData) { for (unsigned int i = 0; i < n; i++) aValues[i] = fillData[i]; bValue = fillData[n]; }
  void set(T* fillData, T bData) { for (unsigned int i = 0; i < n; i++) aValues[i] = fillData[i]; bValue = bData; }
  void get(T* rowData, T& bData) const { for (unsigned int i = 0; i < n; i++) rowData[i] = aValues[i]; bData = bValue; }

  friend std::ostream& operator<<(std::ostream& output, const AugmentedRow& r) {
    output << "[ ";
    for (unsigned int i = 0; i < n; i++) output << r.aValues[i] << ", ";
    output << r.bValue << " ]";
    return output;
  }

 private:
  T aValues[n];   // contiguous with bValue: the engine reads a row as n+1 consecutive values
  T bValue;
};

template <class T, unsigned int n>
class DenseLinearEquationSystemParametersEstimator : public B200Estimator<AugmentedRow<T, n> > {
 public:
  DenseLinearEquationSystemParametersEstimator(T delta) : B200Estimator<AugmentedRow<T, n> >(n) { this->delta = delta; }
  void setDelta(T delta) { this->delta = delta; }

  // Rows of a row-major (rows x n) matrix A and a vector b -> augmented rows.  (The reference's
  // getAugmentedRows takes vnl_matrix / vnl_vector; VNL is not a dependency here.)
  static void getAugmentedRows(const T* A, const T* b, unsigned int rows, std::vector<AugmentedRow<T, n> >& out) {
    out.clear();
    T tmp[n + 1];
    for (unsigned int r = 0; r < rows; r++) {
      for (unsigned int c = 0; c < n; c++) tmp[c] = A[r * n + c];
      tmp[n] = b[r];
      out.push_back(AugmentedRow<T, n>(tmp));
    }
  }

  virtual bool b200Describe(B200EstimatorDesc& d) const {
    // the engine's arithmetic is double; rows must be n+1 packed doubles
    if (sizeof(T) != sizeof(double) || sizeof(AugmentedRow<T, n>) != (n + 1) * sizeof(double)) return false;
    const int m = lsqr_model_dense(n);
    if (m < 0) return false;
    d.model = m;
    d.delta = delta;
    return true;
  }

 private:
  T delta;
};

}  // namespace lsqrRecipes
#endif
