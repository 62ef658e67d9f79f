// Drop-in counterpart of parametersEstimators/ParametersEstimator.h:26-64 of zivy/LSQRRecipes
// (re-authored).  The public interface -- estimate x2, leastSquaresEstimate x2, agree,
// numForEstimate, protected minForEstimate -- is the reference's, so user code and user-defined
// estimators compile unchanged.
//
// One addition, invisible to callers of the reference API: b200Describe().  The reference keeps
// every estimator's threshold in a private member with no getter (e.g.
// PlaneParametersEstimator.h:88), so the GPU driver in RANSAC.h asks the estimator to describe
// itself instead.  Estimators that do not override it (user-defined classes) have no GPU path;
// RANSAC::compute then reports failure through the reference's own convention (empty parameter
// vector, return value 0) and lsqrRecipes::b200LastError() says why.  There is no CPU fallback.
#ifndef LSQR_B200_PARAMETERS_ESTIMATOR_H
#define LSQR_B200_PARAMETERS_ESTIMATOR_H

#include <stdint.h>
#include <string>
#include <vector>

#include "../lsqr_b200.h"

namespace lsqrRecipes {

struct B200EstimatorDesc {
  int model;      // lsqr_model
  double delta;   // the estimator's distance threshold
  double aux;     // RayIntersectionParametersEstimator: minimalAngularDeviation, else 0
  int lsType;     // lsqr_ls_type
  B200EstimatorDesc() : model(-1), delta(0), aux(0), lsType(LSQR_LS_GEOMETRIC) {}
};

namespace b200 {

inline std::string& lastErrorStorage() { static thread_local std::string s; return s; }

// One engine context per host thread (the C ABI is "one context per host thread").
inline lsqr_ctx* context() {
  struct Holder {
    lsqr_ctx* ctx;
    Holder() : ctx(NULL) {
      if (lsqr_ctx_create(&ctx, 0) != LSQR_OK) { ctx = NULL; lastErrorStorage() = "lsqr_ctx_create failed: no usable sm_100 CUDA device (there is no CPU fallback)"; }
    }
    ~Holder() { if (ctx) lsqr_ctx_destroy(ctx); }
  };
  static thread_local Holder h;
  return h.ctx;
}

// ... and one context spanning every visible GPU (lsqr_ctx_create_multi: data replicated, hypotheses partitioned, native
// NCCL for the two exchange steps), created on first use; NULL when the process sees a single device.
inline lsqr_ctx* multiContext() {
  struct Holder {
    lsqr_ctx* ctx;
    Holder() : ctx(NULL) {
      if (lsqr_device_count() > 1 && lsqr_ctx_create_multi(&ctx, 0) != LSQR_OK) ctx = NULL;
    }
    ~Holder() { if (ctx) lsqr_ctx_destroy(ctx); }
  };
  static thread_local Holder h;
  return h.ctx;
}

inline bool check(lsqr_ctx* ctx, int rc) {
  if (rc == LSQR_OK) return true;
  lastErrorStorage() = lsqr_last_error(ctx);
  return false;
}

inline lsqr_ctx* configured(const B200EstimatorDesc& d, bool allGpus = false) {
  lsqr_ctx* ctx = allGpus ? multiContext() : NULL;
  if (!ctx) ctx = context();
  if (!ctx) return NULL;
  if (!check(ctx, lsqr_set_estimator(ctx, d.model, d.delta, d.aux, d.lsType))) return NULL;
  return ctx;
}

// Packs the leading `dim` doubles of each pointed-to record (the reference passes
// std::vector<T*> to estimate() / leastSquaresEstimate()).
template <class T>
inline void gather(const std::vector<T*>& data, int dim, std::vector<double>& out) {
  out.resize(data.size() * static_cast<size_t>(dim));
  for (size_t i = 0; i < data.size(); i++) {
    const double* rec = reinterpret_cast<const double*>(data[i]);
    for (int j = 0; j < dim; j++) out[i * dim + j] = rec[j];
  }
}

// std::vector<bool> from packed bits (datum i = bit i & 31 of word i >> 5).  libstdc++ stores vector<bool> as 64-bit words
// with the same bit order, reachable through the iterator's word pointer: one memcpy instead of n bit insertions (10 M bits:
// ~0.1 ms against ~15 ms).  Any other standard library takes the portable loop.
inline void assignBits(std::vector<bool>& out, const std::vector<uint32_t>& words, size_t n) {
  out.assign(n, false);
  if (n == 0) return;
#if defined(__GLIBCXX__) && defined(__BYTE_ORDER__) && (__BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__)
  unsigned char* dst = reinterpret_cast<unsigned char*>(out.begin()._M_p);
  const unsigned char* src = reinterpret_cast<const unsigned char*>(words.data());
  const size_t full = n / 8;
  for (size_t i = 0; i < full; i++) dst[i] = src[i];
  for (size_t i = full * 8; i < n; i++) out[i] = ((words[i >> 5] >> (i & 31)) & 1u) != 0;
#else
  for (size_t i = 0; i < n; i++) out[i] = ((words[i >> 5] >> (i & 31)) & 1u) != 0;
#endif
}

}  // namespace b200

// Text of the last failure on this thread (CUDA errors, unsupported estimator, ...).
inline const char* b200LastError() { return b200::lastErrorStorage().c_str(); }

template <class T, class S>
class ParametersEstimator {
 public:
  // minElements: number of data objects needed for an exact estimate
  ParametersEstimator(unsigned int minElements) { this->minForEstimate = minElements; }
  virtual ~ParametersEstimator() {}

  // exact estimate from the minimal number of data objects; `parameters` is cleared, then filled
  virtual void estimate(std::vector<T*>& data, std::vector<S>& parameters) = 0;
  virtual void estimate(std::vector<T>& data, std::vector<S>& parameters) = 0;

  // least-squares estimate from an over-determined data set; `parameters` is cleared, then filled
  virtual void leastSquaresEstimate(std::vector<T*>& data, std::vector<S>& parameters) = 0;
  virtual void leastSquaresEstimate(std::vector<T>& data, std::vector<S>& parameters) = 0;

  // does the datum agree with the model?
  virtual bool agree(std::vector<S>& parameters, T& data) = 0;

  unsigned int numForEstimate() { return this->minForEstimate; }

  // GPU path hook (see the header comment).  false = this estimator has no GPU path.
  virtual bool b200Describe(B200EstimatorDesc& /*desc*/) const { return false; }
  // Data types whose leading doubles are NOT the engine's datum (members separated by padding, e.g. the
  // ultrasound-calibration records {Frame T2; Point2D q;}) write the packed datum to `out` and return true.
  virtual bool b200PackDatum(const T& /*datum*/, double* /*out*/) const { return false; }

 protected:
  unsigned int minForEstimate;
};

// Shared implementation of the five virtuals for the estimators that live on the GPU path.
// T's leading doubles are the datum (see lsqr_b200.h); everything is forwarded to the C ABI.
template <class T>
class B200Estimator : public ParametersEstimator<T, double> {
 public:
  B200Estimator(unsigned int minElements) : ParametersEstimator<T, double>(minElements) {}

  virtual void estimate(std::vector<T*>& data, std::vector<double>& parameters) {
    parameters.clear();
    if (this->minForEstimate == 0 || data.size() < this->minForEstimate) return;
    std::vector<double> packed;
    runEstimate(pack(data, packed), data.size(), parameters);
  }
  virtual void estimate(std::vector<T>& data, std::vector<double>& parameters) {
    std::vector<T*> ptrs(data.size());
    for (size_t i = 0; i < data.size(); i++) ptrs[i] = &data[i];
    estimate(ptrs, parameters);
  }
  virtual void leastSquaresEstimate(std::vector<T*>& data, std::vector<double>& parameters) {
    if (clearsBeforeLeastSquares()) parameters.clear();
    if (guardsLeastSquaresSize() && data.size() < this->minForEstimate) return;
    std::vector<double> packed;
    runLeastSquares(pack(data, packed), data.size(), parameters);
  }
  virtual void leastSquaresEstimate(std::vector<T>& data, std::vector<double>& parameters) {
    std::vector<T*> ptrs(data.size());
    for (size_t i = 0; i < data.size(); i++) ptrs[i] = &data[i];
    leastSquaresEstimate(ptrs, parameters);
  }
  virtual bool agree(std::vector<double>& parameters, T& data) {
    B200EstimatorDesc d;
    this->b200Describe(d);
    lsqr_ctx* ctx = b200::configured(d);
    if (!ctx) return false;
    int dim = 0, np = 0, k = 0;
    lsqr_model_info(d.model, &dim, &np, &k);
    parameters.at(np - 1);  // the reference indexes past the end of a short vector; fail loudly instead
    uint8_t out = 0;
    double packed[32];
    const double* datum = this->b200PackDatum(data, packed) ? packed : reinterpret_cast<const double*>(&data);
    if (!b200::check(ctx, lsqr_agree(ctx, parameters.data(), datum, 1, &out))) return false;
    return out != 0;
  }

 protected:
  // RayIntersectionParametersEstimator::leastSquaresEstimate neither clears `parameters` nor
  // checks the data size (RayIntersectionParametersEstimator.cxx:100-144); everyone else does both.
  virtual bool clearsBeforeLeastSquares() const { return true; }
  virtual bool guardsLeastSquaresSize() const { return true; }

 private:
  const double* pack(std::vector<T*>& data, std::vector<double>& packed) const {
    B200EstimatorDesc d;
    this->b200Describe(d);
    int dim = 0;
    lsqr_model_info(d.model, &dim, NULL, NULL);
    b200::gather(data, dim, packed);
    for (size_t i = 0; i < data.size(); i++) if (!this->b200PackDatum(*data[i], &packed[i * dim])) break;   // custom layouts overwrite the raw copy
    return packed.data();
  }
  void runEstimate(const double* packed, size_t n, std::vector<double>& parameters) const {
    B200EstimatorDesc d;
    this->b200Describe(d);
    lsqr_ctx* ctx = b200::configured(d);
    if (!ctx) return;
    double prm[LSQR_MAX_PARAMS];
    int np = 0;
    if (!b200::check(ctx, lsqr_estimate(ctx, packed, n, prm, &np))) return;
    parameters.insert(parameters.end(), prm, prm + np);
  }
  void runLeastSquares(const double* packed, size_t n, std::vector<double>& parameters) const {
    B200EstimatorDesc d;
    this->b200Describe(d);
    lsqr_ctx* ctx = b200::configured(d);
    if (!ctx) return;
    double prm[LSQR_MAX_PARAMS];
    int np = 0;
    if (!b200::check(ctx, lsqr_least_squares(ctx, packed, n, prm, &np))) return;
    parameters.insert(parameters.end(), prm, prm + np);
  }
};

}  // namespace lsqrRecipes
#endif
