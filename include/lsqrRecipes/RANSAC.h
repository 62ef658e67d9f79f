// Drop-in counterpart of parametersEstimators/RANSAC.{h,hxx} of zivy/LSQRRecipes (re-authored).
// Same namespace, class template and the same two static entry points (RANSAC.h:75-79 and
// :111-113); the bodies (RANSAC.hxx:4-249) are replaced by calls into liblsqr_b200.so:
//
//   reference (host, one thread)                 here (B200)
//   ------------------------------------------   ------------------------------------------------
//   srand/rand subset draw + std::set dedupe      Philox4x32-10 counter sampler on the device
//   estimator->estimate()   per try               one minimal solve per thread
//   N x estimator->agree()  per try               consensus kernel, R hypotheses x point tiles
//   strict '>' best update                        packed (count, ~index) arg-max
//   numTries update after each improvement        same rule, re-evaluated between rounds
//   leastSquaresEstimate(inliers)                 fused mask + moment pass, on-device eigen / LM
//
// Error convention kept: invalid input returns 0 (the randomized overload returns BEFORE clearing
// `parameters`, RANSAC.hxx:16-19 vs :43; the exhaustive overload clears first, :165-169); a
// degenerate data set leaves `parameters` empty.  Estimators without a GPU path (user-defined
// classes that do not override b200Describe) produce the same "empty parameters, return 0"
// result and b200LastError() explains; there is no CPU fallback.
#ifndef LSQR_B200_RANSAC_H
#define LSQR_B200_RANSAC_H

#include <vector>

#include "ParametersEstimator.h"

namespace lsqrRecipes {

// Process-wide knobs that the reference does not have (its random path is time-seeded).
struct B200RansacOptions {
  int precision;        // LSQR_FP32 (default) or LSQR_FP64 for the consensus scoring
  unsigned long long seed;
  // Data sets of at least this many elements are spread over every visible GPU (points replicated, hypotheses partitioned,
  // refine sharded; NCCL inside the library); smaller ones stay on one device, where a call costs fewer fixed latencies.
  size_t multiGpuMinData;
  B200RansacOptions() : precision(LSQR_FP32), seed(0), multiGpuMinData(static_cast<size_t>(1) << 20) {}
};
inline B200RansacOptions& b200Options() { static B200RansacOptions o; return o; }

template <class T, class S>
class RANSAC {
 public:
  // Randomized RANSAC; returns the fraction of the data used for the least-squares estimate.
  static double compute(std::vector<S>& parameters, ParametersEstimator<T, S>* paramEstimator, std::vector<T>& data,
                        double desiredProbabilityForNoOutliers, std::vector<bool>* consensusSet = NULL) {
    const unsigned int numDataObjects = static_cast<unsigned int>(data.size());
    if (numDataObjects < paramEstimator->numForEstimate() || desiredProbabilityForNoOutliers >= 1.0 || desiredProbabilityForNoOutliers <= 0.0) return 0;
    parameters.clear();
    return run(parameters, paramEstimator, data, false, desiredProbabilityForNoOutliers, consensusSet);
  }

  // Brute force over all C(n,k) subsets in lexicographic order; use only when that number is small.
  static double compute(std::vector<S>& parameters, ParametersEstimator<T, S>* paramEstimator, std::vector<T>& data,
                        std::vector<bool>* consensusSet = NULL) {
    parameters.clear();
    if (data.size() < paramEstimator->numForEstimate()) return 0;
    return run(parameters, paramEstimator, data, true, 0.0, consensusSet);
  }

 private:
  static double run(std::vector<S>& parameters, ParametersEstimator<T, S>* paramEstimator, std::vector<T>& data, bool exhaustive, double prob,
                    std::vector<bool>* consensusSet) {
    B200EstimatorDesc d;
    if (!paramEstimator->b200Describe(d)) {
      b200::lastErrorStorage() = "this ParametersEstimator has no GPU path (b200Describe not overridden); lsqr_b200 has no CPU fallback";
      return 0;
    }
    lsqr_ctx* ctx = b200::configured(d, data.size() >= b200Options().multiGpuMinData);
    if (!ctx) return 0;
    // the records as the engine reads them: the caller's own array when the datum's leading doubles are the engine's datum,
    // else (members separated by padding, e.g. the ultrasound-calibration records) a packed copy made here
    double probe[32];
    std::vector<double> packed;
    const void* records = data.data();
    size_t stride = sizeof(T);
    if (!data.empty() && paramEstimator->b200PackDatum(data[0], probe)) {
      int dim = 0;
      lsqr_model_info(d.model, &dim, NULL, NULL);
      packed.resize(data.size() * static_cast<size_t>(dim));
      for (size_t i = 0; i < data.size(); i++) paramEstimator->b200PackDatum(data[i], &packed[i * dim]);
      records = packed.data();
      stride = sizeof(double) * dim;
    }
    lsqr_compute_result r;
    int rc;
    if (exhaustive) {
      if (!b200::check(ctx, lsqr_upload(ctx, records, data.size(), stride))) return 0;
      rc = lsqr_ransac_exhaustive(ctx, LSQR_FP64, NULL, &r);
    } else {   // upload, first scoring round, consensus set and refine in one pipelined call
      rc = lsqr_compute(ctx, records, data.size(), stride, prob, b200Options().precision, b200Options().seed, NULL, &r);
    }
    if (!b200::check(ctx, rc)) return 0;
    if (r.best_count == 0) return 0;
    if (consensusSet) {
      // the consensus set comes back as packed bits (1/8 of a byte mask) and goes into std::vector<bool> word by word
      std::vector<uint32_t> words((data.size() + 31) / 32);
      if (!b200::check(ctx, lsqr_get_mask_bits(ctx, words.data()))) return 0;
      b200::assignBits(*consensusSet, words, data.size());
    }
    for (int i = 0; i < r.n_params; i++) parameters.push_back(static_cast<S>(r.params[i]));
    return r.fraction;
  }
};

}  // namespace lsqrRecipes
#endif
