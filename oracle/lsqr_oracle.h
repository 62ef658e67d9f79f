/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's RANSAC hot path.
 * Never linked, imported or executed by the product (lsqrrecipes_b200/); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (1) the reference's own golden values: pivot-calibration known answers
 *       (testing/PivotCalibrationParametersEstimatorTest.cxx:47-48,82-83 on
 *       testing/Data/pivotCalibrationData.txt) and the four literal circle agree() cases
 *       (testing/SphereParametersEstimatorTest.cxx:280-296);
 *   (2) fixtures in tests/golden/ minted by the reference itself (oracle/_ref, built from
 *       /root/reference by oracle/Makefile; generator: tests/golden/make_golden.py);
 *   (3) oracle/_ref live, whenever that library is present.
 *
 * Model ids, datum layout (D doubles), parameter count P, minimal subset size k:
 *   0 PLANE3  D=3  P=6 [n,a]            k=3   PlaneParametersEstimator<3>
 *   1 LINE2D  D=2  P=4 [n,a] (normal)   k=2   Line2DParametersEstimator
 *   2 LINE2   D=2  P=4 [dir,a]          k=2   LineParametersEstimator<2>
 *   3 LINE3   D=3  P=6 [dir,a]          k=2   LineParametersEstimator<3>
 *   4 CIRCLE2 D=2  P=3 [c,r]            k=3   SphereParametersEstimator<2>
 *   5 SPHERE3 D=3  P=4 [c,r]            k=4   SphereParametersEstimator<3>
 *   6 ABSOR   D=6  P=7 [s,qx,qy,qz,t]   k=3   AbsoluteOrientationParametersEstimator
 *   7 RAY     D=6  P=3 [x,y,z]          k=2   RayIntersectionParametersEstimator (datum = p,n)
 *   8 PIVOT   D=12 P=6 [tDRF,tW]        k=3   PivotCalibrationEstimator (datum = R row-major, t)
 */
#ifndef LSQR_ORACLE_H
#define LSQR_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int orc_model_info(int model, int* D, int* P, int* k);
int orc_num_threads(void);
/* estimate(): data = n>=k data in subset order; returns #params written (0 = degenerate). */
int orc_estimate(int model, double delta, double aux, const double* data, size_t n, double* params);
int orc_least_squares(int model, double delta, double aux, int ls_type, const double* data, size_t n, double* params);
/* MINPACK info (1..4 = success) and number of function evaluations of the last Levenberg-Marquardt refit on this thread */
void orc_last_lm(int* info, int* nfev);
/* AbsoluteOrientationParametersEstimator::weightedLeastSquaresEstimate (.cxx:208-297); returns 7 or 0 */
int orc_weighted_absor(const double* data, size_t n, const double* weights, double* params);
/* agree() of one parameter vector against n data; returns inlier count; out[n] optional. */
int orc_agree(int model, double delta, double aux, const double* params, int np, const double* data, size_t n, uint8_t* out);
/* Full scoring (no early exit) of an ordered subset list: RANSAC.hxx:217-249 per subset. */
int orc_score_subsets(int model, double delta, double aux, const double* data, size_t n, const int32_t* subsets, size_t H,
                      uint32_t* counts, double* params_out, int nthreads);
/* Exhaustive driver, RANSAC.hxx:150-249: lexicographic subsets, strict '>' best update,
 * then least squares on the consensus set.  best_rank = lexicographic rank of the winner. */
int orc_ransac_exhaustive(int model, double delta, double aux, int ls_type, const double* data, size_t n,
                          double* params, uint8_t* mask, double* fraction, uint32_t* best_count, uint64_t* best_rank);
/* RANSAC.hxx:254-280 */
unsigned int orc_choose(unsigned int n, unsigned int m);
/* k-subset of {0..n-1} with the given rank in the enumeration order of RANSAC.hxx:197-213. */
void orc_unrank_lex(uint64_t rank, unsigned int n, unsigned int k, int32_t* out);
/* Stop rule RANSAC.hxx:107-110 (numerator = log(1-p)). */
unsigned int orc_num_tries(double prob, unsigned int votes, unsigned int n, unsigned int k, unsigned int all_tries);

#ifdef __cplusplus
}
#endif
#endif
