/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's RANSAC hot path.
 * See lsqr_oracle.h for the role, the parity status (PINNED) and the model table.
 * Every function cites the reference file:line (under /root/reference) it follows.
 * Compile with: gcc -std=c99 -O2 -ffp-contract=off (no FMA contraction, x86-64 baseline),
 * which makes every plain-double expression below round exactly like the reference
 * compiled with the same flags.
 *
 * The reference delegates eigen / SVD / Levenberg-Marquardt to VNL (third party, absent
 * from /root/reference, un-pinned).  Those three routines are restated here from their
 * published contracts (ascending symmetric eigenvalues with column eigenvectors;
 * pseudo-inverse with singular values <= EPS zeroed; Levenberg-Marquardt = MINPACK's lmder, restated in
 * minpack_lm.h), so results that pass through them match to rounding level, not bit-for-bit.
 */
#include "lsqr_oracle.h"
#include "minpack_lm.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { M_PLANE3 = 0, M_LINE2D = 1, M_LINE2 = 2, M_LINE3 = 3, M_CIRCLE2 = 4, M_SPHERE3 = 5, M_ABSOR = 6, M_RAY = 7, M_PIVOT = 8, M_DENSE5 = 9, M_DENSE6 = 10, M_USXW = 11, M_USCP = 12, M_SPHERE4 = 13, M_PLANE4 = 14,
       /* the rest of the reference's template space: PlaneParametersEstimator<2, 5..8>, SphereParametersEstimator<5..8>,
        * LineParametersEstimator<4..8>, DenseLinearEquationSystemParametersEstimator<double, 2..4, 7, 8> */
       M_PLANE2 = 15, M_PLANE5 = 16, M_PLANE8 = 19, M_SPHERE5 = 20, M_SPHERE8 = 23, M_LINE4 = 24, M_LINE8 = 28, M_DENSE2 = 29, M_DENSE4 = 31, M_DENSE7 = 32, M_DENSE8 = 33, M_COUNT = 34 };
#define MAXD 9   /* doubles per datum of the dimension-templated estimators (8-unknown dense row) */

/* Family and dimension of a dimension-templated estimator: every branch below that depends on `dimension` in the reference
 * takes it from here.  F_NONE: one of the fixed-size estimators. */
enum { F_NONE = 0, F_PLANE_ND, F_SPHERE_ND, F_LINE, F_DENSE };
static int family(int model, int* dim) {
  int d = 0, f = F_NONE;
  if (model == M_PLANE4) { f = F_PLANE_ND; d = 4; }
  else if (model == M_PLANE2) { f = F_PLANE_ND; d = 2; }
  else if (model >= M_PLANE5 && model <= M_PLANE8) { f = F_PLANE_ND; d = model - M_PLANE5 + 5; }
  else if (model == M_SPHERE4) { f = F_SPHERE_ND; d = 4; }
  else if (model >= M_SPHERE5 && model <= M_SPHERE8) { f = F_SPHERE_ND; d = model - M_SPHERE5 + 5; }
  else if (model == M_LINE2) { f = F_LINE; d = 2; }
  else if (model == M_LINE3) { f = F_LINE; d = 3; }
  else if (model >= M_LINE4 && model <= M_LINE8) { f = F_LINE; d = model - M_LINE4 + 4; }
  else if (model == M_DENSE5) { f = F_DENSE; d = 5; }
  else if (model == M_DENSE6) { f = F_DENSE; d = 6; }
  else if (model >= M_DENSE2 && model <= M_DENSE4) { f = F_DENSE; d = model - M_DENSE2 + 2; }
  else if (model == M_DENSE7 || model == M_DENSE8) { f = F_DENSE; d = model - M_DENSE7 + 7; }
  if (dim) *dim = d;
  return f;
}

/* common/Epsilon.h:19 */
static const double EPS = 2.220446049250313e-016;
/* SphereParametersEstimator.hxx:11 */
static const double SPHERE_EPS = 1e-9;
/* common/Frame.cxx:8-10 */
static const double FRAME_SMALL_ANGLE = 0.008726535498373935;
static const double FRAME_HALF_PI = 3.14159265358979323846 / 2.0;

int orc_model_info(int model, int* D, int* P, int* k) {
  static const int tab[15][3] = {{3, 6, 3}, {2, 4, 2}, {2, 4, 2}, {3, 6, 2}, {2, 3, 3}, {3, 4, 4}, {6, 7, 3}, {6, 3, 2}, {12, 6, 3}, {6, 5, 5}, {7, 6, 6}, {14, 20, 4}, {17, 17, 3}, {4, 5, 5}, {4, 8, 4}};
  int d;
  if (model < 0 || model >= M_COUNT) return -1;
  switch (model > 14 ? family(model, &d) : F_NONE) {
    case F_PLANE_ND: *D = d; *P = 2 * d; *k = d; return 0;
    case F_SPHERE_ND: *D = d; *P = d + 1; *k = d + 1; return 0;
    case F_LINE: *D = d; *P = 2 * d; *k = 2; return 0;
    case F_DENSE: *D = d + 1; *P = d; *k = d; return 0;
    default: break;
  }
  *D = tab[model][0]; *P = tab[model][1]; *k = tab[model][2];
  return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------ */
/* Small dense kernels standing in for VNL                                              */
/* ------------------------------------------------------------------------------------ */

/* Cyclic Jacobi for a symmetric n x n matrix (row-major a, destroyed).  On return d[] holds
 * the eigenvalues ASCENDING and the columns of v the matching unit eigenvectors -- the
 * contract of vnl_symmetric_eigensystem used at PlaneParametersEstimator.hxx:163-169,
 * LineParametersEstimator.hxx:102-108, AbsoluteOrientationParametersEstimator.cxx:192-198. */
static void sym_eig(int n, double* a, double* v, double* d) {
  int i, j, k, p, q, sweep;
  for (i = 0; i < n; i++) for (j = 0; j < n; j++) v[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (sweep = 0; sweep < 100; sweep++) {
    double off = 0.0;
    for (p = 0; p < n; p++) for (q = p + 1; q < n; q++) off += a[p * n + q] * a[p * n + q];
    if (off == 0.0) break;
    for (p = 0; p < n; p++) {
      for (q = p + 1; q < n; q++) {
        double apq = a[p * n + q], theta, t, c, s;
        if (apq == 0.0) continue;
        theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
        t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        c = 1.0 / sqrt(t * t + 1.0);
        s = t * c;
        for (k = 0; k < n; k++) { double x = a[k * n + p], y = a[k * n + q]; a[k * n + p] = c * x - s * y; a[k * n + q] = s * x + c * y; }
        for (k = 0; k < n; k++) { double x = a[p * n + k], y = a[q * n + k]; a[p * n + k] = c * x - s * y; a[q * n + k] = s * x + c * y; }
        for (k = 0; k < n; k++) { double x = v[k * n + p], y = v[k * n + q]; v[k * n + p] = c * x - s * y; v[k * n + q] = s * x + c * y; }
      }
    }
  }
  for (i = 0; i < n; i++) d[i] = a[i * n + i];
  for (i = 0; i + 1 < n; i++) {
    int m = i;
    for (j = i + 1; j < n; j++) if (d[j] < d[m]) m = j;
    if (m != i) {
      double t = d[i]; d[i] = d[m]; d[m] = t;
      for (k = 0; k < n; k++) { t = v[k * n + i]; v[k * n + i] = v[k * n + m]; v[k * n + m] = t; }
    }
  }
}

/* Least-squares / pseudo-inverse solve x = pinv(A) b for A (m x n, m >= n, row-major), with
 * singular values <= tol zeroed; returns the rank.  One-sided Jacobi SVD.  Contract of
 * vnl_matrix_inverse + zero_out_absolute(EPS) + rank() + operator* used at
 * SphereParametersEstimator.hxx:288-294, RayIntersectionParametersEstimator.cxx:132-139,
 * PivotCalibrationParametersEstimator.cxx:40-47,85-92. */
static int pinv_solve(int m, int n, const double* Ain, const double* b, double tol, double* x) {
  double* A = (double*)malloc(sizeof(double) * (size_t)m * n);
  double V[12 * 12], w[12], y[12];
  int i, j, k, p, q, sweep, rank = 0;
  memcpy(A, Ain, sizeof(double) * (size_t)m * n);
  for (i = 0; i < n; i++) for (j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (sweep = 0; sweep < 60; sweep++) {
    int rotated = 0;
    for (p = 0; p < n; p++) {
      for (q = p + 1; q < n; q++) {
        double alpha = 0, beta = 0, gamma = 0, zeta, t, c, s;
        for (k = 0; k < m; k++) { double ap = A[(size_t)k * n + p], aq = A[(size_t)k * n + q]; alpha += ap * ap; beta += aq * aq; gamma += ap * aq; }
        if (gamma == 0.0 || fabs(gamma) <= 1e-16 * sqrt(alpha * beta)) continue;
        rotated = 1;
        zeta = (beta - alpha) / (2.0 * gamma);
        t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        c = 1.0 / sqrt(1.0 + t * t);
        s = c * t;
        for (k = 0; k < m; k++) { double xk = A[(size_t)k * n + p], yk = A[(size_t)k * n + q]; A[(size_t)k * n + p] = c * xk - s * yk; A[(size_t)k * n + q] = s * xk + c * yk; }
        for (k = 0; k < n; k++) { double xk = V[k * n + p], yk = V[k * n + q]; V[k * n + p] = c * xk - s * yk; V[k * n + q] = s * xk + c * yk; }
      }
    }
    if (!rotated) break;
  }
  /* A now holds U*diag(w); x = V diag(1/w) U^T b = sum_j V[:,j] (A[:,j].b) / w_j^2 */
  for (j = 0; j < n; j++) {
    double s2 = 0, ub = 0;
    for (k = 0; k < m; k++) { double a = A[(size_t)k * n + j]; s2 += a * a; ub += a * b[k]; }
    w[j] = sqrt(s2);
    if (w[j] <= tol) y[j] = 0.0; else { y[j] = ub / s2; rank++; }
  }
  for (i = 0; i < n; i++) { double s = 0; for (j = 0; j < n; j++) s += V[i * n + j] * y[j]; x[i] = s; }
  free(A);
  return rank;
}

/* Null vector of A (m x n, m < n, row-major) and the number of singular values above tol: the contract of
 * vnl_svd + zero_out_absolute + rank() + nullvector() at PlaneParametersEstimator.hxx:82-90.  vnl_svd carries n
 * singular values (descending, the trailing n-m zero) and an n x n V; nullvector() is V's last column.  Same
 * one-sided Jacobi as pinv_solve, run on the columns of the wide matrix. */
static int null_vector(int m, int n, const double* Ain, double tol, double* x) {
  double A[MAXD * MAXD], V[MAXD * MAXD], w[MAXD];
  int i, j, k, p, q, sweep, rank = 0, last = 0;
  memcpy(A, Ain, sizeof(double) * (size_t)m * n);
  for (i = 0; i < n; i++) for (j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (sweep = 0; sweep < 60; sweep++) {
    int rotated = 0;
    for (p = 0; p < n; p++) {
      for (q = p + 1; q < n; q++) {
        double alpha = 0, beta = 0, gamma = 0, zeta, t, c, s;
        for (k = 0; k < m; k++) { double ap = A[k * n + p], aq = A[k * n + q]; alpha += ap * ap; beta += aq * aq; gamma += ap * aq; }
        if (gamma == 0.0 || fabs(gamma) <= 1e-16 * sqrt(alpha * beta)) continue;
        rotated = 1;
        zeta = (beta - alpha) / (2.0 * gamma);
        t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        c = 1.0 / sqrt(1.0 + t * t);
        s = c * t;
        for (k = 0; k < m; k++) { double xk = A[k * n + p], yk = A[k * n + q]; A[k * n + p] = c * xk - s * yk; A[k * n + q] = s * xk + c * yk; }
        for (k = 0; k < n; k++) { double xk = V[k * n + p], yk = V[k * n + q]; V[k * n + p] = c * xk - s * yk; V[k * n + q] = s * xk + c * yk; }
      }
    }
    if (!rotated) break;
  }
  for (j = 0; j < n; j++) {
    double s2 = 0;
    for (k = 0; k < m; k++) s2 += A[k * n + j] * A[k * n + j];
    w[j] = sqrt(s2);
    if (w[j] > tol) rank++;
    if (w[j] <= w[last]) last = j;   /* the column a stable descending sort puts last */
  }
  for (i = 0; i < n; i++) x[i] = V[i * n + last];
  return rank;
}

/* ------------------------------------------------------------------------------------ */
/* Frame helpers (common/Frame.cxx)                                                     */
/* ------------------------------------------------------------------------------------ */

/* Frame.cxx:174-199 with normalizeQuaternion=false (the default, Frame.h:127) /
 * Frame.cxx:750-772 with the optional normalisation. */
static void quat_to_matrix(double s, double qx, double qy, double qz, int normalize, double R[9]) {
  if (normalize) {
    double norm = sqrt(s * s + qx * qx + qy * qy + qz * qz);
    s /= norm; qx /= norm; qy /= norm; qz /= norm;
  }
  R[0] = 1 - 2 * (qy * qy + qz * qz);
  R[1] = 2 * (qx * qy - s * qz);
  R[2] = 2 * (qx * qz + s * qy);
  R[3] = 2 * (qx * qy + s * qz);
  R[4] = 1 - 2 * (qx * qx + qz * qz);
  R[5] = 2 * (qy * qz - s * qx);
  R[6] = 2 * (qx * qz - s * qy);
  R[7] = 2 * (qy * qz + s * qx);
  R[8] = 1 - 2 * (qx * qx + qy * qy);
}

/* Frame.cxx:208-248 */
static void frame_apply(const double R[9], const double t[3], const double p[3], double out[3]) {
  double x = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + t[0];
  double y = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + t[1];
  double z = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + t[2];
  out[0] = x; out[1] = y; out[2] = z;
}

/* Frame.cxx:952-988 */
static void matrix_to_quat(const double R[9], double q[4]) {
  double startSingularRange = FRAME_HALF_PI - FRAME_SMALL_ANGLE;
  double endSingularRange = FRAME_HALF_PI + FRAME_SMALL_ANGLE;
  double halfTheta;
  q[0] = (0.5 * sqrt(R[0] + R[4] + R[8] + 1));
  halfTheta = acos(q[0]);
  if (!(halfTheta > startSingularRange && halfTheta < endSingularRange)) {
    double denom = 4 * q[0];
    q[1] = (R[7] - R[5]) / denom;
    q[2] = (R[2] - R[6]) / denom;
    q[3] = (R[3] - R[1]) / denom;
  } else {
    int i = 0, j, k;
    double w;
    if (R[4] > R[i * 3 + i]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    j = (i + 1) % 3;
    k = (j + 1) % 3;
    w = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1);
    q[i + 1] = w / 2.0;
    q[j + 1] = (R[i * 3 + j] + R[j * 3 + i]) / (2 * w);
    q[k + 1] = (R[i * 3 + k] + R[k * 3 + i]) / (2 * w);
  }
}

/* ------------------------------------------------------------------------------------ */
/* estimate(): minimal solvers                                                          */
/* ------------------------------------------------------------------------------------ */

/* PlaneParametersEstimator.hxx:36-109, dimension==3 branch :48-69, point :107-108 */
static int plane3_estimate(const double* d, double* prm) {
  const double *p0 = d, *p1 = d + 3, *p2 = d + 6;
  double v1[3], v2[3], nx, ny, nz, norm;
  v1[0] = p1[0] - p0[0]; v1[1] = p1[1] - p0[1]; v1[2] = p1[2] - p0[2];
  v2[0] = p2[0] - p0[0]; v2[1] = p2[1] - p0[1]; v2[2] = p2[2] - p0[2];
  nx = v1[1] * v2[2] - v1[2] * v2[1];
  ny = v1[2] * v2[0] - v1[0] * v2[2];
  nz = v1[0] * v2[1] - v1[1] * v2[0];
  norm = sqrt(nx * nx + ny * ny + nz * nz);
  if (norm < EPS) return 0;
  prm[0] = nx / norm; prm[1] = ny / norm; prm[2] = nz / norm;
  prm[3] = p0[0]; prm[4] = p0[1]; prm[5] = p0[2];
  return 6;
}

/* Line2DParametersEstimator.cxx:11-32 */
static int line2d_estimate(const double* d, double deltaSquared, double* prm) {
  const double *p0 = d, *p1 = d + 2;
  double nx = p1[1] - p0[1];
  double ny = p0[0] - p1[0];
  double normSquared = nx * nx + ny * ny, norm;
  if (normSquared < deltaSquared) return 0;
  norm = sqrt(nx * nx + ny * ny);
  prm[0] = nx / norm; prm[1] = ny / norm; prm[2] = p0[0]; prm[3] = p0[1];
  return 4;
}

/* LineParametersEstimator.hxx:23-48; the separation test uses Point::distanceSquared
 * (common/Point.h:72-74) = sum (a_i-b_i)^2 accumulated left to right. */
static int line_estimate(int dim, const double* d, double deltaSquared, double* prm) {
  const double *p0 = d, *p1 = d + dim;
  double ds = 0, dirNorm = 0.0;
  int i;
  for (i = 0; i < dim; i++) { double t = p0[i] - p1[i]; ds += t * t; }
  if (ds < deltaSquared) return 0;
  for (i = 0; i < dim; i++) {
    prm[i] = p0[i] - p1[i];
    dirNorm += prm[i] * prm[i];
    prm[dim + i] = p0[i];
  }
  dirNorm = sqrt(dirNorm);
  for (i = 0; i < dim; i++) prm[i] /= dirNorm;
  return 2 * dim;
}

/* SphereParametersEstimator.hxx:80-109 (estimate2D) */
static int circle_estimate(const double* d, double* prm) {
  const double *p0 = d, *p1 = d + 2, *p2 = d + 4;
  double A00, A01, A10, A11, b0, b1, detA;
  A00 = p0[0] - p1[0]; A01 = p0[1] - p1[1];
  A10 = p0[0] - p2[0]; A11 = p0[1] - p2[1];
  detA = (A00 * A11 - A01 * A10);
  if (fabs(detA) < SPHERE_EPS) return 0;
  detA *= 2.0;
  b0 = A00 * (p0[0] + p1[0]) + A01 * (p0[1] + p1[1]);
  b1 = A10 * (p0[0] + p2[0]) + A11 * (p0[1] + p2[1]);
  prm[0] = (A11 * b0 - A01 * b1) / detA;
  prm[1] = (A00 * b1 - A10 * b0) / detA;
  prm[2] = sqrt((p0[0] - prm[0]) * (p0[0] - prm[0]) + (p0[1] - prm[1]) * (p0[1] - prm[1]));
  return 3;
}

/* SphereParametersEstimator.hxx:115-163 (estimate3D) */
static int sphere3_estimate(const double* d, double* prm) {
  const double *p0 = d, *p1 = d + 3, *p2 = d + 6, *p3 = d + 9;
  double A00, A01, A02, A10, A11, A12, A20, A21, A22;
  double CT00, CT01, CT02, CT10, CT11, CT12, CT20, CT21, CT22;
  double b0, b1, b2, detA;
  A00 = p0[0] - p1[0]; A01 = p0[1] - p1[1]; A02 = p0[2] - p1[2];
  A10 = p0[0] - p2[0]; A11 = p0[1] - p2[1]; A12 = p0[2] - p2[2];
  A20 = p0[0] - p3[0]; A21 = p0[1] - p3[1]; A22 = p0[2] - p3[2];
  CT00 = A11 * A22 - A12 * A21;
  CT10 = A12 * A20 - A10 * A22;
  CT20 = A10 * A21 - A11 * A20;
  detA = A00 * CT00 + A01 * CT10 + A02 * CT20;
  if (fabs(detA) < SPHERE_EPS) return 0;
  detA *= 2;
  CT01 = A02 * A21 - A01 * A22;
  CT11 = A00 * A22 - A02 * A20;
  CT21 = A01 * A20 - A00 * A21;
  CT02 = A01 * A12 - A02 * A11;
  CT12 = A02 * A10 - A00 * A12;
  CT22 = A00 * A11 - A01 * A10;
  b0 = A00 * (p0[0] + p1[0]) + A01 * (p0[1] + p1[1]) + A02 * (p0[2] + p1[2]);
  b1 = A10 * (p0[0] + p2[0]) + A11 * (p0[1] + p2[1]) + A12 * (p0[2] + p2[2]);
  b2 = A20 * (p0[0] + p3[0]) + A21 * (p0[1] + p3[1]) + A22 * (p0[2] + p3[2]);
  prm[0] = (CT00 * b0 + CT01 * b1 + CT02 * b2) / detA;
  prm[1] = (CT10 * b0 + CT11 * b1 + CT12 * b2) / detA;
  prm[2] = (CT20 * b0 + CT21 * b1 + CT22 * b2) / detA;
  prm[3] = sqrt(((p0[0] - prm[0]) * (p0[0] - prm[0])) + ((p0[1] - prm[1]) * (p0[1] - prm[1])) + ((p0[2] - prm[2]) * (p0[2] - prm[2])));
  return 4;
}

/* One side of the triad construction, AbsoluteOrientationParametersEstimator.cxx:24-51
 * (first set) / :53-81 (second set).  VNL helpers restated: normalize() multiplies by
 * 1/sqrt(sum x^2); dot_product and magnitude accumulate left to right. */
static int triad(const double* P0, const double* P1, const double* P2, double Rm[9], double mean[3]) {
  double x[3], y[3], z[3], s, inv, dot;
  int i;
  mean[0] = (P0[0] + P1[0] + P2[0]) / 3.0;
  mean[1] = (P0[1] + P1[1] + P2[1]) / 3.0;
  mean[2] = (P0[2] + P1[2] + P2[2]) / 3.0;
  for (i = 0; i < 3; i++) x[i] = P0[i] - mean[i];
  s = 0.0; for (i = 0; i < 3; i++) s += x[i] * x[i];
  if (s != 0.0) { inv = 1.0 / sqrt(s); for (i = 0; i < 3; i++) x[i] = inv * x[i]; }
  for (i = 0; i < 3; i++) y[i] = P1[i] - mean[i];
  dot = 0.0; for (i = 0; i < 3; i++) dot += y[i] * x[i];
  for (i = 0; i < 3; i++) y[i] = y[i] - x[i] * dot;
  s = 0.0; for (i = 0; i < 3; i++) s += y[i] * y[i];
  if (s != 0.0) { inv = 1.0 / sqrt(s); for (i = 0; i < 3; i++) y[i] = inv * y[i]; }
  z[0] = x[1] * y[2] - x[2] * y[1];
  z[1] = x[2] * y[0] - x[0] * y[2];
  z[2] = x[0] * y[1] - x[1] * y[0];
  s = 0.0; for (i = 0; i < 3; i++) s += z[i] * z[i];
  if (sqrt(s) < EPS) return 0;
  for (i = 0; i < 3; i++) { Rm[i * 3 + 0] = x[i]; Rm[i * 3 + 1] = y[i]; Rm[i * 3 + 2] = z[i]; }
  return 1;
}

/* AbsoluteOrientationParametersEstimator.cxx:14-101 */
/* PlaneParametersEstimator.hxx:70-108 (dimensions other than 3): the hyperplane normal and offset are the null space of
 * the k x (k+1) matrix [p_i, -1]; rank < k -> linearly dependent points; the normal is scaled to unit length, the point
 * on the plane is the first datum */
static int plane_nd_estimate(int dim, const double* d, double* prm) {
  double A[MAXD * MAXD], x[MAXD], norm = 0;
  int i, j;
  for (i = 0; i < dim; i++) { for (j = 0; j < dim; j++) A[i * (dim + 1) + j] = d[i * dim + j]; A[i * (dim + 1) + dim] = -1; }
  if (null_vector(dim, dim + 1, A, EPS, x) < dim) return 0;
  for (i = 0; i < dim; i++) { norm += x[i] * x[i]; prm[i] = x[i]; }
  norm = 1.0 / sqrt(norm);
  for (i = 0; i < dim; i++) prm[i] *= norm;
  for (i = 0; i < dim; i++) prm[dim + i] = d[i];
  return 2 * dim;
}

/* SphereParametersEstimator.hxx:169-202 (estimateND, dimensions other than 2 and 3): rows p0 - p_i, pseudo-inverse with
 * singular values <= EPS zeroed, rank < dim -> coplanar points */
static int sphere_nd_estimate(int dim, const double* d, double* prm) {
  double A[MAXD * MAXD], b[MAXD] = {0}, x[MAXD], rSquared = 0.0;
  int i, j, rank;
  for (i = 0; i < dim; i++)
    for (j = 0; j < dim; j++) { A[i * dim + j] = d[j] - d[(i + 1) * dim + j]; b[i] += A[i * dim + j] * (d[j] + d[(i + 1) * dim + j]); }
  rank = pinv_solve(dim, dim, A, b, EPS, x);
  if (rank < dim) return 0;
  for (i = 0; i < dim; i++) { x[i] = x[i] * 0.5; prm[i] = x[i]; rSquared += (d[i] - x[i]) * (d[i] - x[i]); }
  prm[dim] = sqrt(rSquared);
  return dim + 1;
}

static int absor_estimate(const double* d, double* prm) {
  double R1[9], R2[9], R[9], m1[3], m2[3], t[3], q[4];
  int i, j, k;
  if (!triad(d + 0, d + 6, d + 12, R1, m1)) return 0;
  if (!triad(d + 3, d + 9, d + 15, R2, m2)) return 0;
  for (i = 0; i < 3; i++)
    for (k = 0; k < 3; k++) { double s = 0.0; for (j = 0; j < 3; j++) s += R2[i * 3 + j] * R1[k * 3 + j]; R[i * 3 + k] = s; }
  for (i = 0; i < 3; i++) { double s = 0.0; for (j = 0; j < 3; j++) s += R[i * 3 + j] * m1[j]; t[i] = m2[i] - s; }
  matrix_to_quat(R, q);
  prm[0] = q[0]; prm[1] = q[1]; prm[2] = q[2]; prm[3] = q[3];
  prm[4] = t[0]; prm[5] = t[1]; prm[6] = t[2];
  return 7;
}

/* RayIntersectionParametersEstimator.cxx:23-70; crossEps = sin^2(angle), :11-16 */
static int ray_estimate(const double* d, double crossEps, double* prm) {
  const double *p1 = d, *n1 = d + 3, *p2 = d + 6, *n2 = d + 9;
  double p21[3], c[3], t1, t2, denominator;
  p21[0] = p2[0] - p1[0]; p21[1] = p2[1] - p1[1]; p21[2] = p2[2] - p1[2];
  c[0] = n1[1] * n2[2] - n1[2] * n2[1];
  c[1] = n1[2] * n2[0] - n1[0] * n2[2];
  c[2] = n1[0] * n2[1] - n1[1] * n2[0];
  denominator = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
  if (denominator < crossEps) return 0;
  t1 = (c[0] * (p21[1] * n2[2] - p21[2] * n2[1]) - c[1] * (p21[0] * n2[2] - p21[2] * n2[0]) + c[2] * (p21[0] * n2[1] - p21[1] * n2[0])) / denominator;
  t2 = (c[0] * (p21[1] * n1[2] - p21[2] * n1[1]) - c[1] * (p21[0] * n1[2] - p21[2] * n1[0]) + c[2] * (p21[0] * n1[1] - p21[1] * n1[0])) / denominator;
  if (t1 < 0 || t2 < 0) return 0;
  prm[0] = (p1[0] + t1 * n1[0] + p2[0] + t2 * n2[0]) / 2.0;
  prm[1] = (p1[1] + t1 * n1[1] + p2[1] + t2 * n2[1]) / 2.0;
  prm[2] = (p1[2] + t1 * n1[2] + p2[2] + t2 * n2[2]) / 2.0;
  return 3;
}

/* Rows [R_i | -I], rhs -t_i: PivotCalibrationParametersEstimator.cxx:9-51 (n=3) and :63-96 */
static int pivot_solve(const double* d, size_t n, double* prm) {
  size_t i; int r, c, rank;
  double* A = (double*)calloc(3 * n * 6, sizeof(double));
  double* b = (double*)malloc(3 * n * sizeof(double));
  for (i = 0; i < n; i++) {
    const double* f = d + 12 * i;
    for (r = 0; r < 3; r++) {
      for (c = 0; c < 3; c++) A[(3 * i + r) * 6 + c] = f[3 * r + c];
      A[(3 * i + r) * 6 + 3 + r] = -1.0;
      b[3 * i + r] = -f[9 + r];
    }
  }
  rank = pinv_solve((int)(3 * n), 6, A, b, EPS, prm);
  free(A); free(b);
  return rank < 6 ? 0 : 6;
}

/* Rows [a_i | b_i]: DenseLinearEquationSystemParametersEstimator.hxx:17-49 (rows = nc) and :64-96 (rows >= nc):
 * x = pinv(A) b with singular values <= EPS zeroed, rank < nc -> no solution */
static int dense_solve(int nc, const double* d, size_t rows, double* prm) {
  size_t i; int j, rank;
  double* A = (double*)malloc(rows * nc * sizeof(double));
  double* b = (double*)malloc(rows * sizeof(double));
  for (i = 0; i < rows; i++) { for (j = 0; j < nc; j++) A[i * nc + j] = d[i * (nc + 1) + j]; b[i] = d[i * (nc + 1) + nc]; }
  rank = pinv_solve((int)rows, nc, A, b, EPS, prm);
  free(A); free(b);
  return rank < nc ? 0 : nc;
}



/* ------------------------------------------------------------------------------------ */
/* Cross-wire (single unknown point target) ultrasound calibration                      */
/* SinglePointTargetUSCalibrationParametersEstimator.cxx:10-329, 415-658                */
/* datum = [R2 row-major (9), t2 (3), u, v]; parameters (20) =                          */
/* [t1, t3, omega_z, omega_y, omega_x, m_x, m_y, m_x R3(:,1), m_y R3(:,2), R3(:,3)]     */
/* ------------------------------------------------------------------------------------ */

/* R <- U V^T of its SVD (the closest rotation in the Frobenius norm, .cxx:226-229), computed as the
 * orthogonal polar factor R (R^T R)^(-1/2) */
static void us_closest_rotation(double R[9]) {
  double S[9], V[9], ev[3], W[9], out[9];
  int i, j, k;
  for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) { double s = 0; for (k = 0; k < 3; k++) s += R[k * 3 + i] * R[k * 3 + j]; S[i * 3 + j] = s; }
  sym_eig(3, S, V, ev);
  for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) { double s = 0; for (k = 0; k < 3; k++) s += V[i * 3 + k] * V[j * 3 + k] / sqrt(ev[k]); W[i * 3 + j] = s; }
  for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) { double s = 0; for (k = 0; k < 3; k++) s += R[i * 3 + k] * W[k * 3 + j]; out[i * 3 + j] = s; }
  memcpy(R, out, sizeof(out));
}

/* .cxx:204-250 (and :862-900): scale factors, re-orthonormalised rotation and Euler angles from the first six
 * unknowns [m_x R3(:,1), m_y R3(:,2)] of the linear solution; out = [omega_z, omega_y, omega_x, m_x, m_y, m_x R3(:,1),
 * m_y R3(:,2), R3(:,3)] (14 values) */
static void us_rotation_part(const double x[6], double out[14]) {
  const double smallAngle = 0.008726535498373935, halfPI = 1.5707963267948966192313216916398;
  double r1[3], r2[3], r3[3], R3[9], m_x, m_y, inv, omega_z, omega_y, omega_x;
  int i;
  for (i = 0; i < 3; i++) { r1[i] = x[i]; r2[i] = x[3 + i]; }
  m_x = sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
  inv = 1.0 / m_x; for (i = 0; i < 3; i++) r1[i] = inv * r1[i];
  m_y = sqrt(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
  inv = 1.0 / m_y; for (i = 0; i < 3; i++) r2[i] = inv * r2[i];
  r3[0] = r1[1] * r2[2] - r1[2] * r2[1];
  r3[1] = r1[2] * r2[0] - r1[0] * r2[2];
  r3[2] = r1[0] * r2[1] - r1[1] * r2[0];
  for (i = 0; i < 3; i++) { R3[i * 3 + 0] = r1[i]; R3[i * 3 + 1] = r2[i]; R3[i * 3 + 2] = r3[i]; }
  us_closest_rotation(R3);
  omega_y = atan2(-R3[6], sqrt(R3[0] * R3[0] + R3[3] * R3[3]));
  if (fabs(omega_y - halfPI) > smallAngle && fabs(omega_y + halfPI) > smallAngle) {
    double cy = cos(omega_y);
    omega_z = atan2(R3[3] / cy, R3[0] / cy);
    omega_x = atan2(R3[7] / cy, R3[8] / cy);
  } else {
    omega_z = 0;
    omega_x = atan2(R3[1], R3[4]);
  }
  out[0] = omega_z; out[1] = omega_y; out[2] = omega_x; out[3] = m_x; out[4] = m_y;
  out[5] = m_x * R3[0]; out[6] = m_x * R3[3]; out[7] = m_x * R3[6];
  out[8] = m_y * R3[1]; out[9] = m_y * R3[4]; out[10] = m_y * R3[7];
  out[11] = R3[2]; out[12] = R3[5]; out[13] = R3[8];
}

/* cross-wire: x = [m_x R3(:,1), m_y R3(:,2), t3, t1] -> the 20 parameters (.cxx:251-268) */
static int us_post(const double x[12], double prm[20]) {
  int i;
  prm[0] = x[9]; prm[1] = x[10]; prm[2] = x[11];
  prm[3] = x[6]; prm[4] = x[7]; prm[5] = x[8];
  us_rotation_part(x, prm + 6);
  for (i = 0; i < 20; i++) if (!(prm[i] == prm[i])) return 0;
  return 20;
}

/* analyticLeastSquaresEstimate, .cxx:120-270: rows [u R2, v R2, R2, -I] x = -t2, pseudo-inverse with singular
 * values <= FLT_EPSILON zeroed, rank < 12 -> no solution */
static int usxw_analytic(const double* d, size_t n, double* prm) {
  size_t i; int r, c, rank;
  double x[12];
  double* A = (double*)calloc(3 * n * 12, sizeof(double));
  double* b = (double*)malloc(3 * n * sizeof(double));
  for (i = 0; i < n; i++) {
    const double* f = d + 14 * i;
    const double ui = f[12], vi = f[13];
    for (r = 0; r < 3; r++) {
      double* row = A + (3 * i + r) * 12;
      for (c = 0; c < 3; c++) { row[c] = f[3 * r + c] * ui; row[3 + c] = f[3 * r + c] * vi; row[6 + c] = f[3 * r + c]; }
      row[9 + r] = -1.0;
      b[3 * i + r] = -f[9 + r];
    }
  }
  rank = pinv_solve((int)(3 * n), 12, A, b, 1.192092896e-07, x);
  free(A); free(b);
  if (rank < 12) return 0;
  return us_post(x, prm);
}

/* e = R2 (u c1 + v c2 + t3) + t2 - t1 for the LM parameters x[11] (f(), .cxx:415-507) and, when J != NULL,
 * its 3 x 11 Jacobian (row-major); the scalar residuals |e_i| and their gradients that the reference hands to MINPACK
 * (f, gradf, .cxx:415-658) are formed from these in usxw_lm_fcn. */
static void us_residual(const double* f, const double* x, double e[3], double* J) {
  const double sz = sin(x[6]), cz = cos(x[6]), sy = sin(x[7]), cy = cos(x[7]), sx = sin(x[8]), cx = cos(x[8]);
  const double mx = x[9], my = x[10], u = f[12], v = f[13];
  const double c1[3] = {cz * cy, sz * cy, -sy};
  const double c2[3] = {cz * sy * sx - sz * cx, sz * sy * sx + cz * cx, cy * sx};
  double w[3];
  int r, k;
  for (k = 0; k < 3; k++) w[k] = u * mx * c1[k] + v * my * c2[k] + x[3 + k];
  for (r = 0; r < 3; r++) e[r] = f[3 * r] * w[0] + f[3 * r + 1] * w[1] + f[3 * r + 2] * w[2] + f[9 + r] - x[r];
  if (J) {
    /* d c1 / d omega_z, omega_y, omega_x and the same for c2 */
    const double dc1[3][3] = {{-sz * cy, cz * cy, 0}, {-cz * sy, -sz * sy, -cy}, {0, 0, 0}};
    const double dc2[3][3] = {{-sz * sy * sx - cz * cx, cz * sy * sx - sz * cx, 0}, {cz * cy * sx, sz * cy * sx, -sy * sx},
                              {cz * sy * cx + sz * sx, sz * sy * cx - cz * sx, cy * cx}};
    double dw[11][3];
    int p;
    memset(dw, 0, sizeof(dw));
    for (k = 0; k < 3; k++) {
      dw[3 + k][k] = 1.0;
      for (p = 0; p < 3; p++) dw[6 + p][k] = u * mx * dc1[p][k] + v * my * dc2[p][k];
      dw[9][k] = u * c1[k];
      dw[10][k] = v * c2[k];
    }
    for (r = 0; r < 3; r++)
      for (p = 0; p < 11; p++) J[r * 11 + p] = (p < 3) ? ((p == r) ? -1.0 : 0.0) : f[3 * r] * dw[p][0] + f[3 * r + 1] * dw[p][1] + f[3 * r + 2] * dw[p][2];
  }
}

/* The functor the reference hands to vnl_levenberg_marquardt (SumSquaresCalibrationPointsDistanceFunction, .cxx:402-658):
 * m scalar residuals d_i = |e_i| (f, :415-507) with Jacobian rows (e_i^T de_i/dx) / d_i (gradf, :510-658). */
typedef struct { const double* d; size_t n; } us_lm_ctx;
/* MINPACK's info and function-evaluation count of the last Levenberg-Marquardt run on this thread (for tests that need a
 * case away from the reference's evaluation cap) */
static __thread int g_lm_info = 0, g_lm_nfev = 0;
void orc_last_lm(int* info, int* nfev) { if (info) *info = g_lm_info; if (nfev) *nfev = g_lm_nfev; }
static int usxw_lm_fcn(void* user, int m, int n, const double* x, double* fvec, double* fjac, int ldfjac, int iflag) {
  const us_lm_ctx* c = (const us_lm_ctx*)user;
  int i, p, r;
  (void)n;
  for (i = 0; i < m; i++) {
    double e[3], J[33], dist;
    us_residual(c->d + 14 * (size_t)i, x, e, iflag == 2 ? J : NULL);
    dist = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    if (iflag == 1) fvec[i] = dist;
    else for (p = 0; p < 11; p++) { double t = 0; for (r = 0; r < 3; r++) t += J[r * 11 + p] * e[r]; fjac[i + (size_t)ldfjac * p] = t / dist; }
  }
  return 0;
}

/* iterativeLeastSquaresEstimate, .cxx:272-329: vnl_levenberg_marquardt (= MINPACK lmder, minpack_lm.h) from the analytic
 * estimate, all three tolerances 10e-16, at most 5000 function evaluations, parameters returned only if minimize()
 * reports success (info 1..4); entries 11..19 rebuilt from the angles and scales. */
static int usxw_iterative(const double* d, size_t n, const double* init, double* prm) {
  us_lm_ctx c;
  double x[11], *fvec;
  int a, info;
  if (n < 11) return 0;   /* vnl_levenberg_marquardt: fewer residuals than unknowns -> failure */
  c.d = d; c.n = n;
  for (a = 0; a < 11; a++) x[a] = init[a];
  fvec = (double*)malloc(sizeof(double) * n);
  info = mpk_lmder(usxw_lm_fcn, &c, (int)n, 11, x, fvec, 10e-16, 10e-16, 10e-16, 5000, 100.0, &g_lm_nfev, NULL);
  g_lm_info = info;
  free(fvec);
  if (info < 1 || info > 4) return 0;
  for (a = 0; a < 11; a++) prm[a] = x[a];
  {
    const double cz = cos(x[6]), sz = sin(x[6]), cy = cos(x[7]), sy = sin(x[7]), cx = cos(x[8]), sx = sin(x[8]);
    prm[11] = x[9] * cz * cy; prm[12] = x[9] * sz * cy; prm[13] = -x[9] * sy;
    prm[14] = x[10] * (cz * sy * sx - sz * cx); prm[15] = x[10] * (sz * sy * sx + cz * cx); prm[16] = x[10] * cy * sx;
    prm[17] = cz * sy * cx + sz * sx; prm[18] = sz * sy * cx - cz * sx; prm[19] = cy * cx;
  }
  return 20;
}

/* leastSquaresEstimate, .cxx:36-59 */
static int usxw_lsq(int ls_type, const double* d, size_t n, double* prm) {
  double init[20];
  if (ls_type == 0) return usxw_analytic(d, n, prm);
  if (!usxw_analytic(d, n, init)) return 0;
  return usxw_iterative(d, n, init, prm);
}

/* agree, .cxx:71-107: qInT = (T2*T3)*q with VNL's left-to-right accumulation (terms that multiply the
 * constant 0 / 1 entries of the homogeneous matrices are exact and left out) */
static int usxw_agree(const double* prm, const double* f, double delta) {
  const double u = f[12], v = f[13];
  double err[3], s = 0;
  int i;
  for (i = 0; i < 3; i++) {
    const double a = f[3 * i], b = f[3 * i + 1], c = f[3 * i + 2];
    const double M0 = a * prm[11] + b * prm[12] + c * prm[13];
    const double M1 = a * prm[14] + b * prm[15] + c * prm[16];
    const double M3 = a * prm[3] + b * prm[4] + c * prm[5] + f[9 + i];
    err[i] = (M0 * u + M1 * v + M3) - prm[i];
  }
  s = err[0] * err[0] + err[1] * err[1] + err[2] * err[2];
  return s < delta * delta;
}

/* ------------------------------------------------------------------------------------ */
/* Calibrated-pointer ultrasound calibration                                            */
/* SinglePointTargetUSCalibrationParametersEstimator.cxx:663-985                        */
/* datum = [R2 (9), t2 (3), u, v, p (3)]; parameters (17) =                             */
/* [t3, omega_z, omega_y, omega_x, m_x, m_y, m_x R3(:,1), m_y R3(:,2), R3(:,3)]         */
/* ------------------------------------------------------------------------------------ */

/* analyticLeastSquaresEstimate, .cxx:789-920: rows [u R2, v R2, R2] x = p - t2, singular values <= FLT_EPSILON
 * zeroed, rank < 9 -> no solution */
static int uscp_analytic(const double* d, size_t n, double* prm) {
  size_t i; int r, c, rank;
  double x[9];
  double* A = (double*)calloc(3 * n * 9, sizeof(double));
  double* b = (double*)malloc(3 * n * sizeof(double));
  for (i = 0; i < n; i++) {
    const double* f = d + 17 * i;
    const double ui = f[12], vi = f[13];
    for (r = 0; r < 3; r++) {
      double* row = A + (3 * i + r) * 9;
      for (c = 0; c < 3; c++) { row[c] = f[3 * r + c] * ui; row[3 + c] = f[3 * r + c] * vi; row[6 + c] = f[3 * r + c]; }
      b[3 * i + r] = f[14 + r] - f[9 + r];
    }
  }
  rank = pinv_solve((int)(3 * n), 9, A, b, 1.192092896e-07, x);
  free(A); free(b);
  if (rank < 9) return 0;
  prm[0] = x[6]; prm[1] = x[7]; prm[2] = x[8];
  us_rotation_part(x, prm + 3);
  for (r = 0; r < 17; r++) if (!(prm[r] == prm[r])) return 0;
  return 17;
}

/* e = R2 (u m_x c1 + v m_y c2 + t3) + t2 - p at x[8] = [t3, omega_z, omega_y, omega_x, m_x, m_y] and its 3 x 8
 * Jacobian: the cross-wire residual with t1 replaced by the measured p and no unknown for it */
static void uscp_residual(const double* f, const double* x, double e[3], double* J) {
  double xx[11], ee[3], JJ[33];
  int r, p;
  xx[0] = f[14]; xx[1] = f[15]; xx[2] = f[16];
  for (p = 0; p < 8; p++) xx[3 + p] = x[p];
  us_residual(f, xx, ee, J ? JJ : NULL);
  for (r = 0; r < 3; r++) { e[r] = ee[r]; if (J) for (p = 0; p < 8; p++) J[r * 8 + p] = JJ[r * 11 + 3 + p]; }
}

static int uscp_lm_fcn(void* user, int m, int n, const double* x, double* fvec, double* fjac, int ldfjac, int iflag) {
  const us_lm_ctx* c = (const us_lm_ctx*)user;
  int i, p, r;
  (void)n;
  for (i = 0; i < m; i++) {
    double e[3], J[24], dist;
    uscp_residual(c->d + 17 * (size_t)i, x, e, iflag == 2 ? J : NULL);
    dist = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    if (iflag == 1) fvec[i] = dist;
    else for (p = 0; p < 8; p++) { double t = 0; for (r = 0; r < 3; r++) t += J[r * 8 + p] * e[r]; fjac[i + (size_t)ldfjac * p] = t / dist; }
  }
  return 0;
}

/* iterativeLeastSquaresEstimate, .cxx:916-985: as the cross-wire one with 8 unknowns and all three tolerances 10e-8, so
 * MINPACK stops as soon as the scaled gradient, the relative step or the relative reductions fall below 1e-7 -- often at
 * the analytic start itself (info 4). */
static int uscp_iterative(const double* d, size_t n, const double* init, double* prm) {
  us_lm_ctx c;
  double x[8], *fvec;
  int a, info;
  if (n < 8) return 0;
  c.d = d; c.n = n;
  for (a = 0; a < 8; a++) x[a] = init[a];
  fvec = (double*)malloc(sizeof(double) * n);
  info = mpk_lmder(uscp_lm_fcn, &c, (int)n, 8, x, fvec, 10e-8, 10e-8, 10e-8, 5000, 100.0, &g_lm_nfev, NULL);
  g_lm_info = info;
  free(fvec);
  if (info < 1 || info > 4) return 0;
  for (a = 0; a < 8; a++) prm[a] = x[a];
  {
    const double cz = cos(x[3]), sz = sin(x[3]), cy = cos(x[4]), sy = sin(x[4]), cx = cos(x[5]), sx = sin(x[5]);
    prm[8] = x[6] * cz * cy; prm[9] = x[6] * sz * cy; prm[10] = -x[6] * sy;
    prm[11] = x[7] * (cz * sy * sx - sz * cx); prm[12] = x[7] * (sz * sy * sx + cz * cx); prm[13] = x[7] * cy * sx;
    prm[14] = cz * sy * cx + sz * sx; prm[15] = sz * sy * cx - cz * sx; prm[16] = cy * cx;
  }
  return 17;
}

static int uscp_lsq(int ls_type, const double* d, size_t n, double* prm) {
  double init[17];
  if (ls_type == 0) return uscp_analytic(d, n, prm);
  if (!uscp_analytic(d, n, init)) return 0;
  return uscp_iterative(d, n, init, prm);
}

/* agree, .cxx:726-761 */
static int uscp_agree(const double* prm, const double* f, double delta) {
  const double u = f[12], v = f[13];
  double err[3], s;
  int i;
  for (i = 0; i < 3; i++) {
    const double a = f[3 * i], b = f[3 * i + 1], c = f[3 * i + 2];
    const double M0 = a * prm[8] + b * prm[9] + c * prm[10];
    const double M1 = a * prm[11] + b * prm[12] + c * prm[13];
    const double M3 = a * prm[0] + b * prm[1] + c * prm[2] + f[9 + i];
    err[i] = (M0 * u + M1 * v + M3) - f[14 + i];
  }
  s = err[0] * err[0] + err[1] * err[1] + err[2] * err[2];
  return s < delta * delta;
}

static double ray_cross_eps(double aux) {
  double a = aux > 0 ? aux : 0.017453292519943295769236907684886; /* RayIntersectionParametersEstimator.h:35 */
  double s = sin(a);
  return s * s;
}

int orc_estimate(int model, double delta, double aux, const double* data, size_t n, double* params) {
  int D, P, k;
  if (orc_model_info(model, &D, &P, &k)) return -1;
  if (n < (size_t)k) return 0;
  if (model > 14) {   /* the wider template space: the same generic bodies, `dimension` from the model id */
    int dim;
    switch (family(model, &dim)) {
      case F_PLANE_ND: return plane_nd_estimate(dim, data, params);
      case F_SPHERE_ND: return sphere_nd_estimate(dim, data, params);
      case F_LINE: return line_estimate(dim, data, delta * delta, params);
      case F_DENSE: return dense_solve(dim, data, (size_t)dim, params);
      default: return -1;
    }
  }
  switch (model) {
    case M_PLANE3: return plane3_estimate(data, params);
    case M_LINE2D: return line2d_estimate(data, delta * delta, params);
    case M_LINE2: return line_estimate(2, data, delta * delta, params);
    case M_LINE3: return line_estimate(3, data, delta * delta, params);
    case M_CIRCLE2: return circle_estimate(data, params);
    case M_SPHERE3: return sphere3_estimate(data, params);
    case M_SPHERE4: return sphere_nd_estimate(4, data, params);
    case M_PLANE4: return plane_nd_estimate(4, data, params);
    case M_ABSOR: return absor_estimate(data, params);
    case M_RAY: return ray_estimate(data, ray_cross_eps(aux), params);
    case M_PIVOT: return pivot_solve(data, 3, params);
    case M_DENSE5: return dense_solve(5, data, 5, params);
    case M_DENSE6: return dense_solve(6, data, 6, params);
    case M_USXW: return (n == 4) ? usxw_analytic(data, 4, params) : 0;   /* estimate() insists on exactly 4 data (.cxx:21-22) */
    case M_USCP: return (n == 3) ? uscp_analytic(data, 3, params) : 0;   /* exactly 3 (.cxx:674-675) */
  }
  return -1;
}

/* ------------------------------------------------------------------------------------ */
/* agree()                                                                              */
/* ------------------------------------------------------------------------------------ */

static int agree1(int model, double delta, const double* prm, const double* x) {
  if (model > 14) {
    int dim, i;
    switch (family(model, &dim)) {
      case F_PLANE_ND: { /* PlaneParametersEstimator.hxx:196-203 */
        double sd = 0;
        for (i = 0; i < dim; i++) sd += prm[i] * (x[i] - prm[dim + i]);
        return (sd * sd) < delta * delta;
      }
      case F_LINE: { /* LineParametersEstimator.hxx:135-150 */
        double v[MAXD], vDotN = 0.0, ds = 0.0;
        for (i = 0; i < dim; i++) { v[i] = x[i] - prm[dim + i]; vDotN += v[i] * prm[i]; }
        for (i = 0; i < dim; i++) ds += (v[i] - vDotN * prm[i]) * (v[i] - vDotN * prm[i]);
        return ds < delta * delta;
      }
      case F_SPHERE_ND: { /* SphereParametersEstimator.hxx:255-264 */
        double dl = 0;
        for (i = 0; i < dim; i++) dl += ((x[i] - prm[i]) * (x[i] - prm[i]));
        dl = fabs(sqrt(dl) - prm[dim]);
        return dl < delta;
      }
      case F_DENSE: { /* DenseLinearEquationSystemParametersEstimator.hxx:111-119 */
        double sum = 0.0;
        for (i = 0; i < dim; i++) sum += x[i] * prm[i];
        sum -= x[dim];
        return fabs(sum) < delta;
      }
      default: return 0;
    }
  }
  switch (model) {
    case M_PLANE3:
    case M_PLANE4: { /* PlaneParametersEstimator.hxx:196-203 */
      const int dim = (model == M_PLANE3) ? 3 : 4;
      double sd = 0; int i;
      for (i = 0; i < dim; i++) sd += prm[i] * (x[i] - prm[dim + i]);
      return (sd * sd) < delta * delta;
    }
    case M_LINE2D: { /* Line2DParametersEstimator.cxx:119-123 */
      double sd = prm[0] * (x[0] - prm[2]) + prm[1] * (x[1] - prm[3]);
      return (sd * sd) < delta * delta;
    }
    case M_LINE2:
    case M_LINE3: { /* LineParametersEstimator.hxx:135-150 */
      int dim = (model == M_LINE2) ? 2 : 3, i;
      double v[3], vDotN = 0.0, ds = 0.0;
      for (i = 0; i < dim; i++) { v[i] = x[i] - prm[dim + i]; vDotN += v[i] * prm[i]; }
      for (i = 0; i < dim; i++) ds += (v[i] - vDotN * prm[i]) * (v[i] - vDotN * prm[i]);
      return ds < delta * delta;
    }
    case M_CIRCLE2:
    case M_SPHERE3:
    case M_SPHERE4: { /* SphereParametersEstimator.hxx:255-264 (distance vs delta, not squared) */
      int dim = (model == M_CIRCLE2) ? 2 : (model == M_SPHERE3 ? 3 : 4), i;
      double dl = 0;
      for (i = 0; i < dim; i++) dl += ((x[i] - prm[i]) * (x[i] - prm[i]));
      dl = fabs(sqrt(dl) - prm[dim]);
      return dl < delta;
    }
    case M_ABSOR: { /* AbsoluteOrientationParametersEstimator.cxx:316-327 */
      double R[9], q[3], dx, dy, dz;
      quat_to_matrix(prm[0], prm[1], prm[2], prm[3], 0, R);
      frame_apply(R, prm + 4, x, q);
      dx = q[0] - x[3]; dy = q[1] - x[4]; dz = q[2] - x[5];
      return (dx * dx + dy * dy + dz * dz) < delta * delta;
    }
    case M_RAY: { /* RayIntersectionParametersEstimator.cxx:164-179 */
      const double *p = x, *n = x + 3;
      double t = n[0] * (prm[0] - p[0]) + n[1] * (prm[1] - p[1]) + n[2] * (prm[2] - p[2]);
      double dx = prm[0] - p[0] - t * n[0];
      double dy = prm[1] - p[1] - t * n[1];
      double dz = prm[2] - p[2] - t * n[2];
      return t >= 0 && (dx * dx + dy * dy + dz * dz < delta * delta);
    }
    case M_PIVOT: { /* PivotCalibrationParametersEstimator.cxx:108-123; l2Norm: common/Vector.h:134-139 */
      double q[3], r[3], s = 0; int i;
      frame_apply(x, x + 9, prm, q);
      for (i = 0; i < 3; i++) r[i] = q[i] - prm[3 + i];
      for (i = 0; i < 3; i++) s += r[i] * r[i];
      return sqrt(s) < delta;
    }
    case M_USXW: return usxw_agree(prm, x, delta);
    case M_USCP: return uscp_agree(prm, x, delta);
    case M_DENSE5:
    case M_DENSE6: { /* DenseLinearEquationSystemParametersEstimator.hxx:111-119 */
      const int nc = (model == M_DENSE5) ? 5 : 6;
      double sum = 0.0; int i;
      for (i = 0; i < nc; i++) sum += x[i] * prm[i];
      sum -= x[nc];
      return fabs(sum) < delta;
    }
  }
  return 0;
}

int orc_agree(int model, double delta, double aux, const double* params, int np, const double* data, size_t n, uint8_t* out) {
  int D, P, k, c = 0; size_t i;
  (void)aux; (void)np;
  if (orc_model_info(model, &D, &P, &k)) return -1;
  for (i = 0; i < n; i++) { int a = agree1(model, delta, params, data + i * D); if (out) out[i] = (uint8_t)a; c += a; }
  return c;
}

/* ------------------------------------------------------------------------------------ */
/* leastSquaresEstimate()                                                               */
/* ------------------------------------------------------------------------------------ */

/* PlaneParametersEstimator.hxx:129-172 (col=0: smallest eigenvalue) and
 * LineParametersEstimator.hxx:68-111 (col=dim-1: largest) share the covariance build. */
static int cov_eig_estimate(int dim, int col, const double* d, size_t n, double* prm) {
  double mean[MAXD] = {0}, cov[MAXD * MAXD] = {0}, meanMat[MAXD * MAXD], V[MAXD * MAXD], ev[MAXD], sqrtN = sqrt((double)n);
  size_t i; int j, k;
  for (i = 0; i < n; i++) for (j = 0; j < dim; j++) mean[j] += d[i * dim + j];
  for (j = 0; j < dim; j++) mean[j] /= sqrtN;
  for (j = 0; j < dim; j++) for (k = j; k < dim; k++) meanMat[j * dim + k] = meanMat[k * dim + j] = mean[j] * mean[k];
  for (i = 0; i < n; i++) for (j = 0; j < dim; j++) for (k = j; k < dim; k++) cov[j * dim + k] += d[i * dim + j] * d[i * dim + k];
  for (j = 0; j < dim; j++) for (k = j + 1; k < dim; k++) cov[k * dim + j] = cov[j * dim + k];
  for (j = 0; j < dim * dim; j++) cov[j] -= meanMat[j];
  sym_eig(dim, cov, V, ev);
  for (j = 0; j < dim; j++) prm[j] = V[j * dim + col];
  for (j = 0; j < dim; j++) prm[dim + j] = mean[j] / sqrtN;
  return 2 * dim;
}

/* Line2DParametersEstimator.cxx:50-100 */
static int line2d_lsq(const double* d, size_t n, double* prm) {
  double meanX = 0.0, meanY = 0.0, nx, ny, norm, c11 = 0, c12 = 0, c22 = 0;
  int i, dataSize = (int)n;
  for (i = 0; i < dataSize; i++) {
    meanX += d[2 * i]; meanY += d[2 * i + 1];
    c11 += d[2 * i] * d[2 * i]; c12 += d[2 * i] * d[2 * i + 1]; c22 += d[2 * i + 1] * d[2 * i + 1];
  }
  meanX /= dataSize; meanY /= dataSize;
  c11 -= dataSize * meanX * meanX;
  c12 -= dataSize * meanX * meanY;
  c22 -= dataSize * meanY * meanY;
  if (c11 < 1e-12) {
    nx = 1.0; ny = 0.0;
    if (c22 < 1e-12) return 0;
  } else {
    double lambda1 = (c11 + c22 + sqrt((c11 - c22) * (c11 - c22) + 4 * c12 * c12)) / 2.0;
    nx = -c12; ny = lambda1 - c22;
    norm = sqrt(nx * nx + ny * ny);
    nx /= norm; ny /= norm;
  }
  prm[0] = nx; prm[1] = ny; prm[2] = meanX; prm[3] = meanY;
  return 4;
}

/* SphereParametersEstimator.hxx:267-307 */
static int sphere_algebraic(int dim, const double* d, size_t n, double* prm) {
  int cols = dim + 1, j, rank;
  size_t i;
  double x[MAXD], rSquared;
  double* A = (double*)malloc(sizeof(double) * n * cols);
  double* b = (double*)calloc(n, sizeof(double));
  for (i = 0; i < n; i++) {
    for (j = 0; j < dim; j++) { A[i * cols + j] = -2 * d[i * dim + j]; b[i] += -(d[i * dim + j] * d[i * dim + j]); }
    A[i * cols + dim] = 1;
  }
  rank = pinv_solve((int)n, cols, A, b, EPS, x);
  free(A); free(b);
  if (rank < dim + 1) return 0;
  rSquared = -x[dim];
  for (j = 0; j < dim; j++) { prm[j] = x[j]; rSquared += x[j] * x[j]; }
  if (rSquared > 0) { prm[dim] = sqrt(rSquared); return dim + 1; }
  return 0;
}

/* residuals f (SphereParametersEstimator.hxx:394-409) and the analytic Jacobian gradf (:413-431): J_i = [(c-p_i)/|p_i-c|, -1]. */
typedef struct { const double* d; int dim; } sphere_lm_ctx;
static int sphere_lm_fcn(void* user, int m, int n, const double* x, double* fvec, double* fjac, int ldfjac, int iflag) {
  const sphere_lm_ctx* c = (const sphere_lm_ctx*)user;
  const int dim = c->dim;
  int i, j;
  (void)n;
  for (i = 0; i < m; i++) {
    const double* pt = c->d + (size_t)i * dim;
    double sq = 0, sv;
    for (j = 0; j < dim; j++) sq += (pt[j] - x[j]) * (pt[j] - x[j]);
    sv = sqrt(sq);
    if (iflag == 1) fvec[i] = sv - x[dim];
    else { for (j = 0; j < dim; j++) fjac[i + (size_t)ldfjac * j] = (x[j] - pt[j]) / sv; fjac[i + (size_t)ldfjac * dim] = -1.0; }
  }
  return 0;
}

/* SphereParametersEstimator.hxx:310-338: vnl_levenberg_marquardt (= MINPACK lmder, minpack_lm.h) with xtol = gtol = 10e-16,
 * ftol left at VNL's default (xtol_default * 0.01 = 1e-10), maxfev = 500; parameters are returned only when minimize()
 * reports success. */
static int sphere_geometric(int dim, const double* d, size_t n, const double* init, double* prm) {
  sphere_lm_ctx c;
  double x[MAXD], *fvec;
  int p = dim + 1, a, info;
  if ((int)n < p) return 0;
  c.d = d; c.dim = dim;
  for (a = 0; a < p; a++) x[a] = init[a];
  fvec = (double*)malloc(sizeof(double) * n);
  info = mpk_lmder(sphere_lm_fcn, &c, (int)n, p, x, fvec, 1e-8 * 0.01, 10e-16, 10e-16, 500, 100.0, &g_lm_nfev, NULL);
  g_lm_info = info;
  free(fvec);
  if (info < 1 || info > 4) return 0;
  for (a = 0; a < p; a++) prm[a] = x[a];
  return p;
}

/* SphereParametersEstimator.hxx:209-232 */
static int sphere_lsq(int dim, int ls_type, const double* d, size_t n, double* prm) {
  double init[MAXD];
  if (ls_type == 0) return sphere_algebraic(dim, d, n, prm);
  if (!sphere_algebraic(dim, d, n, init)) return 0;
  return sphere_geometric(dim, d, n, init, prm);
}

/* AbsoluteOrientationParametersEstimator.cxx:120-206 (Horn's quaternion method) */
static int absor_lsq(const double* d, size_t n, double* prm) {
  double mF[3] = {0, 0, 0}, mS[3] = {0, 0, 0}, mu[9], M[9] = {0}, N[16] = {0}, tmp[9], V[16], ev[4], R[9], tF[3], zero[3] = {0, 0, 0};
  double A12, A20, A01, traceM = 0.0;
  size_t i; int r, c;
  for (i = 0; i < n; i++) for (r = 0; r < 3; r++) { mF[r] += d[6 * i + r]; mS[r] += d[6 * i + 3 + r]; }
  for (r = 0; r < 3; r++) { mF[r] /= (unsigned int)n; mS[r] /= (unsigned int)n; }
  for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) mu[r * 3 + c] = mF[r] * mS[c];
  for (i = 0; i < n; i++) for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) M[r * 3 + c] += d[6 * i + r] * d[6 * i + 3 + c];
  for (r = 0; r < 9; r++) M[r] += mu[r] * (double)(-(int)n);
  for (r = 0; r < 3; r++) traceM += M[r * 3 + r];
  for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) tmp[r * 3 + c] = ((r == c) ? -traceM : 0.0) + (M[r * 3 + c] + M[c * 3 + r]);
  A12 = M[5] - M[7]; A20 = M[6] - M[2]; A01 = M[1] - M[3];
  N[0] = traceM; N[1] = A12; N[2] = A20; N[3] = A01;
  N[4] = A12; N[8] = A20; N[12] = A01;
  for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) N[(r + 1) * 4 + c + 1] = tmp[r * 3 + c];
  sym_eig(4, N, V, ev);
  for (r = 0; r < 4; r++) prm[r] = V[r * 4 + 3];
  quat_to_matrix(prm[0], prm[1], prm[2], prm[3], 1, R);
  frame_apply(R, zero, mF, tF);
  for (r = 0; r < 3; r++) prm[4 + r] = mS[r] - tF[r];
  return 7;
}

/* AbsoluteOrientationParametersEstimator.cxx:208-297: Horn's method with one weight per pair */
int orc_weighted_absor(const double* d, size_t n, const double* w, double* prm) {
  double mF[3] = {0, 0, 0}, mS[3] = {0, 0, 0}, mu[9], M[9] = {0}, N[16] = {0}, tmp[9], V[16], ev[4], R[9], tF[3], zero[3] = {0, 0, 0};
  double A12, A20, A01, traceM = 0.0, sumWeights = 0.0;
  size_t i; int r, c;
  if (n < 3) return 0;
  for (i = 0; i < n; i++) sumWeights += w[i];
  for (i = 0; i < n; i++) for (r = 0; r < 3; r++) { mF[r] += d[6 * i + r] * w[i]; mS[r] += d[6 * i + 3 + r] * w[i]; }
  for (r = 0; r < 3; r++) { mF[r] /= sumWeights; mS[r] /= sumWeights; }
  for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) mu[r * 3 + c] = mF[r] * mS[c];
  for (i = 0; i < n; i++) for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) M[r * 3 + c] += (d[6 * i + r] * d[6 * i + 3 + c]) * w[i];
  for (r = 0; r < 9; r++) M[r] += mu[r] * (-sumWeights);
  for (r = 0; r < 3; r++) traceM += M[r * 3 + r];
  for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) tmp[r * 3 + c] = ((r == c) ? -traceM : 0.0) + (M[r * 3 + c] + M[c * 3 + r]);
  A12 = M[5] - M[7]; A20 = M[6] - M[2]; A01 = M[1] - M[3];
  N[0] = traceM; N[1] = A12; N[2] = A20; N[3] = A01;
  N[4] = A12; N[8] = A20; N[12] = A01;
  for (r = 0; r < 3; r++) for (c = 0; c < 3; c++) N[(r + 1) * 4 + c + 1] = tmp[r * 3 + c];
  sym_eig(4, N, V, ev);
  for (r = 0; r < 4; r++) prm[r] = V[r * 4 + 3];
  quat_to_matrix(prm[0], prm[1], prm[2], prm[3], 1, R);
  frame_apply(R, zero, mF, tF);
  for (r = 0; r < 3; r++) prm[4 + r] = mS[r] - tF[r];
  return 7;
}

/* RayIntersectionParametersEstimator.cxx:100-144 */
static int ray_lsq(const double* d, size_t m, double* prm) {
  double A[9] = {0}, b[3] = {0, 0, 0};
  size_t i; int rank;
  for (i = 0; i < m; i++) {
    const double *p = d + 6 * i, *n = p + 3;
    double s;
    A[0] += -(n[0] * n[0]); A[1] += -(n[0] * n[1]); A[2] += -(n[0] * n[2]);
    A[4] += -(n[1] * n[1]); A[5] += -(n[1] * n[2]); A[8] += -(n[2] * n[2]);
    s = n[0] * p[0] + n[1] * p[1] + n[2] * p[2];
    b[0] += p[0] - s * n[0]; b[1] += p[1] - s * n[1]; b[2] += p[2] - s * n[2];
  }
  A[0] += m; A[3] = A[1]; A[4] += m; A[6] = A[2]; A[7] = A[5]; A[8] += m;
  rank = pinv_solve(3, 3, A, b, EPS, prm);
  return rank < 3 ? 0 : 3;
}

int orc_least_squares(int model, double delta, double aux, int ls_type, const double* data, size_t n, double* params) {
  int D, P, k;
  (void)delta; (void)aux;
  if (orc_model_info(model, &D, &P, &k)) return -1;
  if (n < (size_t)k && model != M_RAY) return 0; /* ray LS has no size guard (:100-144) */
  if (model > 14) {
    int dim;
    switch (family(model, &dim)) {
      case F_PLANE_ND: return cov_eig_estimate(dim, 0, data, n, params);          /* PlaneParametersEstimator.hxx:129-172 */
      case F_LINE: return cov_eig_estimate(dim, dim - 1, data, n, params);        /* LineParametersEstimator.hxx:68-111 */
      case F_SPHERE_ND: return sphere_lsq(dim, ls_type, data, n, params);
      case F_DENSE: return dense_solve(dim, data, n, params);
      default: return -1;
    }
  }
  switch (model) {
    case M_PLANE3: return cov_eig_estimate(3, 0, data, n, params);
    case M_PLANE4: return cov_eig_estimate(4, 0, data, n, params);
    case M_LINE2D: return line2d_lsq(data, n, params);
    case M_LINE2: return cov_eig_estimate(2, 1, data, n, params);
    case M_LINE3: return cov_eig_estimate(3, 2, data, n, params);
    case M_CIRCLE2: return sphere_lsq(2, ls_type, data, n, params);
    case M_SPHERE3: return sphere_lsq(3, ls_type, data, n, params);
    case M_SPHERE4: return sphere_lsq(4, ls_type, data, n, params);
    case M_ABSOR: return absor_lsq(data, n, params);
    case M_RAY: return ray_lsq(data, n, params);
    case M_PIVOT: return pivot_solve(data, n, params);
    case M_DENSE5: return dense_solve(5, data, n, params);
    case M_DENSE6: return dense_solve(6, data, n, params);
    case M_USXW: return usxw_lsq(ls_type, data, n, params);
    case M_USCP: return uscp_lsq(ls_type, data, n, params);
  }
  return -1;
}

/* ------------------------------------------------------------------------------------ */
/* Driver pieces (parametersEstimators/RANSAC.hxx)                                      */
/* ------------------------------------------------------------------------------------ */

/* RANSAC.hxx:217-249 body for one subset (full scoring, no early exit). */
static uint32_t score_one(int model, int D, int k, double delta, double aux, const double* data, size_t n,
                          const int32_t* sub, double* prm, int* nprm) {
  double pts[10 * 17];   /* up to 9 data (8-D hypersphere) of up to 17 doubles */
  int j; size_t m; uint32_t c = 0;
  for (j = 0; j < k; j++) memcpy(pts + j * D, data + (size_t)sub[j] * D, sizeof(double) * D);
  *nprm = orc_estimate(model, delta, aux, pts, (size_t)k, prm);
  if (*nprm <= 0) return 0;
  for (m = 0; m < n; m++) c += (uint32_t)agree1(model, delta, prm, data + m * D);
  return c;
}

int orc_score_subsets(int model, double delta, double aux, const double* data, size_t n, const int32_t* subsets, size_t H,
                      uint32_t* counts, double* params_out, int nthreads) {
  int D, P, k; long long h;
  if (orc_model_info(model, &D, &P, &k)) return -1;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
  for (h = 0; h < (long long)H; h++) {
    double prm[20]; int np, j;
    uint32_t c = score_one(model, D, k, delta, aux, data, n, subsets + h * k, prm, &np);
    if (counts) counts[h] = c;
    if (params_out) for (j = 0; j < P; j++) params_out[h * P + j] = (np > 0) ? prm[j] : NAN;
  }
  (void)nthreads;
  return 0;
}

/* RANSAC.hxx:254-280 */
unsigned int orc_choose(unsigned int n, unsigned int m) {
  double denominatorEnd, numeratorStart, numerator, denominator, i, result;
  if ((n - m) > m) { numeratorStart = n - m + 1; denominatorEnd = m; }
  else { numeratorStart = m + 1; denominatorEnd = n - m; }
  for (i = numeratorStart, numerator = 1; i <= n; i++) numerator *= i;
  for (i = 1, denominator = 1; i <= denominatorEnd; i++) denominator *= i;
  result = numerator / denominator;
  if (denominator > 1.7976931348623157e308 || numerator > 1.7976931348623157e308 || (double)UINT_MAX < result) return UINT_MAX;
  return (unsigned int)result;
}

static uint64_t binom64(unsigned int n, unsigned int k) {
  uint64_t r = 1; unsigned int i;
  if (k > n) return 0;
  for (i = 0; i < k; i++) r = r * (n - i) / (i + 1);
  return r;
}

/* Enumeration order of computeAllChoices, RANSAC.hxx:197-213: ascending index tuples in
 * lexicographic order.  Subsets starting with c number C(n-1-c, k-1). */
void orc_unrank_lex(uint64_t rank, unsigned int n, unsigned int k, int32_t* out) {
  unsigned int j, c = 0;
  for (j = 0; j < k; j++) {
    for (;; c++) {
      uint64_t cnt = binom64(n - 1 - c, k - 1 - j);
      if (rank < cnt) break;
      rank -= cnt;
    }
    out[j] = (int32_t)c++;
  }
}

/* RANSAC.hxx:107-110 */
unsigned int orc_num_tries(double prob, unsigned int votes, unsigned int n, unsigned int k, unsigned int all_tries) {
  double numerator = log(1.0 - prob);
  double denominator = log(1.0 - pow((double)votes / (double)n, (double)(k)));
  unsigned int numTries = (unsigned int)(int)(numerator / denominator + 0.5);
  return numTries < all_tries ? numTries : all_tries;
}

/* RANSAC.hxx:150-192 + :197-213 + :217-249 + least squares on the consensus set :175-185 */
int orc_ransac_exhaustive(int model, double delta, double aux, int ls_type, const double* data, size_t n,
                          double* params, uint8_t* mask, double* fraction, uint32_t* best_count, uint64_t* best_rank) {
  int D, P, k, np = 0, j;
  int32_t sub[10];
  uint32_t best = 0; uint64_t rank = 0, brank = 0; double bprm[20], prm[20];
  size_t m, nin = 0;
  double* inl;
  if (orc_model_info(model, &D, &P, &k)) return -1;
  *fraction = 0; if (best_count) *best_count = 0; if (best_rank) *best_rank = 0;
  if (n < (size_t)k) return 0;
  for (j = 0; j < k; j++) sub[j] = j;
  for (;;) {
    int npc; uint32_t c = score_one(model, D, k, delta, aux, data, n, sub, prm, &npc);
    if (c > best) { best = c; brank = rank; memcpy(bprm, prm, sizeof(prm)); }
    rank++;
    /* next ascending tuple in lexicographic order */
    for (j = k - 1; j >= 0 && sub[j] == (int32_t)(n - k + j); j--) {}
    if (j < 0) break;
    sub[j]++;
    for (j = j + 1; j < k; j++) sub[j] = sub[j - 1] + 1;
  }
  if (best_count) *best_count = best;
  if (best_rank) *best_rank = brank;
  if (best == 0) return 0;
  inl = (double*)malloc(sizeof(double) * n * D);
  for (m = 0; m < n; m++) {
    int a = agree1(model, delta, bprm, data + m * D);
    if (mask) mask[m] = (uint8_t)a;
    if (a) { memcpy(inl + nin * D, data + m * D, sizeof(double) * D); nin++; }
  }
  np = orc_least_squares(model, delta, aux, ls_type, inl, nin, params);
  free(inl);
  *fraction = (double)best / (double)(unsigned int)n;
  return np;
}
