/* TEST INFRASTRUCTURE ONLY -- restatement of MINPACK's Levenberg-Marquardt driver `lmder` and its helpers
 * `lmpar`, `qrfac`, `qrsolv`, `enorm` (J. J. More', B. S. Garbow, K. E. Hillstrom, "User Guide for MINPACK-1",
 * Argonne National Laboratory report ANL-80-74, 1980; the algorithm of J. J. More', "The Levenberg-Marquardt
 * algorithm: implementation and theory", Lecture Notes in Mathematics 630, 1978).
 *
 * Why it is here: every Levenberg-Marquardt call of the reference goes through VNL's vnl_levenberg_marquardt
 * (SphereParametersEstimator.hxx:321-331, SinglePointTargetUSCalibrationParametersEstimator.cxx:284-297, :928-941),
 * which for a functor with an analytic gradient is a thin wrapper over netlib's lmder (mode 1 = internal scaling,
 * factor 100, nprint 1; minimize() returns true iff MINPACK's info is 1, 2, 3 or 4).  VNL is a third-party
 * dependency that is absent from /root/reference and from this image (no version pinned by the reference's
 * CMakeLists.txt:26-70), so its published algorithm is restated here, step for step: trust region on the scaled
 * step, QR with column pivoting of the m x n Jacobian, the lmpar iteration for the Levenberg-Marquardt parameter,
 * the actual/predicted reduction ratio, and the four convergence tests in MINPACK's order.
 *
 * Used by oracle/vnl_shim (the reference's own sources compiled against it -> oracle/_ref) and by the plain-C port
 * oracle/lsqr_oracle.c.  Column-major storage as in the Fortran original: a(i,j) = a[i + lda * j].
 */
#ifndef LSQR_ORACLE_MINPACK_LM_H
#define LSQR_ORACLE_MINPACK_LM_H

#include <math.h>
#include <stdlib.h>

#define MPK_EPSMCH 2.220446049250313e-16   /* dpmpar(1) */
#define MPK_DWARF 2.2250738585072014e-308  /* dpmpar(2) */

/* fcn(user, m, n, x, fvec, fjac, ldfjac, iflag): iflag 1 -> fvec = f(x); iflag 2 -> fjac = J(x) (column-major).
 * A negative return value terminates (info = that value). */
typedef int (*mpk_fcn)(void* user, int m, int n, const double* x, double* fvec, double* fjac, int ldfjac, int iflag);

static double mpk_enorm(int n, const double* x) {
  const double rdwarf = 3.834e-20, rgiant = 1.304e19;
  double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0, agiant;
  int i;
  if (n <= 0) return 0.0;
  agiant = rgiant / (double)n;
  for (i = 0; i < n; i++) {
    const double xabs = fabs(x[i]);
    if (xabs > rdwarf && xabs < agiant) {
      s2 += xabs * xabs;                                   /* intermediate components */
    } else if (xabs <= rdwarf) {                           /* small components */
      if (xabs > x3max) { const double t = x3max / xabs; s3 = 1.0 + s3 * (t * t); x3max = xabs; }
      else if (xabs != 0.0) { const double t = xabs / x3max; s3 += t * t; }
    } else {                                               /* large components */
      if (xabs > x1max) { const double t = x1max / xabs; s1 = 1.0 + s1 * (t * t); x1max = xabs; }
      else { const double t = xabs / x1max; s1 += t * t; }
    }
  }
  if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
  if (s2 != 0.0) {
    if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
    return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
  }
  return x3max * sqrt(s3);
}

/* Householder QR with column pivoting: a P = Q R.  On return the strict upper triangle of a holds that of R, rdiag its
 * diagonal, the lower trapezoid the Householder vectors; acnorm the input column norms; ipvt the permutation. */
static void mpk_qrfac(int m, int n, double* a, int lda, int* ipvt, double* rdiag, double* acnorm, double* wa) {
  int i, j, k, minmn = m < n ? m : n;
  for (j = 0; j < n; j++) {
    acnorm[j] = mpk_enorm(m, a + (size_t)lda * j);
    rdiag[j] = acnorm[j];
    wa[j] = rdiag[j];
    ipvt[j] = j;
  }
  for (j = 0; j < minmn; j++) {
    double ajnorm;
    int kmax = j;
    for (k = j; k < n; k++) if (rdiag[k] > rdiag[kmax]) kmax = k;    /* bring the column of largest norm into the pivot position */
    if (kmax != j) {
      for (i = 0; i < m; i++) { const double t = a[i + (size_t)lda * j]; a[i + (size_t)lda * j] = a[i + (size_t)lda * kmax]; a[i + (size_t)lda * kmax] = t; }
      rdiag[kmax] = rdiag[j];
      wa[kmax] = wa[j];
      k = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = k;
    }
    ajnorm = mpk_enorm(m - j, a + j + (size_t)lda * j);
    if (ajnorm != 0.0) {
      if (a[j + (size_t)lda * j] < 0.0) ajnorm = -ajnorm;
      for (i = j; i < m; i++) a[i + (size_t)lda * j] /= ajnorm;
      a[j + (size_t)lda * j] += 1.0;
      for (k = j + 1; k < n; k++) {                              /* apply the transformation to the remaining columns, update the norms */
        double sum = 0.0, temp;
        for (i = j; i < m; i++) sum += a[i + (size_t)lda * j] * a[i + (size_t)lda * k];
        temp = sum / a[j + (size_t)lda * j];
        for (i = j; i < m; i++) a[i + (size_t)lda * k] -= temp * a[i + (size_t)lda * j];
        if (rdiag[k] != 0.0) {
          double d;
          temp = a[j + (size_t)lda * k] / rdiag[k];
          d = 1.0 - temp * temp;
          rdiag[k] *= sqrt(d > 0.0 ? d : 0.0);
          temp = rdiag[k] / wa[k];
          if (0.05 * (temp * temp) <= MPK_EPSMCH) {
            rdiag[k] = mpk_enorm(m - j - 1, a + (j + 1) + (size_t)lda * k);
            wa[k] = rdiag[k];
          }
        }
      }
    }
    rdiag[j] = -ajnorm;
  }
}

/* Solves min |a x - b|, |D x| appended, given the pivoted QR of a (R in r, upper triangle) and the first n components of Q^T b. */
static void mpk_qrsolv(int n, double* r, int ldr, const int* ipvt, const double* diag, const double* qtb, double* x, double* sdiag, double* wa) {
  int i, j, k, l, nsing;
  for (j = 0; j < n; j++) {
    for (i = j; i < n; i++) r[i + (size_t)ldr * j] = r[j + (size_t)ldr * i];
    x[j] = r[j + (size_t)ldr * j];
    wa[j] = qtb[j];
  }
  for (j = 0; j < n; j++) {                                       /* eliminate the diagonal matrix d with Givens rotations */
    l = ipvt[j];
    if (diag[l] != 0.0) {
      double qtbpj = 0.0;
      for (k = j; k < n; k++) sdiag[k] = 0.0;
      sdiag[j] = diag[l];
      for (k = j; k < n; k++) {
        double c, s, temp;
        if (sdiag[k] == 0.0) continue;
        if (fabs(r[k + (size_t)ldr * k]) < fabs(sdiag[k])) {
          const double cotan = r[k + (size_t)ldr * k] / sdiag[k];
          s = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
          c = s * cotan;
        } else {
          const double tn = sdiag[k] / r[k + (size_t)ldr * k];
          c = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
          s = c * tn;
        }
        r[k + (size_t)ldr * k] = c * r[k + (size_t)ldr * k] + s * sdiag[k];
        temp = c * wa[k] + s * qtbpj;
        qtbpj = -s * wa[k] + c * qtbpj;
        wa[k] = temp;
        for (i = k + 1; i < n; i++) {
          temp = c * r[i + (size_t)ldr * k] + s * sdiag[i];
          sdiag[i] = -s * r[i + (size_t)ldr * k] + c * sdiag[i];
          r[i + (size_t)ldr * k] = temp;
        }
      }
    }
    sdiag[j] = r[j + (size_t)ldr * j];
    r[j + (size_t)ldr * j] = x[j];
  }
  nsing = n;
  for (j = 0; j < n; j++) {
    if (sdiag[j] == 0.0 && nsing == n) nsing = j;
    if (nsing < n) wa[j] = 0.0;
  }
  for (k = 1; k <= nsing; k++) {
    double sum = 0.0;
    j = nsing - k;
    for (i = j + 1; i < nsing; i++) sum += r[i + (size_t)ldr * j] * wa[i];
    wa[j] = (wa[j] - sum) / sdiag[j];
  }
  for (j = 0; j < n; j++) x[ipvt[j]] = wa[j];
}

/* The Levenberg-Marquardt parameter par with | |D x| - delta | <= 0.1 delta (or par = 0 and |D x| <= 1.1 delta). */
static void mpk_lmpar(int n, double* r, int ldr, const int* ipvt, const double* diag, const double* qtb, double delta, double* par, double* x,
                      double* sdiag, double* wa1, double* wa2) {
  int i, j, k, l, nsing = n, iter = 0;
  double dxnorm, fp, gnorm, parl = 0.0, paru, temp;
  for (j = 0; j < n; j++) {                                        /* Gauss-Newton direction (least-squares solution if rank deficient) */
    wa1[j] = qtb[j];
    if (r[j + (size_t)ldr * j] == 0.0 && nsing == n) nsing = j;
    if (nsing < n) wa1[j] = 0.0;
  }
  for (k = 1; k <= nsing; k++) {
    j = nsing - k;
    wa1[j] /= r[j + (size_t)ldr * j];
    temp = wa1[j];
    for (i = 0; i < j; i++) wa1[i] -= r[i + (size_t)ldr * j] * temp;
  }
  for (j = 0; j < n; j++) x[ipvt[j]] = wa1[j];
  for (j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
  dxnorm = mpk_enorm(n, wa2);
  fp = dxnorm - delta;
  if (fp <= 0.1 * delta) { *par = 0.0; return; }
  if (nsing >= n) {                                                /* the Newton step gives a lower bound parl for the zero of the function */
    for (j = 0; j < n; j++) { l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
    for (j = 0; j < n; j++) {
      double sum = 0.0;
      for (i = 0; i < j; i++) sum += r[i + (size_t)ldr * j] * wa1[i];
      wa1[j] = (wa1[j] - sum) / r[j + (size_t)ldr * j];
    }
    temp = mpk_enorm(n, wa1);
    parl = ((fp / delta) / temp) / temp;
  }
  for (j = 0; j < n; j++) {                                        /* upper bound paru */
    double sum = 0.0;
    for (i = 0; i <= j; i++) sum += r[i + (size_t)ldr * j] * qtb[i];
    wa1[j] = sum / diag[ipvt[j]];
  }
  gnorm = mpk_enorm(n, wa1);
  paru = gnorm / delta;
  if (paru == 0.0) paru = MPK_DWARF / (delta < 0.1 ? delta : 0.1);
  if (*par < parl) *par = parl;
  if (*par > paru) *par = paru;
  if (*par == 0.0) *par = gnorm / dxnorm;
  for (;;) {
    double parc;
    iter++;
    if (*par == 0.0) { const double t = 0.001 * paru; *par = MPK_DWARF > t ? MPK_DWARF : t; }
    temp = sqrt(*par);
    for (j = 0; j < n; j++) wa1[j] = temp * diag[j];
    mpk_qrsolv(n, r, ldr, ipvt, wa1, qtb, x, sdiag, wa2);
    for (j = 0; j < n; j++) wa2[j] = diag[j] * x[j];
    dxnorm = mpk_enorm(n, wa2);
    temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
    for (j = 0; j < n; j++) { l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }     /* Newton correction */
    for (j = 0; j < n; j++) {
      wa1[j] /= sdiag[j];
      temp = wa1[j];
      for (i = j + 1; i < n; i++) wa1[i] -= r[i + (size_t)ldr * j] * temp;
    }
    temp = mpk_enorm(n, wa1);
    parc = ((fp / delta) / temp) / temp;
    if (fp > 0.0 && *par > parl) parl = *par;
    if (fp < 0.0 && *par < paru) paru = *par;
    temp = *par + parc;
    *par = parl > temp ? parl : temp;
  }
}

/* lmder, mode 1 (diag set internally from the column norms), nprint ignored.  Returns MINPACK's info:
 *   0 improper input; 1 ftol (actual and predicted relative reductions); 2 xtol (relative step); 3 both; 4 gtol (fvec
 *   orthogonal to the columns of the Jacobian); 5 maxfev reached; 6 / 7 / 8: ftol / xtol / gtol too small. */
static int mpk_lmder(mpk_fcn fcn, void* user, int m, int n, double* x, double* fvec, double ftol, double xtol, double gtol, int maxfev, double factor,
                     int* nfev_out, int* njev_out) {
  int i, j, l, info = 0, iflag, nfev = 0, njev = 0, iter = 1, done = 0;
  double actred, delta = 0, dirder, fnorm, fnorm1, gnorm, par = 0, pnorm, prered, ratio, temp, temp1, temp2, xnorm = 0;
  double *fjac, *diag, *qtf, *wa1, *wa2, *wa3, *wa4;
  int* ipvt;
  if (n <= 0 || m < n || ftol < 0.0 || xtol < 0.0 || gtol < 0.0 || maxfev <= 0 || factor <= 0.0) return 0;
  fjac = (double*)malloc(sizeof(double) * (size_t)m * n);
  diag = (double*)malloc(sizeof(double) * n * 5);
  wa4 = (double*)malloc(sizeof(double) * m);
  ipvt = (int*)malloc(sizeof(int) * n);
  if (!fjac || !diag || !wa4 || !ipvt) { free(fjac); free(diag); free(wa4); free(ipvt); return 0; }
  qtf = diag + n; wa1 = qtf + n; wa2 = wa1 + n; wa3 = wa2 + n;
  iflag = fcn(user, m, n, x, fvec, fjac, m, 1);
  nfev = 1;
  if (iflag < 0) { info = iflag; goto finish; }
  fnorm = mpk_enorm(m, fvec);
  while (!done) {                                                     /* outer loop */
    iflag = fcn(user, m, n, x, fvec, fjac, m, 2);
    njev++;
    if (iflag < 0) { info = iflag; break; }
    mpk_qrfac(m, n, fjac, m, ipvt, wa1, wa2, wa3);
    if (iter == 1) {                                                  /* scale according to the norms of the columns of the initial Jacobian */
      for (j = 0; j < n; j++) { diag[j] = wa2[j]; if (wa2[j] == 0.0) diag[j] = 1.0; }
      for (j = 0; j < n; j++) wa3[j] = diag[j] * x[j];
      xnorm = mpk_enorm(n, wa3);
      delta = factor * xnorm;
      if (delta == 0.0) delta = factor;
    }
    for (i = 0; i < m; i++) wa4[i] = fvec[i];                         /* qtf = first n components of Q^T fvec */
    for (j = 0; j < n; j++) {
      if (fjac[j + (size_t)m * j] != 0.0) {
        double sum = 0.0;
        for (i = j; i < m; i++) sum += fjac[i + (size_t)m * j] * wa4[i];
        temp = -sum / fjac[j + (size_t)m * j];
        for (i = j; i < m; i++) wa4[i] += fjac[i + (size_t)m * j] * temp;
      }
      fjac[j + (size_t)m * j] = wa1[j];
      qtf[j] = wa4[j];
    }
    gnorm = 0.0;                                                      /* norm of the scaled gradient */
    if (fnorm != 0.0) {
      for (j = 0; j < n; j++) {
        l = ipvt[j];
        if (wa2[l] != 0.0) {
          double sum = 0.0;
          for (i = 0; i <= j; i++) sum += fjac[i + (size_t)m * j] * (qtf[i] / fnorm);
          temp = fabs(sum / wa2[l]);
          if (temp > gnorm) gnorm = temp;
        }
      }
    }
    if (gnorm <= gtol) { info = 4; break; }
    for (j = 0; j < n; j++) if (wa2[j] > diag[j]) diag[j] = wa2[j];   /* rescale */
    for (;;) {                                                        /* inner loop */
      mpk_lmpar(n, fjac, m, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3, wa4);
      for (j = 0; j < n; j++) { wa1[j] = -wa1[j]; wa2[j] = x[j] + wa1[j]; wa3[j] = diag[j] * wa1[j]; }
      pnorm = mpk_enorm(n, wa3);
      if (iter == 1 && pnorm < delta) delta = pnorm;                  /* on the first iteration, adjust the initial step bound */
      iflag = fcn(user, m, n, wa2, wa4, fjac, m, 1);
      nfev++;
      if (iflag < 0) { info = iflag; done = 1; break; }
      fnorm1 = mpk_enorm(m, wa4);
      actred = -1.0;                                                  /* scaled actual reduction */
      if (0.1 * fnorm1 < fnorm) { temp = fnorm1 / fnorm; actred = 1.0 - temp * temp; }
      for (j = 0; j < n; j++) {                                       /* scaled predicted reduction and directional derivative */
        wa3[j] = 0.0;
        l = ipvt[j];
        temp = wa1[l];
        for (i = 0; i <= j; i++) wa3[i] += fjac[i + (size_t)m * j] * temp;
      }
      temp1 = mpk_enorm(n, wa3) / fnorm;
      temp2 = (sqrt(par) * pnorm) / fnorm;
      prered = temp1 * temp1 + temp2 * temp2 / 0.5;
      dirder = -(temp1 * temp1 + temp2 * temp2);
      ratio = 0.0;
      if (prered != 0.0) ratio = actred / prered;
      if (ratio <= 0.25) {                                            /* update the step bound */
        if (actred >= 0.0) temp = 0.5; else temp = 0.5 * dirder / (dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
        delta = temp * (delta < pnorm / 0.1 ? delta : pnorm / 0.1);
        par /= temp;
      } else if (par == 0.0 || ratio >= 0.75) {
        delta = pnorm / 0.5;
        par *= 0.5;
      }
      if (ratio >= 1e-4) {                                            /* successful iteration: update x, fvec and their norms */
        for (j = 0; j < n; j++) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
        for (i = 0; i < m; i++) fvec[i] = wa4[i];
        xnorm = mpk_enorm(n, wa2);
        fnorm = fnorm1;
        iter++;
      }
      if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) info = 1;      /* tests for convergence */
      if (delta <= xtol * xnorm) info = 2;
      if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0 && info == 2) info = 3;
      if (info != 0) { done = 1; break; }
      if (nfev >= maxfev) info = 5;                                                   /* tests for termination and stringent tolerances */
      if (fabs(actred) <= MPK_EPSMCH && prered <= MPK_EPSMCH && 0.5 * ratio <= 1.0) info = 6;
      if (delta <= MPK_EPSMCH * xnorm) info = 7;
      if (gnorm <= MPK_EPSMCH) info = 8;
      if (info != 0) { done = 1; break; }
      if (ratio >= 1e-4) break;                                       /* end of the inner loop: repeat if the iteration was unsuccessful */
    }
  }
finish:
  if (nfev_out) *nfev_out = nfev;
  if (njev_out) *njev_out = njev;
  free(fjac); free(diag); free(wa4); free(ipvt);
  return info;
}

#endif /* LSQR_ORACLE_MINPACK_LM_H */
