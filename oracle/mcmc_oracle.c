/*
 * mcmc_oracle.c -- CPU restatement of mcmcf90's sampling hot path (see mcmc_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 * PARITY: NOT pinned to a run of the Fortran reference (no Fortran compiler here or on the GPU box --
 * profiles/r02_probe_fortran.txt -- and the reference ships no expected outputs).  Pinned instead by
 * (1) an independent second restatement of the Fortran (oracle/restate_np.py, real BLAS/LAPACK through scipy):
 * identical chain indices / counters / draws consumed and values to 1e-9 on every sampler
 * (tests/test_ref_parity.py), (2) the reference's one binary fixture testcases/data.mat, (3) analytic known
 * answers, scipy LAPACK cross-checks and identities (tests/test_oracle_pins.py).  oracle/Makefile.ref builds the
 * reference itself against the injected stream wherever gfortran exists.
 *
 * All citations are file:line into /root/reference.  Arrays are column-major with
 * 1-based Fortran indices mapped through the IDX macro so that loops read like the
 * reference.  Build: gcc -O2 -ffp-contract=off (the reference's flags, linux64.mk:38,
 * are -O2 -mtune=native: no FMA contraction).
 */
#include "mcmc_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

#define IDX(i, j, ld) ((size_t)((j)-1) * (size_t)(ld) + (size_t)((i)-1))

/* log(tiny(0d0)), mcmcprec.F90:34-41 */
static const double LOG_REALMIN = -708.3964185322641;

typedef struct {
  int mode; /* 0 injected, 1 philox */
  const double* inj;
  long ninj, pos;
  uint64_t seed, chain, ndrawn;
  int exhausted;
  int saved; /* polar spare, mcmcrand.F90:172-173 */
  double saved_y;
} orc_rng;

typedef struct {
  int id, npar, ny;
  const double* blob;
  long blob_len;
  /* default Gaussian prior (priorfun.f90:97-100); NULL => flat */
  double *pmu, *psig;
} orc_model;

struct orc_chain {
  orc_cfg cfg;
  int npar, nycol, ncolchain;
  orc_model model;
  double *par0, *cmat0, *sigma2, *oldpar;
  int* nobs;
  double *R, *R2, *iC, *qcovstd, *chaincmat, *chainmean;
  double chainwsum;
  double *chain, *sschain, *s2chain;
  int stayed, bndstayed, draccepted, drtries, chainind, simuind;
  int erstayed; /* mcmc.F90:49 */
  /* saved locals of MCMC_adapt, MCMC_adapt.F90:19 */
  int istart, istartind, lastind, lastfreq, newfreq;
  orc_rng rng;
  int status;
  double S02;
  /* loop-carried locals of MCMC_run / MCMC_run_ram / MCMC_run_scam, kept here so that a run can be
   * advanced in pieces (orc_advance) -- the pooled-adaptation tests stop every chain at the ticks */
  double *ss1, sspri1, alpha12;
  int started, next_i;
  int pool; /* 1: MCMC_adapt updates chaincmat/chainmean only; R comes from orc_factor_from_cov */
};

/* ------------------------------------------------------------------ Philox */
/* Philox4x32-10 (Salmon et al. 2011, Random123); replaces the compiler's
 * random_number (mcmcrand.F90:55,104,138,156,177; MCMC_DRAM.F90:132,151). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* uniform number k of stream (seed, chain): block k>>1, half k&1, 53-bit mantissa in
 * [0,1) -- same support as gfortran's random_number. */
double orc_philox_uniform(uint64_t seed, uint64_t chain, uint64_t k) {
  uint64_t blk = k >> 1;
  uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)chain, (uint32_t)(chain >> 32)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t o[4];
  orc_philox4x32_10(ctr, key, o);
  uint64_t bits = (k & 1) ? ((uint64_t)o[2] | ((uint64_t)o[3] << 32)) : ((uint64_t)o[0] | ((uint64_t)o[1] << 32));
  return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}

static double rng_uniform(orc_rng* g) {
  double u;
  if (g->mode == 0) {
    if (g->pos >= g->ninj) {
      g->exhausted = 1;
      return 0.5;
    }
    u = g->inj[g->pos++];
  } else {
    u = orc_philox_uniform(g->seed, g->chain, g->ndrawn);
  }
  g->ndrawn++;
  return u;
}

/* mcmcrand.F90:166-190 normal_bm: Marsaglia polar; returns z*x(2) first, saves z*x(1) */
static double normal_bm(orc_rng* g) {
  if (!g->saved) {
    double x1, x2, xx;
    for (;;) {
      x1 = rng_uniform(g); /* random_number(x) fills x(1) then x(2) */
      x2 = rng_uniform(g);
      x1 = 2.0 * x1 - 1.0;
      x2 = 2.0 * x2 - 1.0;
      xx = x1 * x1 + x2 * x2;
      if (xx < 1.0 && xx != 0.0) break;
      if (g->exhausted) break;
    }
    double z = sqrt(-2.0 * log(xx) / xx);
    g->saved_y = z * x1;
    g->saved = 1;
    return z * x2;
  }
  g->saved = 0;
  return g->saved_y;
}

/* mcmcrand.F90:120-162 gammar_mt (Marsaglia-Tsang 2000) */
static double gammar_mt(orc_rng* g, double a, double b) {
  double aa = a, bb = b;
  if (aa < 1.0) { /* mcmcrand.F90:136-146 (not reached from random_gamma) */
    double u = rng_uniform(g);
    bb = bb * pow(u, 1.0 / aa);
    aa = aa + 1.0;
  }
  double d = aa - 1.0 / 3.0;
  double c = 1.0 / sqrt(9.0 * d);
  double x, v, u;
  for (;;) {
    for (;;) {
      x = normal_bm(g);
      v = 1.0 + c * x;
      if (v > 0.0) break;
      if (g->exhausted) { v = 1.0; break; }
    }
    v = v * v * v;
    u = rng_uniform(g);
    double x2 = x * x;
    if (u < 1.0 - 0.0331 * (x2 * x2)) break;
    if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) break;
    if (g->exhausted) break;
  }
  return bb * d * v;
}

/* mcmcrand.F90:86-111 random_gamma(1,a,b) */
static double random_gamma1(orc_rng* g, double a, double b) {
  if (a < 1.0) {
    double u = rng_uniform(g);
    return gammar_mt(g, 1.0 + a, b) * pow(u, 1.0 / a);
  }
  return gammar_mt(g, a, b);
}

void orc_normals(orc_chain* ch, int n, double* out) { /* mcmcrand.F90:60-83 */
  for (int i = 0; i < n; i++) out[i] = normal_bm(&ch->rng) * 1.0 + 0.0;
}
double orc_gamma(orc_chain* ch, double a, double b) { return random_gamma1(&ch->rng, a, b); }

/* ------------------------------------------------------------- BLAS / LAPACK */
/* netlib reference BLAS dtrmv('U','T','N'), as called at matutils.F90:108-109 */
void orc_dtrmv_ut(int n, const double* A, int lda, double* x) {
  for (int j = n; j >= 1; j--) {
    double temp = x[j - 1];
    temp = temp * A[IDX(j, j, lda)];
    for (int i = j - 1; i >= 1; i--) temp = temp + A[IDX(i, j, lda)] * x[i - 1];
    x[j - 1] = temp;
  }
}

/* netlib dtrmv('U','N','N') on the leading k x k block (used by dtrti2) */
static void dtrmv_un(int k, const double* A, int lda, double* x) {
  for (int j = 1; j <= k; j++) {
    if (x[j - 1] != 0.0) {
      double temp = x[j - 1];
      for (int i = 1; i <= j - 1; i++) x[i - 1] = x[i - 1] + temp * A[IDX(i, j, lda)];
      x[j - 1] = x[j - 1] * A[IDX(j, j, lda)];
    }
  }
}

/* netlib dgemv, alpha=1, beta=0, square; matutils.F90:161 */
void orc_dgemv(char trans, int n, const double* A, int lda, const double* x, double* y) {
  for (int i = 0; i < n; i++) y[i] = 0.0;
  if (trans == 'N' || trans == 'n') {
    for (int j = 1; j <= n; j++) {
      double temp = x[j - 1];
      for (int i = 1; i <= n; i++) y[i - 1] = y[i - 1] + temp * A[IDX(i, j, lda)];
    }
  } else {
    for (int j = 1; j <= n; j++) {
      double temp = 0.0;
      for (int i = 1; i <= n; i++) temp = temp + A[IDX(i, j, lda)] * x[i - 1];
      y[j - 1] = y[j - 1] + temp;
    }
  }
}

/* netlib dsymv('U'), alpha=1, beta=0; matutils.F90:180 */
void orc_dsymv_u(int n, const double* A, int lda, const double* x, double* y) {
  for (int i = 0; i < n; i++) y[i] = 0.0;
  for (int j = 1; j <= n; j++) {
    double temp1 = x[j - 1], temp2 = 0.0;
    for (int i = 1; i <= j - 1; i++) {
      y[i - 1] = y[i - 1] + temp1 * A[IDX(i, j, lda)];
      temp2 = temp2 + A[IDX(i, j, lda)] * x[i - 1];
    }
    y[j - 1] = y[j - 1] + temp1 * A[IDX(j, j, lda)] + temp2;
  }
}

static double ddot_s(int n, const double* x, int incx, const double* y, int incy) {
  double t = 0.0;
  for (int i = 0; i < n; i++) t = t + x[(size_t)i * incx] * y[(size_t)i * incy];
  return t;
}

/* classic netlib dnrm2 (scaled ssq recurrence), used by dchdd.f:149 */
static double dnrm2_c(int n, const double* x) {
  if (n < 1) return 0.0;
  if (n == 1) return fabs(x[0]);
  double scale = 0.0, ssq = 1.0;
  for (int i = 0; i < n; i++) {
    if (x[i] != 0.0) {
      double absxi = fabs(x[i]);
      if (scale < absxi) {
        double t = scale / absxi;
        ssq = 1.0 + ssq * t * t;
        scale = absxi;
      } else {
        double t = absxi / scale;
        ssq = ssq + t * t;
      }
    }
  }
  return scale * sqrt(ssq);
}

/* unblocked LAPACK dpotf2('U'): the order dpotrf uses for n < 64 (matutils.F90:363) */
int orc_dpotf2_u(int n, double* A, int lda) {
  for (int j = 1; j <= n; j++) {
    double ajj = A[IDX(j, j, lda)] - ddot_s(j - 1, &A[IDX(1, j, lda)], 1, &A[IDX(1, j, lda)], 1);
    if (ajj <= 0.0 || ajj != ajj) {
      A[IDX(j, j, lda)] = ajj;
      return j;
    }
    ajj = sqrt(ajj);
    A[IDX(j, j, lda)] = ajj;
    if (j < n) {
      for (int k = j + 1; k <= n; k++) { /* dgemv('T', j-1, n-j, -1, A(1,j+1), lda, A(1,j), 1, 1, A(j,j+1), lda) */
        double temp = 0.0;
        for (int i = 1; i <= j - 1; i++) temp = temp + A[IDX(i, k, lda)] * A[IDX(i, j, lda)];
        A[IDX(j, k, lda)] = A[IDX(j, k, lda)] + (-1.0) * temp;
      }
      double rajj = 1.0 / ajj; /* dscal(n-j, 1/ajj, A(j,j+1), lda) */
      for (int k = j + 1; k <= n; k++) A[IDX(j, k, lda)] = rajj * A[IDX(j, k, lda)];
    }
  }
  return 0;
}

/* dpotri('U') = dtrtri (dtrti2 order) + dlauum (dlauu2 order); MCMC_adapt.F90:219 */
int orc_dpotri_u(int n, double* A, int lda) {
  for (int i = 1; i <= n; i++)
    if (A[IDX(i, i, lda)] == 0.0) return i;
  for (int j = 1; j <= n; j++) { /* dtrti2 */
    A[IDX(j, j, lda)] = 1.0 / A[IDX(j, j, lda)];
    double ajj = -A[IDX(j, j, lda)];
    dtrmv_un(j - 1, A, lda, &A[IDX(1, j, lda)]);
    for (int i = 1; i <= j - 1; i++) A[IDX(i, j, lda)] = ajj * A[IDX(i, j, lda)];
  }
  for (int i = 1; i <= n; i++) { /* dlauu2 */
    double aii = A[IDX(i, i, lda)];
    if (i < n) {
      A[IDX(i, i, lda)] = ddot_s(n - i + 1, &A[IDX(i, i, lda)], lda, &A[IDX(i, i, lda)], lda);
      /* dgemv('N', i-1, n-i, 1, A(1,i+1), lda, A(i,i+1), lda, aii, A(1,i), 1) */
      for (int r = 1; r <= i - 1; r++) A[IDX(r, i, lda)] = aii * A[IDX(r, i, lda)];
      for (int k = i + 1; k <= n; k++) {
        double temp = A[IDX(i, k, lda)];
        for (int r = 1; r <= i - 1; r++) A[IDX(r, i, lda)] = A[IDX(r, i, lda)] + temp * A[IDX(r, k, lda)];
      }
    } else {
      for (int r = 1; r <= i; r++) A[IDX(r, i, lda)] = aii * A[IDX(r, i, lda)];
    }
  }
  return 0;
}

/* classic netlib drotg (called at dchud.f:138) */
void orc_drotg(double* da, double* db, double* c, double* s) {
  double roe = *db, r, z;
  if (fabs(*da) > fabs(*db)) roe = *da;
  double scale = fabs(*da) + fabs(*db);
  if (scale == 0.0) {
    *c = 1.0; *s = 0.0; r = 0.0; z = 0.0;
  } else {
    double ta = *da / scale, tb = *db / scale;
    r = scale * sqrt(ta * ta + tb * tb);
    r = (roe < 0.0 ? -1.0 : 1.0) * r; /* dsign(1,roe)*r */
    *c = *da / r;
    *s = *db / r;
    z = 1.0;
    if (fabs(*da) > fabs(*db)) z = *s;
    if (fabs(*db) >= fabs(*da) && *c != 0.0) z = 1.0 / *c;
  }
  *da = r;
  *db = z;
}

/* dchud.f:122-139 (R part only; nz=0 at matutils.F90:680-682) */
void orc_dchud(double* r, int ldr, int p, const double* x, double* c, double* s) {
  for (int j = 1; j <= p; j++) {
    double xj = x[j - 1];
    for (int i = 1; i <= j - 1; i++) {
      double t = c[i - 1] * r[IDX(i, j, ldr)] + s[i - 1] * xj;
      xj = c[i - 1] * xj - s[i - 1] * r[IDX(i, j, ldr)];
      r[IDX(i, j, ldr)] = t;
    }
    orc_drotg(&r[IDX(j, j, ldr)], &xj, &c[j - 1], &s[j - 1]);
  }
}

/* dchdd.f:141-179 (R part only) */
int orc_dchdd(double* r, int ldr, int p, const double* x, double* c, double* s) {
  s[0] = x[0] / r[IDX(1, 1, ldr)];
  for (int j = 2; j <= p; j++) {
    s[j - 1] = x[j - 1] - ddot_s(j - 1, &r[IDX(1, j, ldr)], 1, s, 1);
    s[j - 1] = s[j - 1] / r[IDX(j, j, ldr)];
  }
  double norm = dnrm2_c(p, s);
  if (!(norm < 1.0)) return -1;
  double alpha = sqrt(1.0 - norm * norm);
  for (int ii = 1; ii <= p; ii++) {
    int i = p - ii + 1;
    double scale = alpha + fabs(s[i - 1]);
    double a = alpha / scale;
    double b = s[i - 1] / scale;
    norm = sqrt(a * a + b * b);
    c[i - 1] = a / norm;
    s[i - 1] = b / norm;
    alpha = scale * norm;
  }
  for (int j = 1; j <= p; j++) {
    double xx = 0.0;
    for (int ii = 1; ii <= j; ii++) {
      int i = j - ii + 1;
      double t = c[i - 1] * xx + s[i - 1] * r[IDX(i, j, ldr)];
      r[IDX(i, j, ldr)] = c[i - 1] * r[IDX(i, j, ldr)] - s[i - 1] * xx;
      xx = t;
    }
  }
  return 0;
}

/* matutils.F90:232-341 covmat.  x is n x p (leading dim ldx), w has nw entries
 * (nw==n: per-row weights; nw==1: scalar weight; nw==0: absent). */
void orc_covmat(const double* x, int n, int ldx, int p, double* cmat, const double* w, int nw,
                double* xmean, double* wsum, int update) {
  double w2, wsum2;
  double* xmean2 = (double*)malloc(sizeof(double) * (size_t)p);
  if (nw == n && nw > 0) { /* matutils.F90:254-256 (size(w)==n is tested first) */
    w2 = -1.0;
    wsum2 = 0.0;
    for (int i = 0; i < n; i++) wsum2 = wsum2 + w[i];
  } else if (nw == 1) {
    w2 = w[0];
    wsum2 = (double)n * w2;
  } else {
    w2 = 1.0;
    wsum2 = (double)n;
  }
  int doupdate = (update && wsum != NULL && *wsum > 0.0);
  if (doupdate) { /* matutils.F90:283-310 */
    for (int i = 1; i <= n; i++) {
      for (int k = 1; k <= p; k++) xmean2[k - 1] = x[IDX(i, k, ldx)] - xmean[k - 1];
      double w3 = (w2 == -1.0) ? w[i - 1] : w2;
      double f1 = w3 / (*wsum + w3 - 1.0);
      double f2 = *wsum / (*wsum + w3);
      for (int b = 1; b <= p; b++)
        for (int a = 1; a <= p; a++)
          cmat[IDX(a, b, p)] = cmat[IDX(a, b, p)] + f1 * (f2 * (xmean2[a - 1] * xmean2[b - 1]) - cmat[IDX(a, b, p)]);
      double f3 = w3 / (*wsum + w3);
      for (int k = 1; k <= p; k++) xmean[k - 1] = xmean[k - 1] + f3 * xmean2[k - 1];
      *wsum = w3 + *wsum;
    }
  } else { /* matutils.F90:312-337 */
    for (int k = 1; k <= p; k++) {
      double acc = 0.0;
      for (int i = 1; i <= n; i++) acc = acc + x[IDX(i, k, ldx)] * ((w2 == -1.0) ? w[i - 1] : w2);
      xmean2[k - 1] = acc / wsum2;
    }
    for (int a = 1; a <= p; a++)
      for (int b = 1; b <= a; b++) {
        double acc = 0.0;
        for (int i = 1; i <= n; i++)
          acc = acc + (x[IDX(i, a, ldx)] - xmean2[a - 1]) *
                          ((x[IDX(i, b, ldx)] - xmean2[b - 1]) * ((w2 == -1.0) ? w[i - 1] : w2));
        cmat[IDX(a, b, p)] = acc / (wsum2 - 1.0);
        if (a != b) cmat[IDX(b, a, p)] = cmat[IDX(a, b, p)];
      }
    if (xmean) memcpy(xmean, xmean2, sizeof(double) * (size_t)p);
    if (wsum) *wsum = wsum2;
  }
  free(xmean2);
}

/* Replacement for dgesvd('A','N') on a symmetric PSD matrix (matutils.F90:409,615):
 * cyclic two-sided Jacobi; eigenvalues sorted descending; each column's sign fixed so
 * that its largest-magnitude component is positive (dgesvd's sign is arbitrary).
 * A, U are n x n column-major. Returns 0, or 1 if not converged in 60 sweeps. */
int orc_symeig(int n, const double* Ain, double* U, double* s) {
  double* A = (double*)malloc(sizeof(double) * (size_t)n * n);
  memcpy(A, Ain, sizeof(double) * (size_t)n * n);
  for (int j = 1; j <= n; j++)
    for (int i = 1; i <= n; i++) {
      U[IDX(i, j, n)] = (i == j) ? 1.0 : 0.0;
      if (i > j) A[IDX(i, j, n)] = A[IDX(j, i, n)]; /* upper triangle is authoritative */
    }
  int conv = 0;
  for (int sweep = 0; sweep < 60 && !conv; sweep++) {
    double off = 0.0, dia = 0.0;
    for (int j = 1; j <= n; j++) {
      dia += A[IDX(j, j, n)] * A[IDX(j, j, n)];
      for (int i = 1; i < j; i++) off += A[IDX(i, j, n)] * A[IDX(i, j, n)];
    }
    if (off <= 1e-32 * dia || off == 0.0) { conv = 1; break; }
    for (int p = 1; p <= n - 1; p++)
      for (int q = p + 1; q <= n; q++) {
        double apq = A[IDX(p, q, n)];
        if (apq == 0.0) continue;
        double app = A[IDX(p, p, n)], aqq = A[IDX(q, q, n)];
        double theta = (aqq - app) / (2.0 * apq);
        double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 1; k <= n; k++) { /* A <- A J */
          double akp = A[IDX(k, p, n)], akq = A[IDX(k, q, n)];
          A[IDX(k, p, n)] = c * akp - sn * akq;
          A[IDX(k, q, n)] = sn * akp + c * akq;
        }
        for (int k = 1; k <= n; k++) { /* A <- J' A */
          double apk = A[IDX(p, k, n)], aqk = A[IDX(q, k, n)];
          A[IDX(p, k, n)] = c * apk - sn * aqk;
          A[IDX(q, k, n)] = sn * apk + c * aqk;
        }
        for (int k = 1; k <= n; k++) {
          double ukp = U[IDX(k, p, n)], ukq = U[IDX(k, q, n)];
          U[IDX(k, p, n)] = c * ukp - sn * ukq;
          U[IDX(k, q, n)] = sn * ukp + c * ukq;
        }
      }
  }
  for (int j = 1; j <= n; j++) s[j - 1] = fabs(A[IDX(j, j, n)]);
  /* selection sort, descending (stable for ties) */
  for (int a = 1; a <= n - 1; a++) {
    int best = a;
    for (int b = a + 1; b <= n; b++)
      if (s[b - 1] > s[best - 1]) best = b;
    if (best != a) {
      double ts = s[a - 1]; s[a - 1] = s[best - 1]; s[best - 1] = ts;
      for (int k = 1; k <= n; k++) {
        double tu = U[IDX(k, a, n)]; U[IDX(k, a, n)] = U[IDX(k, best, n)]; U[IDX(k, best, n)] = tu;
      }
    }
  }
  for (int j = 1; j <= n; j++) {
    int im = 1;
    for (int i = 2; i <= n; i++)
      if (fabs(U[IDX(i, j, n)]) > fabs(U[IDX(im, j, n)])) im = i;
    if (U[IDX(im, j, n)] < 0.0)
      for (int i = 1; i <= n; i++) U[IDX(i, j, n)] = -U[IDX(i, j, n)];
  }
  free(A);
  return conv ? 0 : 1;
}

/* ------------------------------------------------------------ user models */
/* blob layouts (shared with include/mcmcb200_model.cuh):
 *  EXPREG: [n, 0, x[npad], y[npad]], npad = n rounded up to even
 *  GAUSS : [d, 0, mu[dpad], Lam[d*d] row-major]
 *  BANANA: [d, b]
 *  HIER  : [G, J, y[J][G] observation-major] params (theta_1..G, mu, log tau)         */
static void model_ss(const orc_model* m, const double* theta, double* ss) {
  const double* b = m->blob;
  switch (m->id) {
    case ORC_MODEL_EXPREG: { /* testcases/mcmcrun.F90:89,104 */
      long n = (long)b[0], npad = (n + 1) & ~1L;
      const double *x = b + 2, *y = b + 2 + npad;
      double acc = 0.0;
      for (long i = 0; i < n; i++) {
        double r = y[i] - theta[0] * exp(-theta[1] * x[i]);
        acc = acc + r * r;
      }
      ss[0] = acc;
      break;
    }
    case ORC_MODEL_GAUSS: { /* testcases/mcmcrun4.F90:47 dot_product(matmul(lam,theta-mu),theta-mu) */
      int d = (int)b[0], dpad = (d + 1) & ~1;
      const double *mu = b + 2, *lam = b + 2 + dpad;
      double acc = 0.0;
      for (int i = 0; i < d; i++) {
        double row = 0.0;
        for (int j = 0; j < d; j++) row = row + lam[(size_t)i * d + j] * (theta[j] - mu[j]);
        acc = acc + row * (theta[i] - mu[i]);
      }
      ss[0] = acc;
      break;
    }
    case ORC_MODEL_BANANA: { /* build-defined target, SURVEY.md 8d (C4) */
      int d = (int)b[0];
      double bb = b[1];
      double p1 = theta[0], p2 = theta[1] + bb * theta[0] * theta[0] - 100.0 * bb;
      double acc = p1 * p1 / 100.0 + p2 * p2;
      for (int i = 2; i < d; i++) acc = acc + theta[i] * theta[i];
      ss[0] = acc;
      break;
    }
    case ORC_MODEL_HIER: { /* build-defined target, SURVEY.md 8d (C5) */
      int G = (int)b[0], J = (int)b[1];
      const double* y = b + 2;
      double mu = theta[G], ltau = theta[G + 1];
      double itau2 = exp(-2.0 * ltau);
      double acc = 0.0;
      for (int g = 0; g < G; g++) {
        double a = 0.0;
        for (int j = 0; j < J; j++) {
          double r = y[(size_t)j * G + g] - theta[g]; /* blob: y[J][G] */
          a = a + r * r;
        }
        double dm = theta[g] - mu;
        acc = acc + (a + dm * dm * itau2);
      }
      /* 2 G log tau (normalisation of N(mu,tau^2)) + weak hyperpriors mu~N(0,10^2), log tau~N(0,2^2) */
      acc = acc + 2.0 * G * ltau + mu * mu / 100.0 + ltau * ltau / 4.0;
      ss[0] = acc;
      break;
    }
    default:
      ss[0] = 0.0;
  }
}

static int model_checkbounds(const orc_model* m, const double* theta) {
  if (m->id == ORC_MODEL_EXPREG) { /* testcases/mcmcrun.F90:112-122: any(theta<=0) -> false */
    for (int i = 0; i < m->npar; i++)
      if (theta[i] <= 0.0) return 0;
  }
  return 1; /* checkbounds0.f90:3-18 */
}

static double model_priorfun(const orc_model* m, const double* theta) { /* priorfun.f90:31-103 */
  if (!m->pmu) return 0.0;
  double p = 0.0;
  for (int i = 0; i < m->npar; i++)
    if (m->psig[i] > 0.0) {
      double t = (theta[i] - m->pmu[i]) / m->psig[i];
      p = p + t * t;
    }
  return p;
}

double orc_model_ss(int model_id, const double* blob, const double* theta, int npar) {
  orc_model m;
  memset(&m, 0, sizeof m);
  m.id = model_id; m.npar = npar; m.ny = 1; m.blob = blob;
  double ss[1];
  model_ss(&m, theta, ss);
  return ss[0];
}

/* ------------------------------------------------------------------ config */
void orc_default_cfg(orc_cfg* c) { /* mcmcinit.F90:184-230 */
  memset(c, 0, sizeof *c);
  c->method = ORC_DRAM;
  c->nsimu = 0; c->doadapt = 1; c->doburnin = 0; c->burnintime = 0; c->badaptint = -1;
  c->greedy = 0; c->scalelimit = 0.05; c->scalefactor = 2.5; c->drscale = 0.0;
  c->adaptint = 100; c->adapthist = 0; c->adaptend = 0; c->initcmatn = 0;
  c->N0 = 1.0; c->S02 = 0.0; c->updatesigma = 1; c->condmax = 0.0;
  c->alphatarget = 0.234; c->nuparam = 0.7;
}

void orc_check_params(orc_cfg* c) { /* mcmcinit.F90:235-368 */
  if (c->adapthist < 0) c->adapthist = 0;
  if (c->adaptint < 0) { c->adaptint = 0; c->doadapt = 0; }
  if (c->burnintime < 0) c->burnintime = 0;
  if (c->badaptint <= 0) c->badaptint = c->adaptint;
  if (c->badaptint == 0) c->doburnin = 0;
  if (c->initcmatn < 0) c->initcmatn = 0;
  if (c->scalefactor < 0.0) c->scalefactor = 1.0;
  if (c->method == ORC_SCAM) { /* 324-333 */
    c->doscam = 1;
    if (c->condmax <= 0.0) c->condmax = 1.0e15;
    c->doburnin = 0;
    c->drscale = 0.0;
  } else {
    c->doscam = 0;
  }
  if (c->method == ORC_RAM) c->drscale = 0.0; /* 336-338 */
  c->dodr = (c->drscale > 0.0);               /* 341-346 */
  c->usesvd = (c->condmax > 0.0);             /* 348-353 */
}

/* ------------------------------------------------------ MH kernels (L3) */
/* matutils.F90:378-453 / 583-653: SVD square root with condmax floor.
 * scam==0: R = U*diag(sqrt(s)) ; scam==1: R = U, std = sqrt(s).  info=-1 if floored. */
static int svd_factor(orc_chain* ch, const double* cmat, double* R, double* std, int scam) {
  int n = ch->npar;
  double* s = (double*)malloc(sizeof(double) * (size_t)n);
  int info2 = orc_symeig(n, cmat, R, s);
  if (info2) ch->status |= ORC_ST_SVDFAIL;
  if (s[0] == 0.0) { free(s); return n; }
  double tol = s[0] / ch->cfg.condmax;
  if (s[n - 1] <= tol) {
    for (int i = 0; i < n; i++)
      if (s[i] < tol) s[i] = tol;
    info2 = -1;
  }
  if (scam) {
    for (int i = 0; i < n; i++) std[i] = sqrt(s[i]);
  } else {
    for (int i = 1; i <= n; i++) {
      double f = sqrt(s[i - 1]); /* dscal(n, sqrt(s(i)), R(1:n,i), 1), matutils.F90:442 */
      for (int k = 1; k <= n; k++) R[IDX(k, i, n)] = f * R[IDX(k, i, n)];
    }
  }
  free(s);
  return info2;
}

/* MCMC_adapt.F90:181-230 */
static int calculate_R(orc_chain* ch, double* cmat) {
  int n = ch->npar, info = 0;
  size_t nn = (size_t)n * n;
  double* R0 = (double*)malloc(sizeof(double) * nn);
  if (ch->cfg.doscam) {
    info = svd_factor(ch, cmat, R0, ch->qcovstd, 1);
    if (info == -1) info = 0;
    if (info == 0) memcpy(ch->R, R0, sizeof(double) * nn); /* s(1)==0 returns early with R untouched... */
    else memcpy(ch->R, R0, sizeof(double) * nn);           /* ...but MCMC_adapt.F90:197 still does R = R0 */
  } else {
    if (ch->cfg.usesvd) {
      info = svd_factor(ch, cmat, R0, NULL, 0);
      if (info == -1) { /* cmat = matmul(R0,transpose(R0)), MCMC_adapt.F90:205-208 */
        for (int j = 1; j <= n; j++)
          for (int i = 1; i <= n; i++) {
            double acc = 0.0;
            for (int k = 1; k <= n; k++) acc = acc + R0[IDX(i, k, n)] * R0[IDX(j, k, n)];
            cmat[IDX(i, j, n)] = acc;
          }
        info = 0;
      }
    } else { /* covtor, matutils.F90:345-374 */
      memcpy(R0, cmat, sizeof(double) * nn);
      info = orc_dpotf2_u(n, R0, n);
    }
    if (info != 0) {
      ch->status |= ORC_ST_CHOLFAIL;
    } else {
      double sq = sqrt((double)n);
      for (size_t k = 0; k < nn; k++) ch->R[k] = R0[k] * 2.4 / sq; /* MCMC_adapt.F90:216 */
      if (ch->cfg.dodr) {
        memcpy(ch->iC, ch->R, sizeof(double) * nn);
        int i2 = orc_dpotri_u(n, ch->iC, n);
        if (i2 != 0) ch->status |= ORC_ST_CHOLFAIL; /* reference stops, MCMC_adapt.F90:220-223 */
        for (size_t k = 0; k < nn; k++) ch->R2[k] = ch->R[k] / ch->cfg.drscale;
      }
    }
  }
  free(R0);
  return info;
}

/* MCMC_DRAM.F90:20-31 */
static void propose(orc_chain* ch, const double* oldpar, const double* R, double* newpar, double* zout) {
  int n = ch->npar;
  double* z = (double*)malloc(sizeof(double) * (size_t)n);
  double* p = (double*)malloc(sizeof(double) * (size_t)n);
  orc_normals(ch, n, z);
  if (zout) memcpy(zout, z, sizeof(double) * (size_t)n);
  if (ch->cfg.usesvd) {
    orc_dgemv('N', n, R, n, z, p);
  } else {
    memcpy(p, z, sizeof(double) * (size_t)n); /* dcopy, matutils.F90:108 */
    orc_dtrmv_ut(n, R, n, p);
  }
  for (int i = 0; i < n; i++) newpar[i] = oldpar[i] + p[i];
  free(z);
  free(p);
}

/* MCMC_DRAM.F90:100-118 */
static double mcmc_alpha(const orc_chain* ch, const double* ss1, double sspri1, const double* ss2, double sspri2) {
  double sum = 0.0;
  for (int j = 0; j < ch->nycol; j++) sum = sum + (ss2[j] - ss1[j]) / ch->sigma2[j];
  double tst = -0.5 * (sum + (sspri2 - sspri1));
  if (tst >= 0.0) return 1.0;
  if (tst < LOG_REALMIN) return 0.0;
  return exp(tst);
}

/* MCMC_DRAM.F90:140-155 */
static int mcmc_reject(orc_chain* ch, double alpha) {
  int reject = 1;
  if (alpha >= 1.0) {
    reject = 0;
  } else if (alpha > 0.0) {
    double u = rng_uniform(&ch->rng);
    if (u <= alpha) reject = 0;
  }
  return reject;
}

/* MCMC_DRAM.F90:162-186 */
static double dr_alpha13(const orc_chain* ch, const double* oldpar, const double* ss1, double sspri1,
                         const double* newpar, const double* ss2, double sspri2, double alpha12,
                         const double* newpar2, const double* ss3, double sspri3) {
  int n = ch->npar;
  double alpha32;
  if (alpha12 == 0.0) {
    alpha32 = 0.0;
  } else {
    double sum = 0.0;
    for (int j = 0; j < ch->nycol; j++) sum = sum + (ss2[j] - ss3[j]) / ch->sigma2[j];
    double tst32 = -0.5 * (sum + (sspri2 - sspri3));
    alpha32 = fmin(1.0, exp(tst32));
  }
  double sum = 0.0;
  for (int j = 0; j < ch->nycol; j++) sum = sum + (ss3[j] - ss1[j]) / ch->sigma2[j];
  double l2 = -0.5 * (sum + (sspri3 - sspri1));
  double* v = (double*)malloc(sizeof(double) * (size_t)n * 2);
  double* w = v + n;
  double qa = 0.0, qb = 0.0;
  for (int i = 0; i < n; i++) v[i] = newpar2[i] - newpar[i];
  orc_dsymv_u(n, ch->iC, n, v, w);
  for (int i = 0; i < n; i++) qa = qa + w[i] * v[i];
  for (int i = 0; i < n; i++) v[i] = oldpar[i] - newpar[i];
  orc_dsymv_u(n, ch->iC, n, v, w);
  for (int i = 0; i < n; i++) qb = qb + w[i] * v[i];
  free(v);
  double q1 = -0.5 * (qa - qb);
  /* min(1, NaN) is processor dependent in Fortran: NaN -> reject via mcmc_reject */
  double a13 = exp(l2 + q1) * (1.0 - alpha32) / (1.0 - alpha12);
  if (a13 != a13) return a13;
  return fmin(1.0, a13);
}

/* MCMC_DRAM.F90:192-206 */
static void updatesigma2(orc_chain* ch, const double* ss) {
  if (ch->cfg.updatesigma != 0) {
    for (int j = 0; j < ch->nycol; j++) {
      double g = random_gamma1(&ch->rng, ch->cfg.N0 / 2.0 + (double)ch->nobs[j] / 2.0,
                               2.0 / (ch->cfg.N0 * ch->S02 + ss[j]));
      ch->sigma2[j] = 1.0 / g;
    }
  }
}

/* MCMC_aux.F90:166-185 (memory mode) */
static void savechain(orc_chain* ch, const double* par, const double* ss, int reject) {
  int ld = ch->cfg.nsimu;
  if (reject) {
    ch->chain[IDX(ch->chainind, ch->ncolchain, ld)] += 1.0;
  } else {
    ch->chainind++;
    for (int k = 1; k <= ch->npar; k++) ch->chain[IDX(ch->chainind, k, ld)] = par[k - 1];
    ch->chain[IDX(ch->chainind, ch->ncolchain, ld)] = 1.0;
    for (int k = 1; k <= ch->nycol; k++) ch->sschain[IDX(ch->chainind, k, ld)] = ss[k - 1];
  }
  ch->sschain[IDX(ch->chainind, ch->nycol + 1, ld)] = ch->chain[IDX(ch->chainind, ch->ncolchain, ld)];
  if (ch->cfg.updatesigma != 0)
    for (int k = 1; k <= ch->nycol; k++) ch->s2chain[IDX(ch->simuind, k, ld)] = ch->sigma2[k - 1];
}

/* call covmat on rows istart..iend of the stored chain with given weights */
static void covmat_rows(orc_chain* ch, int istart, int iend, const double* w, int nw, int update) {
  int ld = ch->cfg.nsimu;
  orc_covmat(&ch->chain[IDX(istart, 1, ld)], iend - istart + 1, ld, ch->npar, ch->chaincmat, w, nw,
             ch->chainmean, &ch->chainwsum, update);
}

/* MCMC_adapt.F90:12-174 */
static void mcmc_adapt(orc_chain* ch, int simuind) {
  const orc_cfg* c = &ch->cfg;
  int n = ch->npar, ld = c->nsimu;
  size_t nn = (size_t)n * n;
  if (c->doadapt == 0 && c->doburnin == 0) return;               /* :42 */
  if (c->adaptend > 0 && simuind > c->adaptend) return;          /* :43 */
  int ma = (c->adaptint > 0) ? (simuind % c->adaptint) : 1;
  int mb = (c->badaptint > 0) ? (simuind % c->badaptint) : 1;
  if (ma != 0 && mb != 0) return;                                /* :45-46 */

  if (simuind < c->burnintime && c->doburnin != 0 && mb == 0) {  /* :60-61 */
    double staypc = (double)ch->stayed / (double)simuind;
    ch->istartind = ch->chainind;
    if (staypc > 1.0 - c->scalelimit) {                          /* :64-72 */
      for (size_t k = 0; k < nn; k++) ch->R[k] = ch->R[k] / c->scalefactor;
      if (c->dodr) {
        for (size_t k = 0; k < nn; k++) ch->R2[k] = ch->R2[k] / c->scalefactor;
        for (size_t k = 0; k < nn; k++) ch->iC[k] = ch->iC[k] * c->scalefactor * c->scalefactor;
      }
      return;
    } else if (staypc < c->scalelimit) {                         /* :73-82 */
      for (size_t k = 0; k < nn; k++) ch->R[k] = ch->R[k] * c->scalefactor;
      if (c->dodr) {
        for (size_t k = 0; k < nn; k++) ch->R2[k] = ch->R2[k] * c->scalefactor;
        for (size_t k = 0; k < nn; k++) ch->iC[k] = ch->iC[k] / c->scalefactor / c->scalefactor;
      }
      return;
    } else if (c->greedy != 0) {                                 /* :83-101 */
      ch->chainwsum = (double)c->initcmatn;
      memcpy(ch->chaincmat, ch->cmat0, sizeof(double) * nn);
      memcpy(ch->chainmean, ch->par0, sizeof(double) * (size_t)n);
      double one = 1.0;
      covmat_rows(ch, 1, ch->chainind, &one, 1, 1);
      ch->lastfreq = (int)ch->chain[IDX(ch->chainind, ch->ncolchain, ld)];
    }
    ch->lastind = ch->chainind;                                  /* :102 */
  } else if (simuind >= c->burnintime + c->adaptint + c->adapthist && c->doadapt != 0) { /* :105 */
    if (simuind == c->burnintime + c->adaptint + c->adapthist) { /* :108-114 */
      ch->chainwsum = (double)c->initcmatn;
      memcpy(ch->chaincmat, ch->cmat0, sizeof(double) * nn);
      memcpy(ch->chainmean, ch->par0, sizeof(double) * (size_t)n);
    }
    if (c->adapthist > 1) {                                      /* :116-136 AP */
      ch->istart = ch->chainind;
      int histsum = (int)ch->chain[IDX(ch->istart, ch->ncolchain, ld)];
      while (histsum < c->adapthist && ch->istart > 1) {
        ch->istart--;
        histsum += (int)ch->chain[IDX(ch->istart, ch->ncolchain, ld)];
      }
      ch->newfreq = (int)ch->chain[IDX(ch->istart, ch->ncolchain, ld)];
      ch->chain[IDX(ch->istart, ch->ncolchain, ld)] = (double)(ch->newfreq - histsum + c->adapthist);
      covmat_rows(ch, ch->istart, ch->chainind, &ch->chain[IDX(ch->istart, ch->ncolchain, ld)],
                  ch->chainind - ch->istart + 1, 0);
      ch->chain[IDX(ch->istart, ch->ncolchain, ld)] = (double)ch->newfreq;
    } else {                                                     /* :138-159 */
      ch->newfreq = (int)ch->chain[IDX(ch->lastind, ch->ncolchain, ld)];
      ch->chain[IDX(ch->lastind, ch->ncolchain, ld)] = (double)(ch->newfreq - ch->lastfreq);
      ch->istart = ch->lastind;
      covmat_rows(ch, ch->istart, ch->chainind, &ch->chain[IDX(ch->istart, ch->ncolchain, ld)],
                  ch->chainind - ch->istart + 1, 1);
      ch->chain[IDX(ch->lastind, ch->ncolchain, ld)] = (double)ch->newfreq;
      ch->lastfreq = (int)ch->chain[IDX(ch->chainind, ch->ncolchain, ld)];
      ch->lastind = ch->chainind;
    }
  } else {
    return;                                                      /* :161-166 */
  }
  if (!ch->pool) calculate_R(ch, ch->chaincmat);                 /* :168-171: on failure keep old R */
}

/* MCMC_run.F90:12-114 */
static void run_dram(orc_chain* ch, int upto) {
  const orc_cfg* c = &ch->cfg;
  int n = ch->npar, m = ch->nycol;
  double* oldpar = ch->oldpar;
  double* newpar = (double*)malloc(sizeof(double) * (size_t)n * 2);
  double* newpar2 = newpar + n;
  double* ss1 = ch->ss1;
  double* ss2 = (double*)malloc(sizeof(double) * (size_t)m * 2);
  double* ss3 = ss2 + m;
  double sspri1 = ch->sspri1, sspri2 = 0, sspri3 = 0, alpha12 = 0, alpha13;
  int reject = 0;
  if (!ch->started) {
    memcpy(oldpar, ch->par0, sizeof(double) * (size_t)n);
    sspri1 = model_priorfun(&ch->model, oldpar);
    model_ss(&ch->model, oldpar, ss1);
    savechain(ch, oldpar, ss1, reject);
    ch->started = 1;
    ch->next_i = 2;
  }
  int i;
  for (i = ch->next_i; i <= upto; i++) {
    ch->simuind = i;
    propose(ch, oldpar, ch->R, newpar, NULL);
    int inbounds = model_checkbounds(&ch->model, newpar);
    if (!inbounds) {
      if (!c->dodr) ch->bndstayed++;
      for (int j = 0; j < m; j++) ss2[j] = DBL_MAX;
      sspri2 = DBL_MAX;
      alpha12 = 0.0;
      reject = 1;
    } else {
      sspri2 = model_priorfun(&ch->model, newpar);
      model_ss(&ch->model, newpar, ss2);
      alpha12 = mcmc_alpha(ch, ss1, sspri1, ss2, sspri2);
      reject = mcmc_reject(ch, alpha12);
    }
    if (reject && c->dodr) {
      ch->drtries++;
      propose(ch, oldpar, ch->R2, newpar2, NULL);
      inbounds = model_checkbounds(&ch->model, newpar2);
      if (!inbounds) {
        ch->bndstayed++;
        reject = 1;
      } else {
        sspri3 = model_priorfun(&ch->model, newpar2);
        model_ss(&ch->model, newpar2, ss3);
        alpha13 = dr_alpha13(ch, oldpar, ss1, sspri1, newpar, ss2, sspri2, alpha12, newpar2, ss3, sspri3);
        if (getenv("ORC_TRACE")) fprintf(stderr, "trace i=%d a12=%.17g a13=%.17g nd=%llu\n", i, alpha12, alpha13, (unsigned long long)ch->rng.ndrawn);
        reject = mcmc_reject(ch, alpha13);
        if (!reject) {
          ch->draccepted++;
          memcpy(newpar, newpar2, sizeof(double) * (size_t)n);
          memcpy(ss2, ss3, sizeof(double) * (size_t)m);
          sspri2 = sspri3;
        }
      }
    }
    if (reject) {
      ch->stayed++;
    } else {
      memcpy(ss1, ss2, sizeof(double) * (size_t)m);
      sspri1 = sspri2;
      memcpy(oldpar, newpar, sizeof(double) * (size_t)n);
    }
    updatesigma2(ch, ss1);
    savechain(ch, oldpar, ss1, reject);
    mcmc_adapt(ch, i);
    if (getenv("ORC_TRACE")) fprintf(stderr, "step i=%d a12=%.17g rej=%d stayed=%d bnd=%d nd=%llu th=%.17g %.17g s2=%.17g\n", i, alpha12, reject, ch->stayed, ch->bndstayed, (unsigned long long)ch->rng.ndrawn, oldpar[0], oldpar[1], ch->sigma2[0]);
    if (ch->rng.exhausted) { ch->status |= ORC_ST_RNG_EXHAUSTED; i++; break; }
  }
  ch->next_i = i;
  ch->sspri1 = sspri1;
  free(newpar);
  free(ss2);
}

/* MCMC_DRAM.F90:124-135: critical value of -2*log(lik); one uniform is ALWAYS drawn */
static double mcmc_sscrit(orc_chain* ch, const double* ss1, double priss1) {
  double u = rng_uniform(&ch->rng);
  double sum = 0.0;
  for (int j = 0; j < ch->nycol; j++) sum = sum + ss1[j] / ch->sigma2[j];
  return -2.0 * log(u) + sum + priss1;
}

/* MCMC_run_er.F90:12-107 -- early-rejection MH.  The user function ssfunction_er may stop summing
 * once its partial sum reaches sscrit; the library default (ssfunction_er0.f90) evaluates the full
 * ssfunction, which is what this restatement does.  Delayed rejection is switched off (:24-27). */
static void run_er(orc_chain* ch, int upto) {
  orc_cfg* c = &ch->cfg;
  int n = ch->npar, m = ch->nycol;
  double* oldpar = ch->oldpar;
  double* newpar = (double*)malloc(sizeof(double) * (size_t)n);
  double* ss1 = ch->ss1;
  double* ss2 = (double*)malloc(sizeof(double) * (size_t)m);
  double sspri1 = ch->sspri1, sspri2 = 0, sscrit;
  int reject = 0;
  c->dodr = 0; /* :24-27 */
  if (!ch->started) {
    memcpy(oldpar, ch->par0, sizeof(double) * (size_t)n);
    sspri1 = model_priorfun(&ch->model, oldpar);
    model_ss(&ch->model, oldpar, ss1);
    savechain(ch, oldpar, ss1, reject);
    ch->started = 1;
    ch->next_i = 2;
  }
  int i;
  for (i = ch->next_i; i <= upto; i++) {
    ch->simuind = i;
    propose(ch, oldpar, ch->R, newpar, NULL);
    int inbounds = model_checkbounds(&ch->model, newpar);
    if (!inbounds) { /* :54-57 */
      ch->bndstayed++;
      reject = 1;
    } else {
      sscrit = mcmc_sscrit(ch, ss1, sspri1);              /* :59 */
      sspri2 = model_priorfun(&ch->model, newpar);        /* :60 */
      if (sspri2 >= sscrit) {                             /* :62-67 */
        reject = 1;
        ch->erstayed++;
      } else {
        sscrit = ch->sigma2[0] * (sscrit - sspri2);       /* :71, "problem here if nycol > 1" */
        model_ss(&ch->model, newpar, ss2);                /* ssfunction_er(newpar,sscrit), default = ssfunction */
        double sum = 0.0;
        for (int j = 0; j < m; j++) sum = sum + ss2[j];
        reject = (sum >= sscrit) ? 1 : 0;                 /* :74-80 */
      }
    }
    if (reject) {
      ch->stayed++;
    } else {
      memcpy(ss1, ss2, sizeof(double) * (size_t)m);
      sspri1 = sspri2;
      memcpy(oldpar, newpar, sizeof(double) * (size_t)n);
    }
    updatesigma2(ch, ss1);
    savechain(ch, oldpar, ss1, reject);
    mcmc_adapt(ch, i);
    if (ch->rng.exhausted) { ch->status |= ORC_ST_RNG_EXHAUSTED; i++; break; }
  }
  ch->next_i = i;
  ch->sspri1 = sspri1;
  free(newpar);
  free(ss2);
}

/* MCMC_run_ram.F90:13-83, 87-101, 104-179 */
static void run_ram(orc_chain* ch, int upto) {
  const orc_cfg* c = &ch->cfg;
  int n = ch->npar, m = ch->nycol;
  double* oldpar = ch->oldpar;
  double* newpar = (double*)malloc(sizeof(double) * (size_t)n * 5);
  double *u = newpar + n, *xv = newpar + 2 * n, *cc = newpar + 3 * n, *sv = newpar + 4 * n;
  double* ss1 = ch->ss1;
  double* ss2 = (double*)malloc(sizeof(double) * (size_t)m);
  double sspri1 = ch->sspri1, sspri2 = 0;
  double alpha12 = ch->alpha12; /* undefined in the reference before the first in-bounds proposal: declared 0 */
  int reject = 0;
  if (!ch->started) {
    memcpy(oldpar, ch->par0, sizeof(double) * (size_t)n);
    sspri1 = model_priorfun(&ch->model, oldpar);
    model_ss(&ch->model, oldpar, ss1);
    savechain(ch, oldpar, ss1, reject);
    ch->started = 1;
    ch->next_i = 2;
  }
  int i;
  for (i = ch->next_i; i <= upto; i++) {
    ch->simuind = i;
    propose(ch, oldpar, ch->R, newpar, u);
    int inbounds = model_checkbounds(&ch->model, newpar);
    if (!inbounds) { /* alpha12 keeps its previous value, MCMC_run_ram.F90:52-55 */
      ch->bndstayed++;
      reject = 1;
    } else {
      sspri2 = model_priorfun(&ch->model, newpar);
      model_ss(&ch->model, newpar, ss2);
      alpha12 = mcmc_alpha(ch, ss1, sspri1, ss2, sspri2);
      reject = mcmc_reject(ch, alpha12);
    }
    if (reject) {
      ch->stayed++;
    } else {
      memcpy(ss1, ss2, sizeof(double) * (size_t)m);
      sspri1 = sspri2;
      memcpy(oldpar, newpar, sizeof(double) * (size_t)n);
    }
    updatesigma2(ch, ss1);
    savechain(ch, oldpar, ss1, reject);
    /* MCMC_adapt_ram, MCMC_run_ram.F90:104-179 */
    if (c->doadapt != 0 && !(i < c->burnintime && c->doburnin != 0)) {
      double a = 1.0 / pow((double)(float)i, c->nuparam) * (alpha12 - c->alphatarget); /* real(simuind) is single */
      double su2 = 0.0;
      for (int k = 0; k < n; k++) su2 = su2 + u[k] * u[k];
      if (a >= 0.0) {
        for (int k = 0; k < n; k++) xv[k] = u[k] / su2 * a;
        orc_dchud(ch->R, n, n, xv, cc, sv);
      } else {
        for (int k = 0; k < n; k++) xv[k] = -u[k] / su2 * a;
        int info = orc_dchdd(ch->R, n, n, xv, cc, sv);
        if (info != 0) ch->status |= ORC_ST_DOWNDATE_FAIL; /* reference stops: declared flag+skip */
      }
    }
    if (ch->rng.exhausted) { ch->status |= ORC_ST_RNG_EXHAUSTED; i++; break; }
  }
  ch->next_i = i;
  ch->sspri1 = sspri1;
  ch->alpha12 = alpha12;
  free(newpar);
  free(ss2);
}

/* MCMC_run_scam.F90:12-91, 94-138 */
static void run_scam(orc_chain* ch, int upto) {
  const orc_cfg* c = &ch->cfg;
  int n = ch->npar, m = ch->nycol, ld = c->nsimu;
  double* oldpar = ch->oldpar;
  double* newpar = (double*)malloc(sizeof(double) * (size_t)n * 2);
  double* rotpar = newpar + n;
  double* ss1 = ch->ss1;
  double* ss2 = (double*)malloc(sizeof(double) * (size_t)m);
  double sspri1 = ch->sspri1, sspri2 = 0, alpha12;
  int rejall = 0, reject;
  if (!ch->started) {
    memcpy(oldpar, ch->par0, sizeof(double) * (size_t)n);
    sspri1 = model_priorfun(&ch->model, oldpar);
    model_ss(&ch->model, oldpar, ss1);
    savechain(ch, oldpar, ss1, rejall);
    ch->started = 1;
    ch->next_i = 2;
  }
  int i;
  for (i = ch->next_i; i <= upto; i++) {
    ch->simuind = i;
    rejall = 1;
    for (int j = 1; j <= n; j++) {
      /* MCMC_propose_sc, :94-117 */
      orc_dgemv('T', n, ch->R, n, oldpar, rotpar);
      double z;
      orc_normals(ch, 1, &z);
      z = z * ch->qcovstd[j - 1];
      rotpar[j - 1] = rotpar[j - 1] + z;
      orc_dgemv('N', n, ch->R, n, rotpar, newpar);
      int inbounds = model_checkbounds(&ch->model, newpar);
      if (!inbounds) {
        if (!c->dodr) ch->bndstayed++;
        alpha12 = 0.0;
        reject = 1;
      } else {
        sspri2 = model_priorfun(&ch->model, newpar);
        model_ss(&ch->model, newpar, ss2);
        alpha12 = mcmc_alpha(ch, ss1, sspri1, ss2, sspri2);
        reject = mcmc_reject(ch, alpha12);
      }
      if (!reject) {
        memcpy(ss1, ss2, sizeof(double) * (size_t)m);
        sspri1 = sspri2;
        memcpy(oldpar, newpar, sizeof(double) * (size_t)n);
        rejall = 0;
      }
    }
    if (rejall) ch->stayed++;
    ch->sschain[IDX(ch->chainind, m + 1, ld)] = ch->chain[IDX(ch->chainind, ch->ncolchain, ld)]; /* :80 */
    updatesigma2(ch, ss1);
    savechain(ch, oldpar, ss1, rejall);
    mcmc_adapt(ch, i);
    if (ch->rng.exhausted) { ch->status |= ORC_ST_RNG_EXHAUSTED; i++; break; }
  }
  ch->next_i = i;
  ch->sspri1 = sspri1;
  free(newpar);
  free(ss2);
}

/* ------------------------------------------------------------ lifecycle */
/* MCMC_init.F90:75-158 */
orc_chain* orc_create(const orc_cfg* cfg, int model_id, const double* blob, long blob_len,
                      int npar, int nycol, const double* par0, const double* cmat0,
                      const double* sigma2, const int* nobs) {
  orc_chain* ch = (orc_chain*)calloc(1, sizeof *ch);
  ch->cfg = *cfg;
  orc_check_params(&ch->cfg);
  ch->npar = npar; ch->nycol = nycol; ch->ncolchain = npar + 1;
  ch->model.id = model_id; ch->model.npar = npar; ch->model.ny = nycol;
  ch->model.blob = blob; ch->model.blob_len = blob_len;
  size_t n = (size_t)npar, nn = n * n, ns = (size_t)(ch->cfg.nsimu > 0 ? ch->cfg.nsimu : 1);
  ch->par0 = (double*)malloc(sizeof(double) * n); memcpy(ch->par0, par0, sizeof(double) * n);
  ch->oldpar = (double*)calloc(n, sizeof(double));
  ch->cmat0 = (double*)malloc(sizeof(double) * nn); memcpy(ch->cmat0, cmat0, sizeof(double) * nn);
  ch->sigma2 = (double*)malloc(sizeof(double) * (size_t)nycol); memcpy(ch->sigma2, sigma2, sizeof(double) * (size_t)nycol);
  ch->nobs = (int*)malloc(sizeof(int) * (size_t)nycol); memcpy(ch->nobs, nobs, sizeof(int) * (size_t)nycol);
  ch->R = (double*)calloc(nn, sizeof(double));
  ch->R2 = (double*)calloc(nn, sizeof(double));
  ch->iC = (double*)calloc(nn, sizeof(double));
  ch->qcovstd = (double*)calloc(n, sizeof(double));
  ch->chaincmat = (double*)malloc(sizeof(double) * nn); memcpy(ch->chaincmat, cmat0, sizeof(double) * nn);
  ch->chainmean = (double*)malloc(sizeof(double) * n); memcpy(ch->chainmean, par0, sizeof(double) * n);
  ch->chainwsum = (double)ch->cfg.initcmatn;
  ch->chain = (double*)calloc(ns * (n + 1), sizeof(double));
  ch->sschain = (double*)calloc(ns * (size_t)(nycol + 1), sizeof(double));
  ch->s2chain = (double*)calloc(ns * (size_t)nycol, sizeof(double));
  ch->istart = 1; ch->istartind = 1; ch->lastind = 1; ch->lastfreq = 0; ch->newfreq = 0;
  ch->chainind = 0; ch->simuind = 1;
  ch->rng.mode = 1; ch->rng.seed = 0; ch->rng.chain = 0;
  /* MCMC_init.F90:108-110: MCMC_calculate_R(cmat0) on the caller's cmat0 */
  double* c0 = (double*)malloc(sizeof(double) * nn);
  memcpy(c0, cmat0, sizeof(double) * nn);
  calculate_R(ch, c0);
  free(c0);
  ch->S02 = ch->cfg.S02;
  if (ch->S02 <= 0.0) ch->S02 = ch->sigma2[0]; /* MCMC_init.F90:114-116 */
  ch->ss1 = (double*)calloc((size_t)nycol, sizeof(double));
  ch->alpha12 = 0.0;
  return ch;
}

void orc_set_prior(orc_chain* ch, const double* mu, const double* sig) {
  size_t n = (size_t)ch->npar;
  ch->model.pmu = (double*)malloc(sizeof(double) * n);
  ch->model.psig = (double*)malloc(sizeof(double) * n);
  memcpy(ch->model.pmu, mu, sizeof(double) * n);
  memcpy(ch->model.psig, sig, sizeof(double) * n);
}

void orc_set_rng_injected(orc_chain* ch, const double* u, long n) {
  ch->rng.mode = 0; ch->rng.inj = u; ch->rng.ninj = n; ch->rng.pos = 0;
}
void orc_set_rng_philox(orc_chain* ch, uint64_t seed, uint64_t chain_id) {
  ch->rng.mode = 1; ch->rng.seed = seed; ch->rng.chain = chain_id;
}

/* advance to step index `upto` (= simuind; <= nsimu); the first call also does the initial point */
int orc_advance(orc_chain* ch, int upto) { /* mcmc_main.F90:29-37 */
  if (ch->cfg.nsimu < 1) return -1;
  if (upto > ch->cfg.nsimu) upto = ch->cfg.nsimu;
  switch (ch->cfg.method) {
    case ORC_SCAM: run_scam(ch, upto); break;
    case ORC_RAM: run_ram(ch, upto); break;
    case ORC_ER: run_er(ch, upto); break;
    default: run_dram(ch, upto);
  }
  return ch->status;
}

int orc_run(orc_chain* ch) { return orc_advance(ch, ch->cfg.nsimu); }

/* pooled adaptation (an extension of the build, see mcmcf90_b200/csrc/pool.cuh): MCMC_adapt keeps
 * updating chaincmat/chainmean/chainwsum but leaves R alone; the test harness merges the chains'
 * accumulators and hands every chain the factor of the pooled covariance */
void orc_set_pool(orc_chain* ch, int on) { ch->pool = on; }
/* MCMC_calculate_R (MCMC_adapt.F90:181-230) on a caller-supplied covariance (column-major n x n) */
int orc_factor_from_cov(orc_chain* ch, const double* cov) {
  size_t nn = (size_t)ch->npar * ch->npar;
  double* c = (double*)malloc(sizeof(double) * nn);
  memcpy(c, cov, sizeof(double) * nn);
  int before = ch->status;
  ch->status = 0;
  int info = calculate_R(ch, c);
  int st = ch->status;
  ch->status = before | st;
  free(c);
  return info != 0 || st != 0;
}
/* overwrite the proposal factor (RAM pooling: R = chol(mean R'R)) */
void orc_set_R(orc_chain* ch, const double* R) { memcpy(ch->R, R, sizeof(double) * (size_t)ch->npar * ch->npar); }

void orc_free(orc_chain* ch) {
  if (!ch) return;
  free(ch->par0); free(ch->oldpar); free(ch->cmat0); free(ch->sigma2); free(ch->nobs);
  free(ch->R); free(ch->R2); free(ch->iC); free(ch->qcovstd); free(ch->chaincmat); free(ch->chainmean);
  free(ch->chain); free(ch->sschain); free(ch->s2chain); free(ch->ss1);
  free(ch->model.pmu); free(ch->model.psig);
  free(ch);
}

const double* orc_chain_ptr(const orc_chain* ch) { return ch->chain; }
const double* orc_sschain_ptr(const orc_chain* ch) { return ch->sschain; }
const double* orc_s2chain_ptr(const orc_chain* ch) { return ch->s2chain; }
const double* orc_R_ptr(const orc_chain* ch) { return ch->R; }
const double* orc_R2_ptr(const orc_chain* ch) { return ch->R2; }
const double* orc_iC_ptr(const orc_chain* ch) { return ch->iC; }
const double* orc_qcovstd_ptr(const orc_chain* ch) { return ch->qcovstd; }
const double* orc_cmat_ptr(const orc_chain* ch) { return ch->chaincmat; }
const double* orc_mean_ptr(const orc_chain* ch) { return ch->chainmean; }
const double* orc_sigma2_ptr(const orc_chain* ch) { return ch->sigma2; }
const double* orc_par_ptr(const orc_chain* ch) { return ch->oldpar; }
double orc_wsum(const orc_chain* ch) { return ch->chainwsum; }
int orc_erstayed(const orc_chain* ch) { return ch->erstayed; }
void orc_counters(const orc_chain* ch, long* o) {
  o[0] = ch->stayed; o[1] = ch->bndstayed; o[2] = ch->draccepted; o[3] = ch->drtries;
  o[4] = ch->chainind; o[5] = ch->simuind; o[6] = ch->status; o[7] = (long)ch->rng.ndrawn;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* weighted sample mean/covariance of one finished chain from its run-length rows */
static void chain_moments(const orc_chain* ch, double* mean, double* cov) {
  int n = ch->npar, ld = ch->cfg.nsimu;
  double W = 0.0;
  for (int k = 0; k < n; k++) mean[k] = 0.0;
  for (int r = 1; r <= ch->chainind; r++) {
    double w = ch->chain[IDX(r, ch->ncolchain, ld)];
    W += w;
    for (int k = 1; k <= n; k++) mean[k - 1] += w * ch->chain[IDX(r, k, ld)];
  }
  for (int k = 0; k < n; k++) mean[k] /= W;
  for (int a = 0; a < n * n; a++) cov[a] = 0.0;
  for (int r = 1; r <= ch->chainind; r++) {
    double w = ch->chain[IDX(r, ch->ncolchain, ld)];
    for (int a = 1; a <= n; a++)
      for (int b = 1; b <= n; b++)
        cov[(a - 1) * n + (b - 1)] += w * (ch->chain[IDX(r, a, ld)] - mean[a - 1]) * (ch->chain[IDX(r, b, ld)] - mean[b - 1]);
  }
  for (int a = 0; a < n * n; a++) cov[a] /= (W - 1.0);
}

typedef struct {
  const orc_cfg* cfg; int model_id; const double* blob; long blob_len; int npar, nycol; long nchains;
  const double *par0, *cmat0, *sigma2; const int* nobs; uint64_t seed, chain0;
  double *last_par, *mean, *cmat; long* counters; double *chain_mean_out, *chain_cov_out;
  long next; pthread_mutex_t mu;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* J = (batch_job*)arg;
  int npar = J->npar;
  for (;;) {
    pthread_mutex_lock(&J->mu);
    long cidx = J->next++;
    pthread_mutex_unlock(&J->mu);
    if (cidx >= J->nchains) break;
    orc_chain* ch = orc_create(J->cfg, J->model_id, J->blob, J->blob_len, npar, J->nycol,
                               J->par0 + (size_t)cidx * npar, J->cmat0, J->sigma2, J->nobs);
    orc_set_rng_philox(ch, J->seed, J->chain0 + (uint64_t)cidx);
    orc_run(ch);
    if (J->last_par) memcpy(J->last_par + (size_t)cidx * npar, ch->oldpar, sizeof(double) * (size_t)npar);
    if (J->mean) memcpy(J->mean + (size_t)cidx * npar, ch->chainmean, sizeof(double) * (size_t)npar);
    if (J->cmat) memcpy(J->cmat + (size_t)cidx * npar * npar, ch->chaincmat, sizeof(double) * (size_t)npar * npar);
    if (J->counters) orc_counters(ch, J->counters + (size_t)cidx * 8);
    if (J->chain_mean_out && J->chain_cov_out)
      chain_moments(ch, J->chain_mean_out + (size_t)cidx * npar, J->chain_cov_out + (size_t)cidx * npar * npar);
    orc_free(ch);
  }
  return NULL;
}

/* one chain per host thread (the reference is one chain per process) */
int orc_run_batch(const orc_cfg* cfg, int model_id, const double* blob, long blob_len,
                  int npar, int nycol, long nchains, const double* par0,
                  const double* cmat0, const double* sigma2, const int* nobs,
                  uint64_t seed, uint64_t chain0, int nthreads,
                  double* last_par, double* mean, double* cmat, long* counters,
                  double* chain_mean_out, double* chain_cov_out, double* seconds) {
  batch_job J = {cfg, model_id, blob, blob_len, npar, nycol, nchains, par0, cmat0, sigma2, nobs, seed, chain0,
                 last_par, mean, cmat, counters, chain_mean_out, chain_cov_out, 0, PTHREAD_MUTEX_INITIALIZER};
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 1024) nthreads = 1024;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  double t0 = now_s();
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &J);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  if (seconds) *seconds = now_s() - t0;
  free(th);
  return 0;
}
