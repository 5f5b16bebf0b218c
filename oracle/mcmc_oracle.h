/*
 * mcmc_oracle.h -- CPU restatement of mcmcf90's adaptive Metropolis-Hastings hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mcmcf90_b200/ may include, link or call
 * this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / reported CPU baseline.
 *
 * PARITY NOT PINNED TO A RUN OF THE REFERENCE: the reference (Fortran 90 + F77) cannot be compiled in this image
 * or on the GPU box (no Fortran compiler; profiles/r02_probe_fortran.txt) and ships no golden vectors,
 * known-answer tests or expected outputs.  This restatement follows the reference line by
 * line (citations are file:line into /root/reference) and is pinned by (0) an independent second restatement
 * of the Fortran, oracle/restate_np.py (numpy + real BLAS/LAPACK), which must reproduce its chain indices, counters
 * and draw counts exactly and its values to 1e-9 on every sampler (tests/test_ref_parity.py); (a) analytic
 * known answers for the shipped testcase, (b) scipy's LAPACK/BLAS (the same third-party
 * routines the reference links: dpotrf, dpotri, dtrmv, dsymv, dgemv, drotg, dnrm2),
 * (c) mathematical identities (R'R == C +/- xx' for dchud/dchdd, distribution moments
 * for the variate generators) and (d) the Random123 known-answer vectors for Philox.
 *
 * Third-party arithmetic that is NOT under /root/reference (un-vendored, un-versioned
 * link-time deps `-llapack -lblas`, testcases/Makefile:11): restated here in the
 * operation order of reference (netlib) BLAS / unblocked LAPACK 3.x:
 * dtrmv, dgemv, dsymv, ddot, dnrm2 (classic scaled form), drotg (classic),
 * dpotf2, dtrti2, dlauu2.  dgesvd is replaced by a cyclic Jacobi eigen-solver
 * (symmetric PSD input only; column signs fixed by convention, see orc_symeig).
 */
#ifndef MCMC_ORACLE_H
#define MCMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* method codes (mcmc_main.F90:29-37) */
enum { ORC_DRAM = 0, ORC_RAM = 1, ORC_SCAM = 2, ORC_ER = 3 };
/* built-in user models (the plugin side of external_inc.h:14-28) */
enum { ORC_MODEL_EXPREG = 0, ORC_MODEL_GAUSS = 1, ORC_MODEL_BANANA = 2, ORC_MODEL_HIER = 3 };
/* status bits */
enum { ORC_ST_CHOLFAIL = 1, ORC_ST_DOWNDATE_FAIL = 2, ORC_ST_RNG_EXHAUSTED = 4, ORC_ST_SVDFAIL = 8 };

/* mirror of namelist &mcmc (mcmcinit.F90:74-82), kernel-relevant fields only */
typedef struct {
  int method;
  int nsimu, doadapt, adaptint, adapthist, adaptend, initcmatn;
  int doburnin, burnintime, badaptint, greedy;
  double scalelimit, scalefactor, drscale, condmax;
  double N0, S02;
  int updatesigma;
  double alphatarget, nuparam;
  /* derived by orc_check_params (mcmcinit.F90:235-368) */
  int dodr, doscam, usesvd;
} orc_cfg;

typedef struct orc_chain orc_chain;

void orc_default_cfg(orc_cfg* c);          /* mcmcinit.F90:184-230 */
void orc_check_params(orc_cfg* c);         /* mcmcinit.F90:235-368 */

orc_chain* orc_create(const orc_cfg* cfg, int model_id, const double* blob, long blob_len,
                      int npar, int nycol, const double* par0, const double* cmat0,
                      const double* sigma2, const int* nobs);
void orc_set_prior(orc_chain* ch, const double* mu, const double* sig);   /* priorfun.f90:97-100 */
void orc_set_rng_injected(orc_chain* ch, const double* u, long n);
void orc_set_rng_philox(orc_chain* ch, uint64_t seed, uint64_t chain_id);
int  orc_run(orc_chain* ch);               /* dispatch, mcmc_main.F90:29-37 */
int  orc_advance(orc_chain* ch, int upto); /* the same loop, stopped at step index `upto` and resumable */
/* pooled adaptation (extension of the build, no reference counterpart; mcmcf90_b200/csrc/pool.cuh) */
void orc_set_pool(orc_chain* ch, int on);
int  orc_factor_from_cov(orc_chain* ch, const double* cov);   /* MCMC_calculate_R on a given covariance */
void orc_set_R(orc_chain* ch, const double* R);
void orc_free(orc_chain* ch);

/* results (column-major like the Fortran arrays; leading dimension nsimu) */
const double* orc_chain_ptr(const orc_chain* ch);     /* chain(nsimu, npar+1) */
const double* orc_sschain_ptr(const orc_chain* ch);   /* sschain(nsimu, nycol+1) */
const double* orc_s2chain_ptr(const orc_chain* ch);   /* s2chain(nsimu, nycol) */
const double* orc_R_ptr(const orc_chain* ch);         /* R(npar,npar) col-major */
const double* orc_R2_ptr(const orc_chain* ch);
const double* orc_iC_ptr(const orc_chain* ch);
const double* orc_qcovstd_ptr(const orc_chain* ch);
const double* orc_cmat_ptr(const orc_chain* ch);
const double* orc_mean_ptr(const orc_chain* ch);
const double* orc_sigma2_ptr(const orc_chain* ch);
const double* orc_par_ptr(const orc_chain* ch);       /* last oldpar */
double orc_wsum(const orc_chain* ch);
int orc_erstayed(const orc_chain* ch);               /* mcmc.F90:49, steps rejected by the prior alone in method 'er' */
/* counters: [stayed, bndstayed, draccepted, drtries, chainind, simuind, status, ndrawn_lo] */
void orc_counters(const orc_chain* ch, long* out8);

/* many independent chains (one per OpenMP thread) -- CPU baseline and statistics.
 * par0 is (nchains x npar) row-major; outputs (all optional, may be NULL):
 * last_par (nchains x npar), mean (nchains x npar), cmat (nchains x npar x npar),
 * counters (nchains x 8).  Returns 0, fills *seconds with the wall time of the
 * stepping loops only (allocation excluded). */
int orc_run_batch(const orc_cfg* cfg, int model_id, const double* blob, long blob_len,
                  int npar, int nycol, long nchains, const double* par0,
                  const double* cmat0, const double* sigma2, const int* nobs,
                  uint64_t seed, uint64_t chain0, int nthreads,
                  double* last_par, double* mean, double* cmat, long* counters,
                  double* chain_mean_out, double* chain_cov_out, double* seconds);

/* primitives exported for unit tests */
void   orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_philox_uniform(uint64_t seed, uint64_t chain, uint64_t k);
void   orc_normals(orc_chain* ch, int n, double* out);              /* mcmcrand.F90:60-83 */
double orc_gamma(orc_chain* ch, double a, double b);                /* mcmcrand.F90:86-111 */
void   orc_dtrmv_ut(int n, const double* A, int lda, double* x);    /* dtrmv('u','t','n') */
void   orc_dgemv(char trans, int n, const double* A, int lda, const double* x, double* y);
void   orc_dsymv_u(int n, const double* A, int lda, const double* x, double* y);
int    orc_dpotf2_u(int n, double* A, int lda);
int    orc_dpotri_u(int n, double* A, int lda);
void   orc_drotg(double* a, double* b, double* c, double* s);
void   orc_dchud(double* r, int ldr, int p, const double* x, double* c, double* s);
int    orc_dchdd(double* r, int ldr, int p, const double* x, double* c, double* s);
void   orc_covmat(const double* x, int n, int ldx, int p, double* cmat, const double* w, int nw,
                  double* xmean, double* wsum, int update);
int    orc_symeig(int n, const double* A, double* U, double* s);    /* replaces dgesvd('A','N') */
double orc_model_ss(int model_id, const double* blob, const double* theta, int npar);

#ifdef __cplusplus
}
#endif
#endif
