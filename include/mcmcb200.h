/*
 * mcmcb200.h -- C ABI of the B200-native batched adaptive Metropolis-Hastings path.
 *
 * This is the drop-in boundary for mcmcf90's sampling hot path.  In the reference the
 * driver dispatches on `method` to MCMC_run / MCMC_run_ram / MCMC_run_scam
 * (mcmc_main.F90:29-37), which run ONE chain on module-global state (mcmc.F90:28-60).
 * Here the same dispatch point hands `nchains` independent chains to CUDA kernels; the
 * host side (namelist, initialize, file output) stays where it is and talks to the
 * device through the plain-C calls below (ISO_C_BINDING from Fortran, ctypes from
 * Python, direct from C/C++).  No torch / C++ types cross this boundary.
 *
 * Conventions: every call returns 0 on success or a negative MCMCB_E* code and never
 * exits the process (the reference `stop`s, e.g. matutils.F90:764-789).  The caller
 * owns all host buffers.  One host thread per handle.  A handle drives `ngpus` GPUs of the box from that one thread
 * (the reference's host is a single-process program, mcmc_main.F90:12-44): chains are sharded over devices
 * `device .. device+ngpus-1` by contiguous global id and every call below fans out; host arrays stay chain-major over
 * ALL chains.  One handle per rank with `chain_offset` set (one process per GPU) works as well.
 * Matrices are column-major (Fortran) unless stated.
 */
#ifndef MCMCB200_H
#define MCMCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCMCB_ABI_VERSION 3

/* error codes */
#define MCMCB_OK 0
#define MCMCB_EINVAL (-1)       /* bad argument / call order */
#define MCMCB_ECUDA (-2)        /* CUDA runtime error, see mcmcb_last_error */
#define MCMCB_EUNSUPPORTED (-3) /* configuration the device path does not implement */
#define MCMCB_ENOMODEL (-4)     /* unknown user model name */
#define MCMCB_ENOMEM (-5)

/* method (namelist `method`, mcmcinit.F90:60; dispatch mcmc_main.F90:29-37) */
#define MCMCB_DRAM 0
#define MCMCB_RAM 1
#define MCMCB_SCAM 2
#define MCMCB_ER 3 /* early-rejection MH, MCMC_run_er.F90:12-107 (no delayed rejection) */

/* rng_mode */
#define MCMCB_RNG_PHILOX 0   /* Philox4x32-10 keyed by (seed, global chain id) -- replaces random_number */
#define MCMCB_RNG_INJECTED 1 /* consume caller-provided uniforms in the reference's draw order */

/* per-chain status bits (replace the reference's `stop`s) */
#define MCMCB_ST_CHOLFAIL 1       /* Cholesky failed in adaptation: old R kept (MCMC_adapt.F90:169-171) */
#define MCMCB_ST_DOWNDATE_FAIL 2  /* dchdd info=-1: update skipped (matutils.F90:716-722 stops) */
#define MCMCB_ST_RNG_EXHAUSTED 4  /* injected uniform stream ran out */
#define MCMCB_ST_SVDFAIL 8
#define MCMCB_ST_STORE_FULL 16    /* stored chain rows exceeded nsimu */

/* Mirror of namelist &mcmc (mcmcinit.F90:74-82; defaults 184-230), kernel-relevant
 * fields, followed by the batch fields the reference does not have. */
typedef struct mcmcb_config {
  int abi_version; /* MCMCB_ABI_VERSION */
  /* --- &mcmc --- */
  int method;
  int nsimu; /* chain length incl. the initial point; also the row capacity of stored chains */
  int doadapt, adaptint, adapthist, adaptend, initcmatn;
  int doburnin, burnintime, badaptint, greedy;
  double scalelimit, scalefactor, drscale, condmax;
  double N0, S02;
  int updatesigma;
  double alphatarget, nuparam;
  /* --- &mcmcb (new) --- */
  long long nchains;      /* chains owned by this handle */
  long long chain_offset; /* global id of this handle's first chain (Philox stream id) */
  unsigned long long seed;
  int rng_mode;
  int device;          /* CUDA device ordinal */
  int store_chains;    /* run-length chain/sschain/s2chain kept in HBM for the first store_chains chains (-1 = all) */
  int lanes_per_chain; /* 0 = auto; 1,2,4,8,16,32 lanes cooperate on one chain's ssfunction */
  int dump_stride;     /* >0: every dump_stride steps all chains' theta are streamed to pinned host buffers */
  int kernel;          /* 0 = auto; 1 = small-npar register kernel; 2 = large-npar warp kernel */
  int pool_adapt;      /* 1: cross-chain pooled adaptation -- at every adaptation tick the chains' (wsum, mean,
                          cmat) accumulators are merged over ALL chains of ALL handles (mcmcb_set_allreduce)
                          and every chain proposes from the factor of the pooled covariance; RAM: the chains'
                          R'R are averaged every adaptint steps (no reference counterpart) */
  int diag_stride;     /* >0: every diag_stride steps theta of every chain is folded into the per-chain running
                          moments behind mcmcb_diagnostics (R-hat / ESS) */
  int diag_lags;       /* autocovariance lags kept for the ESS, in units of diag_stride (0..MCMCB_DIAG_MAXLAGS) */
  int ngpus;           /* 0 or 1: one GPU (`device`); n > 1: this handle drives devices device..device+n-1 from one host
                          thread; pooled adaptation / diagnostics then reduce over NCCL inside the process */
  char model[32];      /* user-model name: "expreg", "gauss", "banana", "hier", or a plugin's name */
} mcmcb_config;

typedef struct mcmcb_handle_s* mcmcb_handle;

/* Load a user-model plugin (a shared library built against include/mcmcb200_plugin.cuh): its models
 * register themselves by name -- the run-time form of the reference's link-time override of ssfunction /
 * priorfun / checkbounds (external_inc.h:4-28).  Returns MCMCB_ENOMODEL when the library cannot be loaded. */
int mcmcb_load_plugin(const char* path);

/* namelist defaults, mcmcinit.F90:184-230 (MCMC_init_namelist) */
int mcmcb_default_config(mcmcb_config* cfg);
/* sanity rules + derived flags, mcmcinit.F90:235-368 (check_mcmcinit_parameters) */
int mcmcb_check_config(mcmcb_config* cfg, int* dodr, int* doscam, int* usesvd);

/* replaces the allocation half of MCMC_init (MCMC_init.F90:81-132) */
int mcmcb_create(const mcmcb_config* cfg, mcmcb_handle* out);
int mcmcb_destroy(mcmcb_handle h); /* MCMC_cleanup, MCMC_aux.F90:90-118 */
const char* mcmcb_last_error(mcmcb_handle h);

/* user-model data (what the plugin's ssfunction loads on first call, e.g. data.dat at
 * testcases/mcmcrun.F90:69-86).  Opaque blob of doubles, staged into shared memory. */
int mcmcb_set_data(mcmcb_handle h, const double* blob, size_t ndoubles);
/* default Gaussian prior read from priorsfile (priorfun.f90:58-100): mu[npar], sig[npar]
 * (sig<=0 disables a component).  NULL/NULL = flat prior. */
int mcmcb_set_priors(mcmcb_handle h, const double* mu, const double* sig, int npar);
/* what `initialize` returns (external_inc.h:33-42, initialize.F90:21-121):
 * par0 is npar values shared by all chains (par0_stride==0) or nchains rows of
 * par0_stride doubles; cmat0 is npar x npar column-major; sigma2/nobs have nycol entries. */
int mcmcb_set_initial(mcmcb_handle h, int npar, int nycol, const double* par0, long long par0_stride,
                      const double* cmat0, const double* sigma2, const int* nobs);
/* parity hook: u holds nchains rows of per_chain uniforms in [0,1), consumed in the
 * reference's draw order instead of Philox. */
int mcmcb_inject_uniforms(mcmcb_handle h, const double* u, size_t per_chain);

/* advance every chain by nsteps iterations of MCMC_LOOP (MCMC_run.F90:41-107,
 * MCMC_run_ram.F90:45-80, MCMC_run_scam.F90:38-88).  The first call also evaluates the
 * initial point (MCMC_run.F90:27-36).  Asynchronous; mcmcb_sync waits. */
int mcmcb_run(mcmcb_handle h, int nsteps);
int mcmcb_sync(mcmcb_handle h);

/* stored chain of one chain in the reference's layout (MCMC_aux.F90:166-185):
 * chain is (ld x (npar+1)) column-major, last column = repeat count; sschain is
 * (ld x (nycol+1)); s2chain is (ld x nycol) indexed by step (not compressed, Q16).
 * *nrows receives chainind.  Any output pointer may be NULL. */
int mcmcb_fetch_chain(mcmcb_handle h, long long chain, int ld, double* chain_out, double* sschain_out,
                      double* s2chain_out, int* nrows);

/* Per-chain state arrays, chain-major on the host side: out[chain * width + k].
 *  "par" (npar) "ss" (nycol) "sspri" (1) "sigma2" (nycol) "mean" (npar) "wsum" (1)
 *  "cmat" "R" "R2" "iC" (npar*npar, column-major, upper triangle authoritative)
 *  "qcovstd" (npar)
 *  "counters" (8 x int64: stayed, bndstayed, draccepted, drtries, chainind, simuind, status, ndrawn)
 *  "erstayed" (1 x int64: steps rejected by the prior alone in method 'er', mcmc.F90:49) */
int mcmcb_fetch(mcmcb_handle h, const char* what, void* out, size_t out_bytes);

/* One chain's adaptation state with typed arguments (no string keys): mean[npar], cmat[npar*npar] and R[npar*npar]
 * column-major, *wsum, sigma2[nycol], counters[8] as in "counters" above -- what a Fortran host copies back into
 * chainmean / chaincmat / chainwsum / R / sigma2 and its counters (mcmc.F90:28-55).  Any pointer may be NULL. */
int mcmcb_fetch_stats(mcmcb_handle h, long long chain, double* mean, double* cmat, double* wsum, double* R,
                      double* sigma2, long long* counters);

/* streamed dumps (MCMC_dump.F90:12-30 hook): pops the oldest completed snapshot of all
 * chains' theta (nchains x npar, chain-major) from the pinned ring; returns 1 if one was
 * copied, 0 if none pending. *step receives simuind of the snapshot. */
int mcmcb_dump_pop(mcmcb_handle h, double* out, size_t out_bytes, int* step);

/* The same snapshot with what MCMC_savechain records beside theta (MCMC_aux.F90:166-185): ss (nchains x nycol) and
 * sigma2 (nchains x nycol) of every chain at that step; ss / sigma2 may be NULL. */
int mcmcb_dump_pop_ex(mcmcb_handle h, double* par, double* ss, double* sigma2, size_t par_bytes, int* step);

/* ---- multi-GPU collectives: chains never interact in the reference, so the step loop has no
 * collective.  The two optional cross-chain features below reduce a few small vectors over every handle of
 * the job.  The library does not link a communication library: the host hands it ONE callback that
 * sum-reduces n doubles in place in DEVICE memory across all ranks, ordered on the given CUDA stream (or
 * synchronously): ncclAllReduce(buf, buf, n, ncclDouble, ncclSum, comm, stream) from C/Fortran,
 * torch.distributed.all_reduce from Python (mcmcf90_b200/parallel.py).  Without a callback the reduction
 * covers the handle's own chains only.  A handle with ngpus > 1 reduces over its own devices with NCCL and refuses a
 * callback (MCMCB_EUNSUPPORTED). */
typedef int (*mcmcb_allreduce_fn)(void* user, double* device_buf, size_t n, void* cuda_stream);
int mcmcb_set_allreduce(mcmcb_handle h, mcmcb_allreduce_fn fn, void* user);

/* pooled statistics of the last pooled adaptation tick (pool_adapt = 1): wsum, mean[npar],
 * cov[npar*npar] column-major.  Any pointer may be NULL. */
int mcmcb_pool_fetch(mcmcb_handle h, double* wsum, double* mean, double* cov);

#define MCMCB_DIAG_MAXLAGS 32
/* Convergence diagnostics over every chain of every handle (diag_stride > 0): potential scale reduction
 * R-hat (Gelman & Rubin 1992, between/within variances of the chains' snapshot means) and the effective
 * sample size of the pooled snapshots from the chains' mean autocovariances at lags 1..diag_lags (Geyer
 * initial-positive-sequence truncation).  Collective: every rank calls it at the same point.
 * rhat, ess, mean, var have npar entries (any may be NULL); *nsnap / *nchains_total receive the snapshots
 * per chain and the number of chains pooled. */
int mcmcb_diagnostics(mcmcb_handle h, double* rhat, double* ess, double* mean, double* var, long long* nsnap,
                      long long* nchains_total);
int mcmcb_diag_reset(mcmcb_handle h);

/* introspection for measurement */
void* mcmcb_stream(mcmcb_handle h);              /* cudaStream_t the kernels run on (first device of a group) */
void* mcmcb_stream_of(mcmcb_handle h, int k);    /* ... on the k-th device of a group handle */
int mcmcb_ngpus(mcmcb_handle h);                 /* devices this handle drives */
long long mcmcb_nccl_calls(mcmcb_handle h);      /* NCCL allreduce groups issued so far (group handles) */
long long mcmcb_launch_count(mcmcb_handle h);    /* kernels launched so far */
int mcmcb_info(mcmcb_handle h, int* npar, int* nycol, int* lanes_per_chain, int* kernel, int* threads_per_block,
               int* blocks, size_t* smem_bytes);
/* chains each thread of the register kernel runs side by side (1 when it does not apply): with B chains per thread
 * one shared-memory read of a datum serves B chains' ssfunction */
int mcmcb_chains_per_thread(mcmcb_handle h);
/* FP64 pipe microbenchmark: dependent-free DFMA chains on every SM; returns measured
 * TFLOP/s (2 flop per DFMA) and the elapsed milliseconds. */
int mcmcb_dfma_peak(int device, double* tflops, double* ms);
/* accuracy evidence for the device exp used by model code (include/mcmcb200_model.cuh):
 * out_fast[i] = mcmcb_exp_fast(a[i]), out_mul[i] = mcmcb_expmul_fast(a[i], scale) = exp(a[i]*scale) */
int mcmcb_exp_selftest(int device, const double* a, double scale, double* out_fast, double* out_mul, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* MCMCB200_H */
