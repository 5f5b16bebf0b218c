/*
 * mcmcb_host.h -- host side of the reference's driver above the C ABI of include/mcmcb200.h.
 *
 * The reference's host is Fortran (mcmc_main.F90, mcmcinit.F90, initialize.F90, MCMC_aux.F90,
 * matutils.F90, matfiles.F90); this image has no Fortran compiler, so the same steps are written in
 * C++ where the reference is compiled code: read namelist &mcmc, run `initialize` on the .dat files,
 * hand the chains to the GPU through the C ABI, write the chain / restart files in the reference's
 * formats.  Pure host code: no CUDA types, usable (and tested) without a GPU.
 */
#ifndef MCMCB_HOST_H
#define MCMCB_HOST_H

#include <stddef.h>

#include "mcmcb200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MCMCBH_PATH 256
#define MCMCBH_OK 0
#define MCMCBH_ENOFILE (-1)  /* file missing / unreadable */
#define MCMCBH_EPARSE (-2)   /* malformed namelist or data file (message in mcmcbh_last_error) */
#define MCMCBH_EINVAL (-3)

/* file names and host-only switches of namelist &mcmc (mcmcinit.F90:40-82; defaults 184-230) */
typedef struct mcmcbh_files {
  char chainfile[MCMCBH_PATH], s2file[MCMCBH_PATH], ssfile[MCMCBH_PATH], priorsfile[MCMCBH_PATH];
  char cov0file[MCMCBH_PATH], covffile[MCMCBH_PATH], covnfile[MCMCBH_PATH], meanfile[MCMCBH_PATH];
  char nmlffile[MCMCBH_PATH], parfile[MCMCBH_PATH], parffile[MCMCBH_PATH];
  char sigma2file[MCMCBH_PATH], sigma2ffile[MCMCBH_PATH];
  char datafile[MCMCBH_PATH]; /* &mcmcb: the user model's data (what its ssfunction loads, e.g. data.dat) */
  int verbosity, printint, dumpint, usrfunlen, filepars;
  /* legacy variables of the Modest interface, read and written back as they are (mcmcinit.F90:218-223) */
  int svddim, sstype;
  double condmaxini, sstrans;
} mcmcbh_files;

const char* mcmcbh_last_error(void);

/* MCMC_init_namelist + read_mcmcinit_namelist (mcmcinit.F90:184-230, 88-182): defaults, then group &mcmc of
 * `path`; group &mcmcb (optional, new) carries the batch fields of mcmcb_config (nchains, seed, model,
 * store_chains, pool_adapt, diag_stride, diag_lags, dump_stride, datafile).  Unknown keys are an error, as a
 * Fortran namelist read would make them. */
int mcmcbh_read_namelist(const char* path, mcmcb_config* cfg, mcmcbh_files* files);

/* loaddata (matutils.F90:1007-1280): ASCII matrix, whitespace/comma separated, comment lines start with one
 * of # % ! C c; returns a malloc'ed row-major matrix the caller frees with mcmcbh_free. */
int mcmcbh_load_dat(const char* path, double** data, int* rows, int* cols);
void mcmcbh_free(void* p);
/* writedata (matutils.F90:841-907): one row per line, values separated by a blank.  The reference prints
 * with the processor-dependent G0 edit descriptor; here "%.17g" (round-trip exact). x is column-major. */
int mcmcbh_write_dat(const char* path, const double* x, int rows, int cols, int ld);
/* writemat4 (matfiles.F90:66-126): MATLAB Level 1.0 (v4) MAT file, little endian doubles, column-major */
int mcmcbh_write_mat4(const char* path, const char* name, const double* x, int rows, int cols, int ld);
/* addtomat (matfiles.F90:187-293): append `cols` columns to the matrix of an existing MAT-v4 file and rewrite its
 * header -- the reference's streaming ('disk') chain output, one column per chain row (MCMC_aux.F90:141-160) */
int mcmcbh_addto_mat4(const char* path, const double* x, int rows, int cols, int ld);
/* write_mcmcinit_namelist (mcmcinit.F90:147-179; `nmlffile`, MCMC_aux.F90:82-83): &mcmc with every variable, then &mcmcb */
int mcmcbh_write_namelist(const char* path, const mcmcb_config* cfg, const mcmcbh_files* files);
/* writes `.mat` by extension like MCMC_writechains (MCMC_aux.F90:25-29), else ASCII */
int mcmcbh_write_matrix(const char* path, const char* name, const double* x, int rows, int cols, int ld);

/* `initialize` (initialize.F90:21-121): par0, cmat0, sigma2, nobs from parfile / cov0file / sigma2file /
 * mcmcnycol.dat / covnfile in directory `dir`.  Outputs are malloc'ed (free with mcmcbh_free). */
int mcmcbh_initialize(const char* dir, const mcmcbh_files* files, int* npar, int* nycol, double** par0, double** cmat0,
                      double** sigma2, int** nobs, int* initcmatn);

/* blob of the built-in models from their data file: "expreg" = two columns x y (testcases/data.dat) */
int mcmcbh_model_blob(const char* model, const char* datapath, double** blob, size_t* n);

#ifdef __cplusplus
}
#endif
#endif
