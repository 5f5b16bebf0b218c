// mcmcb_main -- the reference's driver (mcmc_main, mcmc_main.F90:12-44) above the GPU C ABI.
//
//   MCMC_init      read mcmcinit.nml, run `initialize` on mcmcpar.dat / mcmccov.dat / mcmcsigma2.dat
//                  (MCMC_init.F90:16-158)                      -> mcmcbh_read_namelist, mcmcbh_initialize
//   MCMC_run*      the sampling loop                            -> mcmcb_create / set_* / run  (CUDA library)
//   MCMC_writechains  chain, sschain, s2chain + restart files (MCMC_aux.F90:17-85)  -> mcmcbh_write_matrix
//
// Run it in a directory that holds the reference's input files (e.g. a copy of testcases/): with
// nchains = 1 it writes the same files the Fortran program writes (chain 0 of the batch = the run);
// with &mcmcb nchains > 1 the stored chains beyond the first go to <name>_<chain>.<ext>.
//
// usage: mcmcb_main [directory]        (default: current directory)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mcmcb200.h"
#include "mcmcb_host.h"

static std::string in_dir(const std::string& dir, const char* f) { return (dir.empty() || f[0] == '/') ? f : dir + "/" + f; }

static std::string chain_name(const std::string& base, long long chain) {
  if (chain == 0) return base;
  const size_t dot = base.find_last_of('.');
  char suf[32];
  std::snprintf(suf, sizeof suf, "_%05lld", chain);
  return dot == std::string::npos ? base + suf : base.substr(0, dot) + suf + base.substr(dot);
}

#define HOST(call)                                                             \
  do {                                                                         \
    int rc_ = (call);                                                          \
    if (rc_ != 0) {                                                            \
      std::fprintf(stderr, "%s\n(%s -> %d)\n", mcmcbh_last_error(), #call, rc_); \
      return 1;                                                                \
    }                                                                          \
  } while (0)
#define DEV(call)                                                                                   \
  do {                                                                                              \
    int rc_ = (call);                                                                               \
    if (rc_ < 0) {                                                                                  \
      std::fprintf(stderr, "%s failed: %d %s\n", #call, rc_, h ? mcmcb_last_error(h) : "");         \
      return 2;                                                                                     \
    }                                                                                               \
  } while (0)

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : "";
  mcmcb_config cfg;
  mcmcbh_files files;
  mcmcb_handle h = nullptr;
  HOST(mcmcbh_read_namelist(in_dir(dir, "mcmcinit.nml").c_str(), &cfg, &files));
  if (cfg.nsimu <= 0) {  // mcmcinit.F90: nsimu = 0 means no run
    std::fprintf(stderr, "nsimu = 0, no MCMC run\n");
    return 0;
  }
  int npar = 0, nycol = 0, *nobs = nullptr;
  double *par0 = nullptr, *cmat0 = nullptr, *sigma2 = nullptr, *blob = nullptr;
  size_t nblob = 0;
  HOST(mcmcbh_initialize(dir.c_str(), &files, &npar, &nycol, &par0, &cmat0, &sigma2, &nobs, &cfg.initcmatn));
  HOST(mcmcbh_model_blob(cfg.model, in_dir(dir, files.datafile).c_str(), &blob, &nblob));
  if (cfg.store_chains == 0) cfg.store_chains = 1;
  if (files.verbosity > 0)
    std::printf("mcmcb_main: %lld chain(s), npar = %d, nsimu = %d, model = %s\n", cfg.nchains, npar, cfg.nsimu, cfg.model);

  DEV(mcmcb_create(&cfg, &h));
  DEV(mcmcb_set_data(h, blob, nblob));
  if (files.priorsfile[0]) {  // priorfun.f90:58-100: rows (mu, sig) per parameter
    double* pr = nullptr;
    int r = 0, c = 0;
    HOST(mcmcbh_load_dat(in_dir(dir, files.priorsfile).c_str(), &pr, &r, &c));
    if (r != npar || c < 2) { std::fprintf(stderr, "priors file must be npar x 2\n"); return 1; }
    std::vector<double> mu(npar), sg(npar);
    for (int i = 0; i < npar; i++) { mu[i] = pr[(size_t)i * c]; sg[i] = pr[(size_t)i * c + 1]; }
    DEV(mcmcb_set_priors(h, mu.data(), sg.data(), npar));
    mcmcbh_free(pr);
  }
  DEV(mcmcb_set_initial(h, npar, nycol, par0, 0, cmat0, sigma2, nobs));
  // MCMC_LOOP in pieces of printint steps (the reference prints its statistics there, MCMC_adapt.F90:22)
  int left = cfg.nsimu - 1, done = 0;
  const int piece = files.printint > 0 ? files.printint : left;
  while (left > 0 || done == 0) {
    const int n = left < piece ? left : piece;
    DEV(mcmcb_run(h, n));
    DEV(mcmcb_sync(h));
    done += n;
    left -= n;
    if (files.verbosity > 0) {
      std::vector<long long> cnt((size_t)cfg.nchains * 8);
      DEV(mcmcb_fetch(h, "counters", cnt.data(), cnt.size() * sizeof(long long)));
      double stayed = 0;
      for (long long c = 0; c < cfg.nchains; c++) stayed += (double)cnt[(size_t)c * 8];
      std::printf(" i = %d  rejected = %.1f %%\n", done + 1, 100.0 * stayed / ((double)cfg.nchains * (done > 0 ? done : 1)));
    }
    if (n == 0) break;
  }

  // ---- MCMC_writechains (MCMC_aux.F90:17-85)
  const int ld = cfg.nsimu;
  std::vector<double> chain((size_t)ld * (npar + 1)), sschain((size_t)ld * (nycol + 1)), s2chain((size_t)ld * nycol);
  const long long nstore = cfg.store_chains < 0 ? cfg.nchains : (cfg.store_chains < cfg.nchains ? cfg.store_chains : cfg.nchains);
  std::vector<double> lastpar(npar), lasts2(nycol);
  for (long long c = 0; c < nstore; c++) {
    int rows = 0;
    DEV(mcmcb_fetch_chain(h, c, ld, chain.data(), sschain.data(), s2chain.data(), &rows));
    HOST(mcmcbh_write_matrix(in_dir(dir, chain_name(files.chainfile, c).c_str()).c_str(), "chain", chain.data(), rows, npar + 1, ld));
    HOST(mcmcbh_write_matrix(in_dir(dir, chain_name(files.ssfile, c).c_str()).c_str(), "sschain", sschain.data(), rows, nycol + 1, ld));
    if (cfg.updatesigma)
      HOST(mcmcbh_write_matrix(in_dir(dir, chain_name(files.s2file, c).c_str()).c_str(), "s2chain", s2chain.data(), cfg.nsimu, nycol, ld));
    if (c == 0) {
      for (int k = 0; k < npar; k++) lastpar[k] = chain[(size_t)k * ld + rows - 1];
      for (int k = 0; k < nycol; k++) lasts2[k] = s2chain[(size_t)k * ld + cfg.nsimu - 1];
    }
  }
  // restart files from chain 0 (the pooled statistics when pool_adapt = 1): final covariance, its weight,
  // the mean, the last point, sigma2 + nobs
  double w_final = 0.0;
  {
    const size_t N = (size_t)cfg.nchains;
    std::vector<double> cm(N * npar * npar), mean(N * npar), wsum(N);
    DEV(mcmcb_fetch(h, "cmat", cm.data(), cm.size() * sizeof(double)));
    DEV(mcmcb_fetch(h, "mean", mean.data(), mean.size() * sizeof(double)));
    DEV(mcmcb_fetch(h, "wsum", wsum.data(), wsum.size() * sizeof(double)));
    double w0 = wsum[0];
    if (cfg.pool_adapt && mcmcb_pool_fetch(h, &w0, mean.data(), cm.data()) != 0) w0 = wsum[0];
    w_final = w0;
    HOST(mcmcbh_write_dat(in_dir(dir, files.covffile).c_str(), cm.data(), npar, npar, npar));
    if (files.covnfile[0]) {
      const double wn = (double)(int)w0;
      HOST(mcmcbh_write_dat(in_dir(dir, files.covnfile).c_str(), &wn, 1, 1, 1));
    }
    HOST(mcmcbh_write_dat(in_dir(dir, files.meanfile).c_str(), mean.data(), npar, 1, npar));
    HOST(mcmcbh_write_dat(in_dir(dir, files.parffile).c_str(), lastpar.data(), 1, npar, 1));  // a row, MCMC_aux.F90:59
    if (cfg.updatesigma) {
      std::vector<double> s2n((size_t)2 * nycol);  // 2 x nycol: sigma2 row, nobs row
      for (int k = 0; k < nycol; k++) { s2n[(size_t)k * 2] = lasts2[k]; s2n[(size_t)k * 2 + 1] = (double)nobs[k]; }
      HOST(mcmcbh_write_dat(in_dir(dir, files.sigma2ffile).c_str(), s2n.data(), 2, nycol, 2));
    }
  }
  // the namelist of the CONTINUATION run (MCMC_aux.F90:48-52,79-83): initcmatn = int(chainwsum) when covnfile is
  // set, then initcmatn += simuind and burnintime = 0 -- the restart weights the saved covariance with what it has seen
  if (files.nmlffile[0]) {
    mcmcb_config next = cfg;
    // the reference writes its namelist VARIABLES, i.e. after check_mcmcinit_parameters (mcmcinit.F90:235-368) and
    // after MCMC_init replaced S02 <= 0 by sigma2(1) (MCMC_init.F90:114-116)
    mcmcb_check_config(&next, nullptr, nullptr, nullptr);
    if (next.S02 <= 0.0) next.S02 = sigma2[0];
    next.initcmatn = (files.covnfile[0] ? (int)w_final : next.initcmatn) + cfg.nsimu;
    next.burnintime = 0;
    HOST(mcmcbh_write_namelist(in_dir(dir, files.nmlffile).c_str(), &next, &files));
  }
  if (cfg.diag_stride > 0) {
    std::vector<double> rhat(npar), ess(npar), pm(npar), pv(npar);
    long long ns = 0, nc = 0;
    if (mcmcb_diagnostics(h, rhat.data(), ess.data(), pm.data(), pv.data(), &ns, &nc) == 0)
      for (int k = 0; k < npar; k++)
        std::printf(" par %d: mean %.6g  var %.6g  R-hat %.4f  ESS %.0f  (%lld chains x %lld snapshots)\n", k + 1, pm[k], pv[k],
                    rhat[k], ess[k], nc, ns);
  }
  if (files.verbosity > 0)
    std::printf("note: saved results in %s and %s%s%s.\n", files.chainfile, files.ssfile, cfg.updatesigma ? " and " : "",
                cfg.updatesigma ? files.s2file : "");
  mcmcb_destroy(h);
  mcmcbh_free(par0); mcmcbh_free(cmat0); mcmcbh_free(sigma2); mcmcbh_free(nobs); mcmcbh_free(blob);
  return 0;
}
