// C ABI of the batched adaptive-MH path (include/mcmcb200.h): handle, model registry,
// kernel launchers, state fetch, streamed dumps.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <thread>
#include <vector>

#include "k1_small.cuh"
#include "k2_large.cuh"
#include "k3_scam.cuh"
#include "pool.cuh"
#include "diag.cuh"
#include "mcmcb200.h"
#include "models.cuh"
#include "registry.h"

using namespace mcmcb;

#ifndef MCMCB_K1_DEFAULT_BATCH
#define MCMCB_K1_DEFAULT_BATCH 4
#endif

#include "launchers.cuh"


// ------------------------------------------------------------------ registry
namespace mcmcb {
std::vector<ModelEntry>& registry() {
  static std::vector<ModelEntry> r;
  return r;
}
int register_model(const ModelEntry& e) {
  for (auto& x : registry())
    if (std::strcmp(x.name, e.name) == 0 && x.kernel == e.kernel && x.npar == e.npar) {
      x = e;
      return 0;
    }
  registry().push_back(e);
  return 0;
}
}  // namespace mcmcb

namespace {
using namespace mcmcb::launch;

// the built-in models register themselves from their own translation units (builtin_*.cu), exactly like a
// user plugin does (include/mcmcb200_plugin.cuh)

const ModelEntry* find_model(const char* name, int kernel, int npar = -1) {
  // kernel == 0 (auto): a model registered for both kernel families runs on the register kernel (compile-time npar)
  // when it has a registration for this npar (npar < 0: not known yet -- any), else on the warp-per-chain kernels
  for (int want : {kernel == 0 ? 1 : kernel, kernel == 0 ? 2 : kernel})
    for (auto& e : registry())
      if (std::strcmp(e.name, name) == 0 && e.kernel == want && (npar < 0 || e.npar == 0 || e.npar == npar)) return &e;
  return nullptr;
}

int pick_lanes(mcmcb_handle h) {
  int L = h->cfg.lanes_per_chain;
  if (L == 1 || L == 2 || L == 4 || L == 8 || L == 16 || L == 32) return L;
  // auto: smallest group size that still fills every SM with K1_THREADS resident threads
  long long fill = (long long)h->num_sms * K1_THREADS;
  L = 1;
  while (L < 32 && h->cfg.nchains * L < fill) L *= 2;
  return L;
}

}  // namespace

// ------------------------------------------------------------------ plugins
extern "C" int mcmcb_load_plugin(const char* path) {
  if (!path || !*path) return MCMCB_EINVAL;
  void* lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);  // the plugin's static registrars call register_model()
  if (!lib) {
    std::fprintf(stderr, "mcmcb_load_plugin: %s\n", dlerror());
    return MCMCB_ENOMODEL;
  }
  return MCMCB_OK;
}

// ------------------------------------------------------------------ config
extern "C" int mcmcb_default_config(mcmcb_config* c) {  // mcmcinit.F90:184-230
  if (!c) return MCMCB_EINVAL;
  std::memset(c, 0, sizeof *c);
  c->abi_version = MCMCB_ABI_VERSION;
  c->method = MCMCB_DRAM;
  c->nsimu = 0;
  c->doadapt = 1;
  c->doburnin = 0;
  c->burnintime = 0;
  c->badaptint = -1;
  c->greedy = 0;
  c->scalelimit = 0.05;
  c->scalefactor = 2.5;
  c->drscale = 0.0;
  c->adaptint = 100;
  c->adapthist = 0;
  c->adaptend = 0;
  c->initcmatn = 0;
  c->N0 = 1.0;
  c->S02 = 0.0;
  c->updatesigma = 1;
  c->condmax = 0.0;
  c->alphatarget = 0.234;
  c->nuparam = 0.7;
  c->nchains = 1;
  c->chain_offset = 0;
  c->seed = 0;
  c->rng_mode = MCMCB_RNG_PHILOX;
  c->device = 0;
  c->store_chains = 0;
  c->lanes_per_chain = 0;
  c->dump_stride = 0;
  c->kernel = 0;
  c->pool_adapt = 0;
  c->diag_stride = 0;
  c->diag_lags = 8;
  c->ngpus = 1;
  std::strcpy(c->model, "expreg");
  return MCMCB_OK;
}

extern "C" int mcmcb_check_config(mcmcb_config* c, int* dodr, int* doscam, int* usesvd) {  // mcmcinit.F90:235-368
  if (!c) return MCMCB_EINVAL;
  if (c->adapthist < 0) c->adapthist = 0;
  if (c->adaptint < 0) { c->adaptint = 0; c->doadapt = 0; }
  if (c->burnintime < 0) c->burnintime = 0;
  if (c->badaptint <= 0) c->badaptint = c->adaptint;
  if (c->badaptint == 0) c->doburnin = 0;
  if (c->initcmatn < 0) c->initcmatn = 0;
  if (c->scalelimit < 0.0 || c->scalelimit > 0.5) return MCMCB_EINVAL;  // reference stops, :254-257
  if (c->scalefactor < 0.0) c->scalefactor = 1.0;
  int sc = 0;
  if (c->method == MCMCB_SCAM) {
    sc = 1;
    if (c->condmax <= 0.0) c->condmax = 1.0e15;
    c->doburnin = 0;
    c->drscale = 0.0;
  }
  if (c->method == MCMCB_RAM) c->drscale = 0.0;
  if (c->method == MCMCB_ER) c->drscale = 0.0;  // "no dr with er", MCMC_run_er.F90:24-27
  if (c->method < MCMCB_DRAM || c->method > MCMCB_ER) return MCMCB_EINVAL;
  if (dodr) *dodr = c->drscale > 0.0;
  if (doscam) *doscam = sc;
  if (usesvd) *usesvd = c->condmax > 0.0;
  return MCMCB_OK;
}

// ------------------------------------------------------------------ lifecycle
// ------------------------------------------------------------------ one handle driving several GPUs
// cfg.ngpus > 1: the handle is a GROUP of per-device handles ("kids") inside this one process and host thread --
// the reference's host is a single-process program (mcmc_main.F90:12-44), so a Fortran user gets the whole box
// without MPI.  Chains are sharded by contiguous global id; every call fans out; the kernels of the
// kids run concurrently because launches are asynchronous.  The pooled-adaptation and diagnostics reductions go
// over NCCL (one communicator per device from ncclCommInitAll, calls bracketed by ncclGroupStart/End); the library
// is dlopen'ed so that a process that never asks for ngpus > 1 with pooling does not need it.
namespace grp {
struct Nccl {
  void* lib = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static Nccl& nccl() { static Nccl n; return n; }
static int nccl_load(mcmcb_handle g) {
  Nccl& n = nccl();
  if (n.lib) return 0;
  const char* env = std::getenv("MCMCB_NCCL_LIB");
  const char* cand[] = {env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"};
  for (const char* c : cand) {
    n.lib = dlopen(c, RTLD_NOW | RTLD_LOCAL);
    if (n.lib) break;
  }
  if (!n.lib) { g->err = std::string("cannot load NCCL (set MCMCB_NCCL_LIB): ") + dlerror(); return MCMCB_EUNSUPPORTED; }
  n.CommInitAll = (int (*)(void**, int, const int*))dlsym(n.lib, "ncclCommInitAll");
  n.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(n.lib, "ncclAllReduce");
  n.GroupStart = (int (*)())dlsym(n.lib, "ncclGroupStart");
  n.GroupEnd = (int (*)())dlsym(n.lib, "ncclGroupEnd");
  n.CommDestroy = (int (*)(void*))dlsym(n.lib, "ncclCommDestroy");
  n.GetErrorString = (const char* (*)(int))dlsym(n.lib, "ncclGetErrorString");
  if (!n.CommInitAll || !n.AllReduce || !n.GroupStart || !n.GroupEnd || !n.CommDestroy) {
    g->err = "NCCL library lacks the expected symbols";
    n.lib = nullptr;
    return MCMCB_EUNSUPPORTED;
  }
  return 0;
}
static int comms(mcmcb_handle g) {  // lazily: only pooled adaptation / diagnostics need them
  if (!g->nccl_comms.empty()) return 0;
  int rc = nccl_load(g);
  if (rc) return rc;
  std::vector<int> devs;
  for (auto* k : g->kids) devs.push_back(k->cfg.device);
  g->nccl_comms.assign(g->kids.size(), nullptr);
  const int e = nccl().CommInitAll(g->nccl_comms.data(), (int)devs.size(), devs.data());
  if (e != 0) {
    g->err = std::string("ncclCommInitAll: ") + (nccl().GetErrorString ? nccl().GetErrorString(e) : "error");
    g->nccl_comms.clear();
    return MCMCB_ECUDA;
  }
  return 0;
}
// sum-reduce n doubles at bufs[k] (device memory of kid k) in place over all kids, ordered on each kid's stream
static int allreduce(mcmcb_handle g, const std::vector<double*>& bufs, size_t n) {
  int rc = comms(g);
  if (rc) return rc;
  Nccl& nc = nccl();
  int e = nc.GroupStart();
  for (size_t k = 0; k < g->kids.size() && e == 0; k++)
    e = nc.AllReduce(bufs[k], bufs[k], n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, g->nccl_comms[k], g->kids[k]->stream);
  const int e2 = nc.GroupEnd();
  if (e == 0) e = e2;
  if (e != 0) { g->err = std::string("ncclAllReduce: ") + (nc.GetErrorString ? nc.GetErrorString(e) : "error"); return MCMCB_ECUDA; }
  g->nccl_calls++;
  return 0;
}
// run fn(kid index) for every device of a group at once (one host thread per device for the duration of the call):
// uploads, downloads and their synchronisations then overlap across devices instead of running one device after another
template <class F>
static int for_kids(mcmcb_handle g, F fn) {
  const size_t n = g->kids.size();
  std::vector<int> rc(n, 0);
  std::vector<std::thread> th;
  for (size_t k = 1; k < n; k++) th.emplace_back([&, k] { rc[k] = fn((int)k); });
  rc[0] = fn(0);
  for (auto& t : th) t.join();
  for (int r : rc)
    if (r) return r;
  return 0;
}
static void shard(long long N, int G, int k, long long* n, long long* off) {  // contiguous ranges, remainder to the first kids
  const long long base = N / G, rem = N % G;
  *n = base + (k < rem ? 1 : 0);
  *off = (long long)k * base + (k < rem ? k : rem);
}
}  // namespace grp

static int create_group(const mcmcb_config* cfg, mcmcb_handle* out) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) return MCMCB_ECUDA;
  const int G = cfg->ngpus;
  if (cfg->device < 0 || cfg->device + G > ndev || cfg->nchains < G) return MCMCB_EINVAL;
  mcmcb_handle g = new mcmcb_handle_s();
  g->cfg = *cfg;
  for (int k = 0; k < G; k++) {
    mcmcb_config c = *cfg;
    c.ngpus = 1;
    c.device = cfg->device + k;
    long long n = 0, off = 0;
    grp::shard(cfg->nchains, G, k, &n, &off);
    c.nchains = n;
    c.chain_offset = cfg->chain_offset + off;
    // stored chains are the first store_chains GLOBAL chains
    const long long want = cfg->store_chains < 0 ? cfg->nchains : cfg->store_chains;
    const long long mine = std::max<long long>(0, std::min<long long>(want - off, n));
    c.store_chains = (int)mine;
    mcmcb_handle kid = nullptr;
    const int rc = mcmcb_create(&c, &kid);
    if (rc) {
      for (auto* q : g->kids) mcmcb_destroy(q);
      delete g;
      return rc;
    }
    g->kids.push_back(kid);
  }
  g->npar = g->kids[0]->npar;
  g->nycol = g->kids[0]->nycol;
  g->model = g->kids[0]->model;
  *out = g;
  return MCMCB_OK;
}

extern "C" int mcmcb_create(const mcmcb_config* cfg, mcmcb_handle* out) {
  if (!cfg || !out || cfg->abi_version != MCMCB_ABI_VERSION || cfg->nchains < 1 || cfg->nsimu < 1) return MCMCB_EINVAL;
  if (cfg->ngpus < 0) return MCMCB_EINVAL;
  if (cfg->ngpus > 1) return create_group(cfg, out);
  mcmcb_handle h = new mcmcb_handle_s();
  h->cfg = *cfg;
  int rc = mcmcb_check_config(&h->cfg, &h->dodr, &h->doscam, &h->usesvd);
  if (rc) { delete h; return rc; }
  const mcmcb_config& c = h->cfg;
  // configurations that need the stored row history of the reference (AP)
  h->model = find_model(c.model, c.kernel);
  if (!h->model) { delete h; return MCMCB_ENOMODEL; }
  // greedy burn-in (MCMC_adapt.F90:83-101) is built for the Cholesky-factor samplers (K1, K2), not with an SVD factor
  // (condmax > 0; SCAM switches burn-in off anyway).  AP windows (:116-136) are built everywhere.
  if (h->model->kernel != 1 && c.method != MCMCB_RAM && h->usesvd && c.greedy && c.doburnin) { delete h; return MCMCB_EUNSUPPORTED; }
  if (c.pool_adapt && (c.adapthist > 1 || c.method == MCMCB_ER)) { delete h; return MCMCB_EUNSUPPORTED; }
  // SVD factor paths (SCAM, condmax > 0) live in the warp-per-chain kernels only: take the model's registration for
  // those kernels when it has one (the built-in "expreg" does)
  if (h->model->kernel == 1 && (h->doscam || h->usesvd)) {
    const ModelEntry* alt = c.kernel == 0 ? find_model(c.model, 2) : nullptr;
    if (!alt) { delete h; return MCMCB_EUNSUPPORTED; }
    h->model = alt;
  }
  // pooled adaptation replaces the chains' own factor updates at the AM ticks; the burn-in scaling branch
  // (per-chain acceptance driven) and the usesvd DR combination stay per chain and are not pooled
  if (c.pool_adapt && (c.doburnin || !c.doadapt || c.adaptint <= 0)) { delete h; return MCMCB_EUNSUPPORTED; }
  if (c.diag_stride < 0 || c.diag_lags < 0 || c.diag_lags > MCMCB_DIAG_MAXLAGS) { delete h; return MCMCB_EINVAL; }
  DevCfg& d = h->dc;
  d.pool = c.pool_adapt ? 1 : 0;
  d.method = c.method; d.nsimu = c.nsimu; d.doadapt = c.doadapt; d.adaptint = c.adaptint; d.adapthist = c.adapthist;
  d.adaptend = c.adaptend; d.initcmatn = c.initcmatn; d.doburnin = c.doburnin; d.burnintime = c.burnintime;
  d.badaptint = c.badaptint; d.greedy = c.greedy; d.updatesigma = c.updatesigma; d.dodr = h->dodr;
  d.doscam = h->doscam; d.usesvd = h->usesvd; d.scalelimit = c.scalelimit; d.scalefactor = c.scalefactor;
  d.drscale = c.drscale; d.condmax = c.condmax; d.N0 = c.N0; d.S02 = c.S02; d.alphatarget = c.alphatarget;
  d.nuparam = c.nuparam;
  if (cudaSetDevice(c.device) != cudaSuccess) { delete h; return MCMCB_ECUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, c.device) != cudaSuccess) { delete h; return MCMCB_ECUDA; }
  h->num_sms = prop.multiProcessorCount;
  h->max_smem = (size_t)prop.sharedMemPerBlockOptin;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return MCMCB_ECUDA; }
  if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return MCMCB_ECUDA; }
  h->store_chains = c.store_chains < 0 ? (int)std::min<long long>(c.nchains, 1 << 30) : (int)std::min<long long>(c.store_chains, c.nchains);
  // npar is fixed by the model only if every registration of this name a later mcmcb_set_initial may pick agrees on it
  // (a name registered for several compile-time npar, or for the run-time-npar kernels too: known at set_initial)
  h->npar = h->model->npar;
  for (auto& e : registry())
    if (std::strcmp(e.name, c.model) == 0 && (c.kernel == 0 || e.kernel == c.kernel) && e.npar != h->npar) h->npar = 0;
  h->nycol = h->model->ny;
  *out = h;
  return MCMCB_OK;
}

static void free_dev(mcmcb_handle h) {
  void* ptrs[] = {h->d_st, h->d_ist, h->d_par0, h->d_cmat0, h->d_sigma2, h->d_nobs, h->d_blob, h->d_prior,
                  h->d_inj, h->d_store_rows, h->d_store_cnt, h->d_store_s2, h->d_hist, h->d_tile, h->d_theta, h->d_mean, h->d_Rm,
                  h->d_cmat, h->d_gcm, h->d_gmean, h->d_gw, h->d_rowbuf, h->d_coef, h->d_Rp, h->d_scratch, h->d_cmat0_full, h->d_qstd,
                  h->d_pool, h->d_pool_partial, h->d_Rpool, h->d_fail, h->d_diag, h->d_diag_buf, h->d_diag_partial,
                  h->d_fetch};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (auto& s : h->dump_slots) {
    if (s.dev) cudaFree(s.dev);
    if (s.host) cudaFreeHost(s.host);
    if (s.ev) cudaEventDestroy(s.ev);
    if (s.ready) cudaEventDestroy(s.ready);
  }
}

extern "C" int mcmcb_destroy(mcmcb_handle h) {
  if (!h) return MCMCB_EINVAL;
  if (!h->kids.empty()) {
    for (size_t k = 0; k < h->kids.size(); k++) {
      if (k < h->nccl_comms.size() && h->nccl_comms[k]) { cudaSetDevice(h->kids[k]->cfg.device); grp::nccl().CommDestroy(h->nccl_comms[k]); }
      mcmcb_destroy(h->kids[k]);
    }
    delete h;
    return MCMCB_OK;
  }
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->copy_stream);
  free_dev(h);
  cudaStreamDestroy(h->stream);
  cudaStreamDestroy(h->copy_stream);
  delete h;
  return MCMCB_OK;
}

extern "C" const char* mcmcb_last_error(mcmcb_handle h) {
  if (!h) return "null handle";
  if (h->err.empty())
    for (auto* k : h->kids)
      if (!k->err.empty()) return k->err.c_str();
  return h->err.c_str();
}

extern "C" int mcmcb_set_data(mcmcb_handle h, const double* blob, size_t n) {
  if (!h || !blob || n == 0) return MCMCB_EINVAL;
  if (!h->kids.empty()) return grp::for_kids(h, [&](int k) { return mcmcb_set_data(h->kids[k], blob, n); });
  CK(cudaSetDevice(h->cfg.device));
  size_t bytes = ((n * sizeof(double) + 15) / 16) * 16;
  if (h->d_blob && bytes != h->blob_bytes) { cudaFree(h->d_blob); h->d_blob = nullptr; }
  if (!h->d_blob) CK(cudaMalloc(&h->d_blob, bytes));
  CK(cudaMemsetAsync(h->d_blob, 0, bytes, h->stream));
  CK(cudaMemcpyAsync(h->d_blob, blob, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->blob_n = n;
  h->blob_bytes = bytes;
  // TMA-stage the blob into shared memory when it fits beside the kernel's static smem
  h->smem_blob = bytes + MCMCB_EXP_TAB_DOUBLES * sizeof(double) + 1024 <= h->max_smem;
  h->attr_set = false;
  return MCMCB_OK;
}

extern "C" int mcmcb_set_priors(mcmcb_handle h, const double* mu, const double* sig, int npar) {
  if (!h) return MCMCB_EINVAL;
  if (!h->kids.empty()) {
    for (auto* k : h->kids) { const int rc = mcmcb_set_priors(k, mu, sig, npar); if (rc) return rc; }
    return MCMCB_OK;
  }
  CK(cudaSetDevice(h->cfg.device));
  if (h->d_prior) { cudaFree(h->d_prior); h->d_prior = nullptr; }
  if (!mu || !sig) return MCMCB_OK;
  if (npar != h->npar && h->npar > 0) return MCMCB_EINVAL;
  std::vector<double> buf(2 * (size_t)npar);
  std::copy(mu, mu + npar, buf.begin());
  std::copy(sig, sig + npar, buf.begin() + npar);
  CK(cudaMalloc(&h->d_prior, sizeof(double) * buf.size()));
  CK(cudaMemcpy(h->d_prior, buf.data(), sizeof(double) * buf.size(), cudaMemcpyHostToDevice));
  return MCMCB_OK;
}

extern "C" int mcmcb_set_initial(mcmcb_handle h, int npar, int nycol, const double* par0, long long par0_stride,
                                 const double* cmat0, const double* sigma2, const int* nobs) {
  if (!h || !par0 || !cmat0 || !sigma2 || !nobs || npar < 1 || nycol < 1) return MCMCB_EINVAL;
  if (!h->kids.empty()) {
    const int rc = grp::for_kids(h, [&](int i) {
      mcmcb_handle k = h->kids[i];
      const double* p0 = par0 + (par0_stride ? (size_t)(k->cfg.chain_offset - h->cfg.chain_offset) * par0_stride : 0);
      return mcmcb_set_initial(k, npar, nycol, p0, par0_stride, cmat0, sigma2, nobs);
    });
    if (rc) return rc;
    h->npar = npar; h->nycol = nycol; h->initial_set = true;
    return MCMCB_OK;
  }
  if (!h->d_st) {
    // now that npar is known: the register kernel if the model has a compile-time registration for it (and the
    // sampler needs no SVD factor), else its run-time-npar registration for the warp-per-chain kernels
    const int want = (h->doscam || h->usesvd) ? (h->cfg.kernel == 1 ? 1 : 2) : h->cfg.kernel;
    const ModelEntry* m = find_model(h->cfg.model, want, npar);
    if (!m) return h->model->npar > 0 && npar != h->model->npar ? MCMCB_EINVAL : MCMCB_EUNSUPPORTED;
    if (m->kernel != 1 && h->cfg.method != MCMCB_RAM && h->usesvd && h->cfg.greedy && h->cfg.doburnin) return MCMCB_EUNSUPPORTED;
    h->model = m;
  }
  if (h->model->npar > 0 && npar != h->model->npar) return MCMCB_EINVAL;
  if (nycol != h->model->ny) return MCMCB_EINVAL;
  if (par0_stride != 0 && par0_stride < npar) return MCMCB_EINVAL;
  CK(cudaSetDevice(h->cfg.device));
  h->npar = npar;
  h->nycol = nycol;
  const long long N = h->cfg.nchains;
  if (!h->d_par0) CK(cudaMalloc(&h->d_par0, sizeof(double) * (size_t)N * npar));
  if (par0_stride == npar) {
    CK(cudaMemcpyAsync(h->d_par0, par0, sizeof(double) * (size_t)N * npar, cudaMemcpyHostToDevice, h->stream));
  } else {
    h->h_tmp.resize((size_t)N * npar);
    for (long long c = 0; c < N; c++)
      for (int k = 0; k < npar; k++) h->h_tmp[(size_t)c * npar + k] = par0[(size_t)c * par0_stride + k];
    CK(cudaMemcpyAsync(h->d_par0, h->h_tmp.data(), sizeof(double) * (size_t)N * npar, cudaMemcpyHostToDevice, h->stream));
  }
  // packed upper triangle of cmat0 (column-major input; the upper part is authoritative, matutils.F90:73-75)
  int T = npar * (npar + 1) / 2;
  std::vector<double> pkd((size_t)T);
  for (int j = 0; j < npar; j++)
    for (int i = 0; i <= j; i++) pkd[(size_t)j * (j + 1) / 2 + i] = cmat0[(size_t)j * npar + i];
  if (!h->d_cmat0) CK(cudaMalloc(&h->d_cmat0, sizeof(double) * T));
  if (!h->d_sigma2) CK(cudaMalloc(&h->d_sigma2, sizeof(double) * nycol));
  if (!h->d_nobs) CK(cudaMalloc(&h->d_nobs, sizeof(int) * nycol));
  if (!h->d_tile) CK(cudaMalloc(&h->d_tile, sizeof(unsigned) * 4));
  if (!h->d_cmat0_full) CK(cudaMalloc(&h->d_cmat0_full, sizeof(double) * (size_t)npar * npar));
  {  // full symmetric copy built from the authoritative upper triangle
    std::vector<double> full((size_t)npar * npar);
    for (int j = 0; j < npar; j++)
      for (int i = 0; i < npar; i++) full[(size_t)j * npar + i] = (i <= j) ? cmat0[(size_t)j * npar + i] : cmat0[(size_t)i * npar + j];
    CK(cudaMemcpyAsync(h->d_cmat0_full, full.data(), sizeof(double) * full.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  CK(cudaMemcpyAsync(h->d_cmat0, pkd.data(), sizeof(double) * T, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_sigma2, sigma2, sizeof(double) * nycol, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_nobs, nobs, sizeof(int) * nycol, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->dc.S02 <= 0.0) h->dc.S02 = sigma2[0];  // MCMC_init.F90:114-116
  if (!h->d_st) {
    int rc = h->model->alloc(h);
    if (rc) return rc;
  }
  if (h->cfg.pool_adapt && !h->d_pool) {
    const size_t nd = (size_t)npar * npar, pn = 1 + (size_t)npar + nd;
    CK(cudaMalloc(&h->d_pool, sizeof(double) * pn));
    CK(cudaMemsetAsync(h->d_pool, 0, sizeof(double) * pn, h->stream));
    CK(cudaMalloc(&h->d_pool_partial, sizeof(double) * (size_t)POOL_BLOCKS * nd));
    CK(cudaMalloc(&h->d_Rpool, sizeof(double) * nd));
    CK(cudaMalloc(&h->d_fail, sizeof(int)));
    CK(cudaMemsetAsync(h->d_fail, 0, sizeof(int), h->stream));
  }
  if (h->cfg.diag_stride > 0 && !h->d_diag) {
    h->diag_K = h->cfg.diag_lags;
    const size_t ne = (size_t)N * npar, nf = 3 + 3 * (size_t)h->diag_K, nv = 2 + (size_t)h->diag_K;
    CK(cudaMalloc(&h->d_diag, sizeof(double) * nf * ne));
    CK(cudaMalloc(&h->d_diag_buf, sizeof(double) * (1 + (size_t)npar + nv * npar)));
    CK(cudaMalloc(&h->d_diag_partial, sizeof(double) * (size_t)POOL_BLOCKS * nv * npar));
  }
  h->diag_n = 0;
  h->L = (h->model->kernel == 1) ? pick_lanes(h) : 32;
  // chains per thread of the register kernel (thread-per-chain mapping): the launcher's tile plan falls back to
  // smaller tiles when there are too few chains to fill every warp (MCMCB_K1_BATCH overrides, tuning only)
  h->k1_batch = 1;
  if (const char* e = std::getenv("MCMCB_EXP_DIRECT")) h->k1_exp_direct = e[0] != '0';  // tuning experiments only
  if (const char* e = std::getenv("MCMCB_K1_SUPERTILE")) h->k1_supertile = e[0] != '0';
  h->k1_threads = K1_THREADS;
  {
    const char* e = std::getenv("MCMCB_ER_EXIT");
    h->er_exit = h->cfg.method == MCMCB_ER && e && e[0] == '1';
  }
  if (h->model->kernel == 1 && h->L == 1 && !h->er_exit) {
    // several chains per thread pay when the sweep over the model's data dominates a step (C3, 10^4 data: +7 %); with a
    // short data loop the step is the cold routines, whose state then thrashes L1/L2 (11 data: 1.9e9 chain-steps/s
    // with one chain per thread, 1.1e9 with four) -- the blob size is the proxy for the length of the data loop
    int want = h->blob_n >= 2048 ? MCMCB_K1_DEFAULT_BATCH : 1;
    if (const char* e = std::getenv("MCMCB_K1_BATCH")) want = std::atoi(e);
    h->k1_batch = (want == 2 || want == 4) ? want : 1;
  }
  // CTA size of the register kernel.  With four chains per thread the chains in flight keep ~1.3 KB of local-memory state
  // per thread; 384 threads per CTA run as fast as 512 (8.85e7 against 8.87e7 chain-steps/s on BASELINE C3) and move
  // 36 GB instead of 65 GB through HBM per launch (profiles/r02_ab_k1_block.txt); 256: 12 GB, 2.5 % slower.
  if (h->k1_batch == 4 && K1_THREADS >= 384) h->k1_threads = 384;
  if (const char* e = std::getenv("MCMCB_K1_BLOCK")) {
    const int t = std::atoi(e);
    if (t >= 32 && t <= K1_THREADS && t % 32 == 0) { h->k1_threads = t; h->k1_threads_fixed = true; }
  }
  h->k1_threads_used = h->k1_threads;
  int rc = h->model->init(h);
  if (rc) return rc;
  h->initial_set = true;
  h->steps_done = 0;
  return MCMCB_OK;
}

extern "C" int mcmcb_inject_uniforms(mcmcb_handle h, const double* u, size_t per_chain) {
  if (!h) return MCMCB_EINVAL;
  if (!h->kids.empty()) {
    for (auto* k : h->kids) {
      const double* uk = u ? u + (size_t)(k->cfg.chain_offset - h->cfg.chain_offset) * per_chain : nullptr;
      const int rc = mcmcb_inject_uniforms(k, uk, per_chain);
      if (rc) return rc;
    }
    return MCMCB_OK;
  }
  CK(cudaSetDevice(h->cfg.device));
  if (h->d_inj) { cudaFree(h->d_inj); h->d_inj = nullptr; }
  h->inj_per_chain = 0;
  if (!u || per_chain == 0) return MCMCB_OK;
  size_t n = (size_t)h->cfg.nchains * per_chain;
  CK(cudaMalloc(&h->d_inj, sizeof(double) * n));
  CK(cudaMemcpy(h->d_inj, u, sizeof(double) * n, cudaMemcpyHostToDevice));
  h->inj_per_chain = per_chain;
  return MCMCB_OK;
}

// ------------------------------------------------------------------ streamed dumps
// What a snapshot holds: theta, ss, sspri and sigma2 of every chain at one step -- the arguments of the reference's
// per-step hooks (MCMC_dump(oldpar), MCMC_userfun; MCMC_dump.F90:12-30, MCMC_userfun.F90:12-19) plus the two values
// MCMC_savechain records beside them (MCMC_aux.F90:166-185), thinned to every dump_stride-th step.
static size_t dump_scalar_fields(mcmcb_handle h) { return 2 * (size_t)h->nycol + 1; }  // ss[ny], sspri, sigma2[ny]
static size_t dump_doubles(mcmcb_handle h) {
  if (h->model->kernel == 2) return (size_t)h->cfg.nchains * h->dp + dump_scalar_fields(h) * h->pitch;
  return ((size_t)h->npar + dump_scalar_fields(h)) * h->pitch;  // K1: theta, ss, sspri, sigma2 are the first SoA fields
}

static int dump_enqueue(mcmcb_handle h) {
  // snapshot device->device on the compute stream, then device->pinned host on the copy stream, so the next
  // launch overlaps the PCIe copy
  const bool k2 = h->model->kernel == 2;
  const size_t bytes = sizeof(double) * dump_doubles(h);
  if (h->dump_slots.empty()) {
    // ring of snapshots in flight between the copy stream and the host's mcmcb_dump_pop: 4..16 slots, about 1 GB of pinned
    // host memory at most (a full ring overwrites its oldest snapshot and counts it as dropped)
    h->dump_slots.resize((size_t)std::max<long long>(4, std::min<long long>(16, (1ll << 30) / (long long)std::max<size_t>(bytes, 1))));
    for (auto& s : h->dump_slots) {
      CK(cudaMalloc(&s.dev, bytes));
      CK(cudaMallocHost(&s.host, bytes));
      CK(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
      s.state = 0;
    }
  }
  int k = h->dump_head % (int)h->dump_slots.size();
  auto& s = h->dump_slots[k];
  if (s.state != 0) {  // ring full: oldest snapshot is overwritten
    CK(cudaEventSynchronize(s.ev));
    for (auto it = h->dump_fifo.begin(); it != h->dump_fifo.end(); ++it)
      if (*it == k) { h->dump_fifo.erase(it); break; }
    h->dumps_dropped++;
  }
  if (k2) {
    const size_t tb = sizeof(double) * (size_t)h->cfg.nchains * h->dp;
    CK(cudaMemcpyAsync(s.dev, h->d_theta, tb, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(s.dev + (size_t)h->cfg.nchains * h->dp, h->d_st /* ss, sspri, sigma2 are fields 0.. */,
                       bytes - tb, cudaMemcpyDeviceToDevice, h->stream));
  } else {
    CK(cudaMemcpyAsync(s.dev, h->d_st, bytes, cudaMemcpyDeviceToDevice, h->stream));
  }
  CK(cudaEventRecord(s.ready, h->stream));
  CK(cudaStreamWaitEvent(h->copy_stream, s.ready, 0));
  CK(cudaMemcpyAsync(s.host, s.dev, bytes, cudaMemcpyDeviceToHost, h->copy_stream));
  CK(cudaEventRecord(s.ev, h->copy_stream));
  s.state = 1;
  s.step = 1 + (int)h->steps_done;
  h->dump_fifo.push_back(k);
  h->dump_head++;
  return 0;
}

static bool dump_front_ready(mcmcb_handle h) {
  if (h->dump_fifo.empty()) return false;
  return cudaEventQuery(h->dump_slots[h->dump_fifo.front()].ev) == cudaSuccess;
}

// one handle: copy the front snapshot out (caller checked dump_front_ready)
static void dump_take(mcmcb_handle h, double* par, double* ss, double* sigma2, int* step) {
  const int k = h->dump_fifo.front();
  auto& s = h->dump_slots[k];
  const long long N = h->cfg.nchains;
  const int d = h->npar, ny = h->nycol;
  const bool k2 = h->model->kernel == 2;
  const double* sc = k2 ? s.host + (size_t)N * h->dp : s.host + (size_t)d * h->pitch;  // [ss.., sspri, sigma2..][pitch]
  if (par)
    for (int f = 0; f < d; f++)
      for (long long c = 0; c < N; c++) par[(size_t)c * d + f] = k2 ? s.host[(size_t)c * h->dp + f] : s.host[(size_t)f * h->pitch + c];
  for (int f = 0; f < ny; f++)
    for (long long c = 0; c < N; c++) {
      if (ss) ss[(size_t)c * ny + f] = sc[(size_t)f * h->pitch + c];
      if (sigma2) sigma2[(size_t)c * ny + f] = sc[(size_t)(ny + 1 + f) * h->pitch + c];
    }
  if (step) *step = s.step;
  s.state = 0;
  h->dump_fifo.pop_front();
}

extern "C" int mcmcb_dump_pop_ex(mcmcb_handle h, double* par, double* ss, double* sigma2, size_t par_bytes, int* step) {
  if (!h || !par) return MCMCB_EINVAL;
  if (par_bytes < sizeof(double) * (size_t)h->cfg.nchains * h->npar) return MCMCB_EINVAL;
  if (!h->kids.empty()) {  // a snapshot is complete when every device has delivered its part
    for (auto* k : h->kids)
      if (!dump_front_ready(k)) return 0;
    for (auto* k : h->kids) {
      const size_t off = (size_t)(k->cfg.chain_offset - h->cfg.chain_offset);
      dump_take(k, par + off * h->npar, ss ? ss + off * h->nycol : nullptr, sigma2 ? sigma2 + off * h->nycol : nullptr, step);
    }
    return 1;
  }
  if (!dump_front_ready(h)) return 0;
  dump_take(h, par, ss, sigma2, step);
  return 1;
}

extern "C" int mcmcb_dump_pop(mcmcb_handle h, double* out, size_t out_bytes, int* step) {
  return mcmcb_dump_pop_ex(h, out, nullptr, nullptr, out_bytes, step);
}

// ------------------------------------------------------------------ pooled adaptation / diagnostics
// `hs` is the handle itself, or the kids of a group handle: the reductions below run over all of them (NCCL inside
// the process) and, for a plain handle, over every rank the host's allreduce callback spans.
static int allreduce_many(mcmcb_handle top, const std::vector<mcmcb_handle>& hs, const std::vector<double*>& bufs, size_t n) {
  if (hs.size() > 1) return grp::allreduce(top, bufs, n);
  mcmcb_handle h = hs[0];
  if (!h->ar_fn) return 0;  // single handle: the local sums are the global sums
  const int rc = h->ar_fn(h->ar_user, bufs[0], n, (void*)h->stream);
  if (rc) { h->err = "allreduce callback failed"; return MCMCB_ECUDA; }
  return 0;
}

// step index i (= simuind after the step) at which the pooled factor is rebuilt: the AM branch of
// MCMC_adapt (MCMC_adapt.F90:105) for DRAM/AM/SCAM, every adaptint steps for RAM
static bool is_pool_tick(const mcmcb_config& c, long long i) {
  if (!c.pool_adapt || c.adaptint <= 0 || i % c.adaptint != 0) return false;
  if (c.adaptend > 0 && i > c.adaptend) return false;
  if (c.method == MCMCB_RAM) return true;
  return i >= (long long)c.burnintime + c.adaptint + c.adapthist;
}

static int pool_tick(mcmcb_handle top, const std::vector<mcmcb_handle>& hs) {
  const size_t d = (size_t)hs[0]->npar;
  std::vector<double*> b1, b2;
  for (auto* h : hs) { b1.push_back(h->d_pool); b2.push_back(h->d_pool + 1 + d); }
  int rc = 0;
  for (auto* h : hs) { if (!rc) { CK(cudaSetDevice(h->cfg.device)); rc = h->model->pool(h, 1); } }
  if (!rc) rc = allreduce_many(top, hs, b1, 1 + d);
  for (auto* h : hs) { if (!rc) { CK(cudaSetDevice(h->cfg.device)); rc = h->model->pool(h, 2); } }
  if (!rc) rc = allreduce_many(top, hs, b2, d * d);
  for (auto* h : hs) { if (!rc) { CK(cudaSetDevice(h->cfg.device)); rc = h->model->pool(h, 3); } h->pool_ticks++; }
  return rc;
}

static DiagParams diag_params(mcmcb_handle h) {
  DiagParams p{};
  if (h->model->kernel == 2) { p.theta = h->d_theta; p.chain_stride = h->dp; p.comp_stride = 1; }
  else { p.theta = h->d_st; p.chain_stride = 1; p.comp_stride = h->pitch; }  // theta = fields 0..npar-1 of the SoA state
  p.nchains = h->cfg.nchains;
  p.d = h->npar;
  p.K = h->diag_K;
  p.nsnap = h->diag_n;
  p.ds = h->d_diag;
  return p;
}

static int diag_snapshot(mcmcb_handle h) {
  DiagParams p = diag_params(h);
  const long long ne = p.nchains * p.d;
  diag_update_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, h->stream>>>(p);
  h->launches++;
  h->diag_n++;
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ run
static int run_many(mcmcb_handle top, const std::vector<mcmcb_handle>& hs, int nsteps) {
  for (auto* h : hs) {
    if (!h->initial_set || !h->d_blob) return MCMCB_EINVAL;
    if (h->cfg.rng_mode == MCMCB_RNG_INJECTED && !h->d_inj) return MCMCB_EINVAL;
  }
  const mcmcb_config& c = hs[0]->cfg;  // the kids of a group differ in nchains / chain_offset / device only
  int left = nsteps;
  do {
    // a launch ends where the host has something to do: streamed dump, diagnostics snapshot, pooled tick.
    // simuind after k more steps = 1 + steps_done + k (the first launch also evaluates the initial point)
    const long long done = hs[0]->steps_done;
    int n = left;
    if (c.dump_stride > 0) n = std::min<long long>(n, c.dump_stride - done % c.dump_stride);
    if (c.diag_stride > 0) n = std::min<long long>(n, c.diag_stride - done % c.diag_stride);
    if (c.pool_adapt) n = std::min<long long>(n, c.adaptint - (1 + done) % c.adaptint);
    for (auto* h : hs) {  // asynchronous launches: the devices of a group run side by side
      CK(cudaSetDevice(h->cfg.device));
      const int rc = h->model->step(h, n);
      if (rc) return rc;
      h->steps_done += n;
    }
    left -= n;
    if (n > 0 && is_pool_tick(c, 1 + hs[0]->steps_done)) {
      const int rc = pool_tick(top, hs);
      if (rc) return rc;
    }
    for (auto* h : hs) {
      CK(cudaSetDevice(h->cfg.device));
      if (n > 0 && c.diag_stride > 0 && h->steps_done % c.diag_stride == 0) {
        const int rc = diag_snapshot(h);
        if (rc) return rc;
      }
      if (n > 0 && c.dump_stride > 0 && h->steps_done % c.dump_stride == 0) {
        const int rc = dump_enqueue(h);
        if (rc) return rc;
      }
    }
  } while (left > 0);
  return MCMCB_OK;
}

extern "C" int mcmcb_run(mcmcb_handle h, int nsteps) {
  if (!h || nsteps < 0) return MCMCB_EINVAL;
  if (!h->kids.empty()) return run_many(h, h->kids, nsteps);
  return run_many(h, std::vector<mcmcb_handle>{h}, nsteps);
}

extern "C" int mcmcb_set_allreduce(mcmcb_handle h, mcmcb_allreduce_fn fn, void* user) {
  if (!h) return MCMCB_EINVAL;
  if (!h->kids.empty()) return fn ? MCMCB_EUNSUPPORTED : MCMCB_OK;  // a group reduces over its own devices (NCCL)
  h->ar_fn = fn;
  h->ar_user = user;
  return MCMCB_OK;
}

extern "C" int mcmcb_pool_fetch(mcmcb_handle h, double* wsum, double* mean, double* cov) {
  if (h && !h->kids.empty()) return mcmcb_pool_fetch(h->kids[0], wsum, mean, cov);  // identical on every device
  if (!h || !h->d_pool) return MCMCB_EINVAL;
  CK(cudaSetDevice(h->cfg.device));
  const size_t d = (size_t)h->npar, n = 1 + d + d * d;
  std::vector<double> b(n);
  CK(cudaMemcpyAsync(b.data(), h->d_pool, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  // after a tick the S2 block holds the pooled covariance itself (k2_pool_factor_kernel divides in place;
  // the K1 path leaves the raw sums) -- normalise here so both kernels report the same thing
  const double W = b[0];
  const bool ram = h->cfg.method == MCMCB_RAM;
  const double den = (h->model->kernel == 2) ? 1.0 : (ram ? W : W - 1.0);
  if (wsum) *wsum = W;
  if (mean) for (size_t k = 0; k < d; k++) mean[k] = W > 0.0 ? b[1 + k] / W : 0.0;
  if (cov) for (size_t k = 0; k < d * d; k++) cov[k] = b[1 + d + k] / den;
  return MCMCB_OK;
}

extern "C" int mcmcb_diag_reset(mcmcb_handle h) {
  if (!h) return MCMCB_EINVAL;
  for (auto* k : h->kids) k->diag_n = 0;
  h->diag_n = 0;
  return MCMCB_OK;
}

// R-hat and ESS from the chain-summed moments (Gelman et al., BDA3 11.4-11.5; Stan's multi-chain ESS with
// Geyer's initial positive sequence, lags limited to diag_lags)
extern "C" int mcmcb_diagnostics(mcmcb_handle top, double* rhat, double* ess, double* mean, double* var, long long* nsnap,
                                 long long* nchains_total) {
  if (!top) return MCMCB_EINVAL;
  std::vector<mcmcb_handle> hs = top->kids.empty() ? std::vector<mcmcb_handle>{top} : top->kids;
  mcmcb_handle h = hs[0];
  if (!h->d_diag || h->diag_n < 2) return MCMCB_EINVAL;
  const int d = h->npar, K = h->diag_K, nv = 2 + K;
  std::vector<double*> b1, b2;
  for (auto* q : hs) { b1.push_back(q->d_diag_buf); b2.push_back(q->d_diag_buf + 1 + d); }
  dim3 grid(POOL_BLOCKS, d);
  for (auto* q : hs) {
    if (cudaSetDevice(q->cfg.device) != cudaSuccess) return MCMCB_ECUDA;
    DiagParams p = diag_params(q);
    diag_reduce_kernel<<<grid, POOL_THREADS, 0, q->stream>>>(p, 1, q->d_diag_buf, q->d_diag_partial);
    diag_final_kernel<<<(d * 2 + 255) / 256, 256, 0, q->stream>>>(q->d_diag_partial, POOL_BLOCKS, 2, d, 1, q->d_diag_buf);
  }
  CK(cudaGetLastError());
  int rc = allreduce_many(top, hs, b1, 1 + (size_t)d);
  if (rc) return rc;
  for (auto* q : hs) {
    if (cudaSetDevice(q->cfg.device) != cudaSuccess) return MCMCB_ECUDA;
    DiagParams p = diag_params(q);
    diag_reduce_kernel<<<grid, POOL_THREADS, 0, q->stream>>>(p, 2, q->d_diag_buf, q->d_diag_partial);
    diag_final_kernel<<<(d * nv + 255) / 256, 256, 0, q->stream>>>(q->d_diag_partial, POOL_BLOCKS, nv, d, 2, q->d_diag_buf + 1 + d);
    q->launches += 4;
  }
  CK(cudaGetLastError());
  rc = allreduce_many(top, hs, b2, (size_t)nv * d);
  if (rc) return rc;
  CK(cudaSetDevice(h->cfg.device));
  double* buf = h->d_diag_buf;
  std::vector<double> b(1 + (size_t)d + (size_t)nv * d);
  CK(cudaMemcpyAsync(b.data(), buf, sizeof(double) * b.size(), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const double M = b[0], n = (double)h->diag_n;
  if (nsnap) *nsnap = h->diag_n;
  if (nchains_total) *nchains_total = (long long)(M + 0.5);
  for (int k = 0; k < d; k++) {
    const double* s = &b[1 + d + (size_t)k * nv];
    const double mu = b[1 + k] / M;
    const double Wv = s[1] / M;                              // mean within-chain variance
    const double Bn = M > 1.0 ? s[0] / (M - 1.0) : 0.0;      // variance of the chain means = B / n
    const double varp = (n - 1.0) / n * Wv + Bn;             // marginal posterior variance estimate
    if (mean) mean[k] = mu;
    if (var) var[k] = varp;
    if (rhat) rhat[k] = std::sqrt(varp / Wv);
    if (ess) {
      // rho_t = 1 - (W - mean_c acov_t,c) / var+, acov scaled n/(n-1) like the within variance
      auto rho = [&](int t) {
        if (t == 0) return 1.0 - (Wv - Wv) / varp;
        return 1.0 - (Wv - (s[1 + t] / M) * n / (n - 1.0)) / varp;
      };
      const int T = (int)std::min<long long>(K, h->diag_n - 1);
      double tau = -1.0;
      for (int t = 0; t <= T; t += 2) {
        const double pair = rho(t) + (t + 1 <= T ? rho(t + 1) : 0.0);
        if (pair <= 0.0 && t > 0) break;
        tau += 2.0 * pair;
      }
      if (tau < 1.0 / std::log10(M * n + 10.0)) tau = 1.0 / std::log10(M * n + 10.0);  // Stan's cap on super-efficiency
      ess[k] = M * n / tau;
    }
  }
  return MCMCB_OK;
}

extern "C" int mcmcb_sync(mcmcb_handle h) {
  if (!h) return MCMCB_EINVAL;
  if (!h->kids.empty()) {
    for (auto* k : h->kids) { const int rc = mcmcb_sync(k); if (rc) return rc; }
    return MCMCB_OK;
  }
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->copy_stream));
  return MCMCB_OK;
}

// ------------------------------------------------------------------ fetch
extern "C" int mcmcb_fetch(mcmcb_handle h, const char* what, void* out, size_t out_bytes) {
  if (!h || !what || !out || !h->initial_set) return MCMCB_EINVAL;
  if (!h->kids.empty()) {  // every array is chain-major: the kids' parts are consecutive slices
    if (h->cfg.nchains <= 0 || out_bytes % (size_t)h->cfg.nchains != 0) return MCMCB_EINVAL;
    const size_t per = out_bytes / (size_t)h->cfg.nchains;  // bytes per chain as the caller sized it
    return grp::for_kids(h, [&](int i) {
      mcmcb_handle k = h->kids[i];
      const size_t off = (size_t)(k->cfg.chain_offset - h->cfg.chain_offset);
      return mcmcb_fetch(k, what, (char*)out + off * per, per * (size_t)k->cfg.nchains);
    });
  }
  CK(cudaSetDevice(h->cfg.device));
  return h->model->fetch(h, what, out, out_bytes);
}

extern "C" int mcmcb_fetch_chain(mcmcb_handle h, long long chain, int ld, double* chain_out, double* sschain_out,
                                 double* s2chain_out, int* nrows) {
  if (h && !h->kids.empty()) {
    for (auto* k : h->kids) {
      const long long off = k->cfg.chain_offset - h->cfg.chain_offset;
      if (chain >= off && chain < off + k->cfg.nchains) return mcmcb_fetch_chain(k, chain - off, ld, chain_out, sschain_out, s2chain_out, nrows);
    }
    return MCMCB_EINVAL;
  }
  if (!h || !h->initial_set || chain < 0 || chain >= h->store_chains) return MCMCB_EINVAL;
  CK(cudaSetDevice(h->cfg.device));
  return h->model->fetch_chain(h, chain, ld, chain_out, sschain_out, s2chain_out, nrows);
}

// One chain's adaptation state and counters with typed arguments -- what a Fortran host copies back into the module
// globals chainmean / chaincmat / chainwsum / R and the counters (mcmc.F90:28-55) without string keys.
extern "C" int mcmcb_fetch_stats(mcmcb_handle h, long long chain, double* mean, double* cmat, double* wsum, double* R,
                                 double* sigma2, long long* counters) {
  if (!h || !h->initial_set || chain < 0 || chain >= h->cfg.nchains) return MCMCB_EINVAL;
  if (!h->kids.empty()) {
    for (auto* k : h->kids) {
      const long long off = k->cfg.chain_offset - h->cfg.chain_offset;
      if (chain >= off && chain < off + k->cfg.nchains) return mcmcb_fetch_stats(k, chain - off, mean, cmat, wsum, R, sigma2, counters);
    }
    return MCMCB_EINVAL;
  }
  CK(cudaSetDevice(h->cfg.device));
  const int d = h->npar, ny = h->nycol;
  const size_t P = (size_t)h->pitch;
  // column `chain` of nf consecutive SoA fields -> contiguous host values (one strided copy)
  auto col = [&](const void* base, size_t elem, int f0, int nf, void* dst) -> cudaError_t {
    return cudaMemcpy2DAsync(dst, elem, (const char*)base + ((size_t)f0 * P + (size_t)chain) * elem, P * elem, elem, (size_t)nf,
                             cudaMemcpyDeviceToHost, h->stream);
  };
  std::vector<double> tri;
  std::vector<int> ic;
  int i_first = 0, i_nd = 0;
  if (h->model->kernel == 1) {
    const K1Layout Lo = k1_layout(d, ny);
    const int T = d * (d + 1) / 2;
    tri.assign(2 * (size_t)T, 0.0);
    if (mean) CK(col(h->d_st, 8, Lo.mean, d, mean));
    if (wsum) CK(col(h->d_st, 8, Lo.wsum, 1, wsum));
    if (sigma2) CK(col(h->d_st, 8, Lo.s2, ny, sigma2));
    if (cmat) CK(col(h->d_st, 8, Lo.cm, T, tri.data()));
    if (R) CK(col(h->d_st, 8, Lo.r, T, tri.data() + T));
    ic.assign(Lo.i_nf, 0);
    if (counters) CK(col(h->d_ist, 4, 0, Lo.i_nf, ic.data()));
    CK(cudaStreamSynchronize(h->stream));
    for (int j = 0; j < d; j++)
      for (int i = 0; i <= j; i++) {
        if (cmat) cmat[(size_t)j * d + i] = cmat[(size_t)i * d + j] = tri[(size_t)j * (j + 1) / 2 + i];
        if (R) { R[(size_t)j * d + i] = tri[T + (size_t)j * (j + 1) / 2 + i]; if (i != j) R[(size_t)i * d + j] = 0.0; }
      }
    i_first = Lo.i_stayed; i_nd = Lo.i_ndlo;
  } else {
    const K2Layout Lo = k2_layout(ny);
    std::vector<double> rm;
    if (mean) CK(cudaMemcpyAsync(mean, h->d_mean + (size_t)chain * h->dp, sizeof(double) * d, cudaMemcpyDeviceToHost, h->stream));
    if (cmat) CK(cudaMemcpyAsync(cmat, h->d_cmat + (size_t)chain * d * d, sizeof(double) * d * d, cudaMemcpyDeviceToHost, h->stream));
    if (R) {
      rm.resize((size_t)d * d);
      CK(cudaMemcpyAsync(rm.data(), h->d_Rm + (size_t)chain * h->r_stride, sizeof(double) * d * d, cudaMemcpyDeviceToHost, h->stream));
    }
    if (wsum) CK(col(h->d_st, 8, Lo.wsum, 1, wsum));
    if (sigma2) CK(col(h->d_st, 8, Lo.s2, ny, sigma2));
    ic.assign(Lo.i_nf, 0);
    if (counters) CK(col(h->d_ist, 4, 0, Lo.i_nf, ic.data()));
    CK(cudaStreamSynchronize(h->stream));
    if (R)  // Cholesky factor: row-major upper -> column-major; SVD factor is stored column-major already
      for (int j = 0; j < d; j++)
        for (int i = 0; i < d; i++)
          R[(size_t)j * d + i] = h->factor_mode == FACTOR_CHOL ? (i <= j ? rm[(size_t)i * d + j] : 0.0) : rm[(size_t)j * d + i];
    i_first = Lo.i_stayed; i_nd = Lo.i_ndlo;
  }
  if (counters) {  // stayed, bndstayed, draccepted, drtries, chainind, simuind, status are the first seven int fields
    for (int k = 0; k < 7; k++) counters[k] = ic[i_first + k];
    counters[7] = (long long)(((unsigned long long)(unsigned)ic[i_nd + 1] << 32) | (unsigned)ic[i_nd]);
  }
  return MCMCB_OK;
}

// ------------------------------------------------------------------ introspection
extern "C" void* mcmcb_stream(mcmcb_handle h) {
  if (h && !h->kids.empty()) return (void*)h->kids[0]->stream;
  return h ? (void*)h->stream : nullptr;
}
extern "C" void* mcmcb_stream_of(mcmcb_handle h, int k) {  // stream of the k-th device of a group handle
  if (!h) return nullptr;
  if (h->kids.empty()) return k == 0 ? (void*)h->stream : nullptr;
  return (k >= 0 && k < (int)h->kids.size()) ? (void*)h->kids[k]->stream : nullptr;
}
extern "C" int mcmcb_ngpus(mcmcb_handle h) { return h ? (h->kids.empty() ? 1 : (int)h->kids.size()) : 0; }
extern "C" long long mcmcb_nccl_calls(mcmcb_handle h) { return h ? h->nccl_calls : 0; }
extern "C" long long mcmcb_launch_count(mcmcb_handle h) {
  if (!h) return 0;
  long long n = h->launches;
  for (auto* k : h->kids) n += k->launches;
  return n;
}
extern "C" int mcmcb_chains_per_thread(mcmcb_handle h) {
  if (h && !h->kids.empty()) return mcmcb_chains_per_thread(h->kids[0]);
  return (h && h->model && h->model->kernel == 1) ? h->k1_batch : 1;
}

extern "C" int mcmcb_info(mcmcb_handle h, int* npar, int* nycol, int* lanes, int* kernel, int* tpb, int* blocks,
                          size_t* smem) {
  if (!h) return MCMCB_EINVAL;
  if (!h->kids.empty()) return mcmcb_info(h->kids[0], npar, nycol, lanes, kernel, tpb, blocks, smem);
  if (npar) *npar = h->npar;
  if (nycol) *nycol = h->nycol;
  if (lanes) *lanes = (h->model && h->model->kernel == 2 && h->k2_group_threads > 0) ? h->k2_group_threads : (h->k5s_lanes > 0 ? h->k5s_lanes : (h->k4 ? 1 : h->L));
  if (kernel) *kernel = h->model ? h->model->kernel : 0;
  if (tpb) *tpb = (h->model && h->model->kernel == 2) ? h->k2_warps * 32 : h->k1_threads_used;
  if (blocks) *blocks = h->blocks;
  if (smem) *smem = h->smem;
  return MCMCB_OK;
}

// FP64 pipe microbenchmark: 8 independent DFMA chains per thread
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678) out[0] = s;
}

extern "C" int mcmcb_dfma_peak(int device, double* tflops, double* ms_out) {
  if (cudaSetDevice(device) != cudaSuccess) return MCMCB_ECUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MCMCB_ECUDA;
  double* d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) return MCMCB_ECUDA;
  const int threads = 512, blocks = prop.multiProcessorCount * 4, iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dfma_peak_kernel<<<blocks, threads>>>(d, 2000, 0.999999, 1e-7);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (e != cudaSuccess) return MCMCB_ECUDA;
  double flops = 2.0 * 64.0 * (double)iters * (double)threads * (double)blocks;
  if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return MCMCB_OK;
}

// ------------------------------------------------------------------ exp self-test (accuracy evidence)
__global__ void exp_selftest_kernel(const double* a, double sc, double* out_fast, double* out_mul, long long n) {
  extern __shared__ double tab[];  // MCMCB_EXP_TAB_DOUBLES
  mcmcb_stage_exp_table(tab);
  __syncthreads();
  const unsigned tl = mcmcb_exp_column(tab);
  const double ks = mcmcb_expmul_scale(sc);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    out_fast[i] = mcmcb_exp_ok(a[i]) ? mcmcb_exp_fast(a[i], tl) : exp(a[i]);
    out_mul[i] = mcmcb_exp_ok(a[i] * sc) ? mcmcb_expmul_fast(a[i], ks, tl) : exp(a[i] * sc);
  }
}

extern "C" int mcmcb_exp_selftest(int device, const double* a, double scale, double* out_fast, double* out_mul, size_t n) {
  if (!a || !out_fast || !out_mul || n == 0) return MCMCB_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return MCMCB_ECUDA;
  double *da = nullptr, *d1 = nullptr, *d2 = nullptr;
  cudaError_t e = cudaMalloc(&da, 8 * n);
  if (e == cudaSuccess) e = cudaMalloc(&d1, 8 * n);
  if (e == cudaSuccess) e = cudaMalloc(&d2, 8 * n);
  if (e == cudaSuccess) e = cudaMemcpy(da, a, 8 * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const int tab_bytes = MCMCB_EXP_TAB_DOUBLES * sizeof(double);
    cudaFuncSetAttribute(exp_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tab_bytes);
    exp_selftest_kernel<<<296, 256, tab_bytes>>>(da, scale, d1, d2, (long long)n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out_fast, d1, 8 * n, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out_mul, d2, 8 * n, cudaMemcpyDeviceToHost);
  cudaFree(da); cudaFree(d1); cudaFree(d2);
  return e == cudaSuccess ? MCMCB_OK : MCMCB_ECUDA;
}
