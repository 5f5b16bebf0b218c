// Kernel launchers of the sampling kernels, templated on the user model: everything a model needs to be
// registered with the library (`register_model(K1<MyModel>::entry())`).  Included by api.cu for the built-in
// models and by user plugins through include/mcmcb200_plugin.cuh -- the run-time analogue of the reference's
// link-time override of ssfunction / priorfun / checkbounds (external_inc.h:4-28).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "diag.cuh"
#include "k1_small.cuh"
#include "k2_large.cuh"
#include "k2g_group.cuh"
#include "k3_scam.cuh"
#include "k4_ram.cuh"
#include "k5_scam.cuh"
#include "k5s_scam.cuh"
#include "mcmcb200.h"
#include "pool.cuh"
#include "registry.h"

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                    \
      return MCMCB_ECUDA;                                                                             \
    }                                                                                                 \
  } while (0)

namespace mcmcb {
namespace launch {

static int fetch_fields(mcmcb_handle h, int f0, int nf, std::vector<double>& buf) {
  buf.resize((size_t)nf * h->pitch);
  CK(cudaMemcpyAsync(buf.data(), h->d_st + (size_t)f0 * h->pitch, sizeof(double) * buf.size(), cudaMemcpyDeviceToHost,
                     h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// SoA state -> the chain-major host layouts of mcmcb_fetch, transposed on the device so that the
// device->host copy is one contiguous transfer straight into the caller's buffer (pinned or not)
static __global__ void k1_gather_fields_kernel(const double* st, long long pitch, long long n, int f0, int width, double* out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  const long long c = t / width;
  const int k = (int)(t - c * width);
  out[t] = st[(size_t)(f0 + k) * pitch + c];
}
// packed upper triangle (column-packed) -> full D x D column-major per chain; sym mirrors, else zero below
static __global__ void k1_gather_tri_kernel(const double* st, long long pitch, long long n, int f0, int D, int sym, double* out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * D * D) return;
  const long long c = t / (D * D);
  const int e = (int)(t - c * D * D), j = e / D, i = e - j * D;
  double v = 0.0;
  if (i <= j) v = st[(size_t)(f0 + j * (j + 1) / 2 + i) * pitch + c];
  else if (sym) v = st[(size_t)(f0 + i * (i + 1) / 2 + j) * pitch + c];
  out[t] = v;
}
static __global__ void k1_gather_counters_kernel(const int* ist, long long pitch, long long n, K1Layout Lo, long long* out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int src[7] = {Lo.i_stayed, Lo.i_bnd, Lo.i_dracc, Lo.i_drtry, Lo.i_chainind, Lo.i_simuind, Lo.i_status};
#pragma unroll
  for (int k = 0; k < 7; k++) out[c * 8 + k] = ist[(size_t)src[k] * pitch + c];
  const unsigned lo = (unsigned)ist[(size_t)Lo.i_ndlo * pitch + c], hi = (unsigned)ist[(size_t)Lo.i_ndhi * pitch + c];
  out[c * 8 + 7] = (long long)(((unsigned long long)hi << 32) | lo);
}

static __global__ void k1_gather_int_kernel(const int* ist, long long pitch, long long n, int field, long long* out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) out[c] = ist[(size_t)field * pitch + c];
}

static int fetch_stage(mcmcb_handle h, size_t bytes) {
  if (h->fetch_bytes >= bytes) return 0;
  if (h->d_fetch) cudaFree(h->d_fetch);
  h->d_fetch = nullptr;
  h->fetch_bytes = 0;
  CK(cudaMalloc(&h->d_fetch, bytes));
  h->fetch_bytes = bytes;
  return 0;
}

static int k1_fetch(mcmcb_handle h, const char* what, void* out, size_t out_bytes) {
  const long long N = h->cfg.nchains;
  const int D = h->npar, NY = h->nycol;
  const K1Layout Lo = k1_layout(D, NY);
  std::string w(what);
  const int threads = 256;
  if (w == "counters") {
    const size_t bytes = sizeof(long long) * 8 * (size_t)N;
    if (out_bytes < bytes) return MCMCB_EINVAL;
    int rc = fetch_stage(h, bytes);
    if (rc) return rc;
    k1_gather_counters_kernel<<<(unsigned)((N + threads - 1) / threads), threads, 0, h->stream>>>(
        h->d_ist, h->pitch, N, Lo, (long long*)h->d_fetch);
    h->launches++;
    CK(cudaMemcpyAsync(out, h->d_fetch, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MCMCB_OK;
  }
  if (w == "erstayed") {  // mcmc.F90:49: steps rejected by the prior alone (method 'er'), one int64 per chain
    const size_t bytes = sizeof(long long) * (size_t)N;
    if (out_bytes < bytes) return MCMCB_EINVAL;
    int rc = fetch_stage(h, bytes);
    if (rc) return rc;
    k1_gather_int_kernel<<<(unsigned)((N + threads - 1) / threads), threads, 0, h->stream>>>(h->d_ist, h->pitch, N, Lo.i_er,
                                                                                           (long long*)h->d_fetch);
    h->launches++;
    CK(cudaMemcpyAsync(out, h->d_fetch, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MCMCB_OK;
  }
  int f0 = -1, width = 0;
  bool tri = false;
  if (w == "par") { f0 = Lo.th; width = D; }
  else if (w == "ss") { f0 = Lo.ss; width = NY; }
  else if (w == "sspri") { f0 = Lo.pri; width = 1; }
  else if (w == "sigma2") { f0 = Lo.s2; width = NY; }
  else if (w == "mean") { f0 = Lo.mean; width = D; }
  else if (w == "wsum") { f0 = Lo.wsum; width = 1; }
  else if (w == "cmat") { f0 = Lo.cm; tri = true; }
  else if (w == "R") { f0 = Lo.r; tri = true; }
  else if (w == "R2") { f0 = Lo.r2; tri = true; }
  else if (w == "iC") { f0 = Lo.ic; tri = true; }
  else return MCMCB_EINVAL;
  const size_t per = tri ? (size_t)D * D : (size_t)width;
  const size_t bytes = sizeof(double) * per * (size_t)N;
  if (out_bytes < bytes) return MCMCB_EINVAL;
  int rc = fetch_stage(h, bytes);
  if (rc) return rc;
  const long long total = (long long)per * N;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (tri)
    k1_gather_tri_kernel<<<blocks, threads, 0, h->stream>>>(h->d_st, h->pitch, N, f0, D, (w == "cmat" || w == "iC") ? 1 : 0,
                                                            (double*)h->d_fetch);
  else
    k1_gather_fields_kernel<<<blocks, threads, 0, h->stream>>>(h->d_st, h->pitch, N, f0, width, (double*)h->d_fetch);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, h->d_fetch, bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MCMCB_OK;
}

// shared by K1 and K2: rows live in the common store, (chainind, cnt, simuind) in ist at the given fields
static int store_fetch_chain(mcmcb_handle h, long long chain, int ld, double* chain_out, double* sschain_out,
                             double* s2chain_out, int* nrows, int f_chainind, int f_cnt, int f_simuind) {
  const int D = h->npar, NY = h->nycol, cap = h->cfg.nsimu;
  int iv[3];
  const int fld[3] = {f_chainind, f_cnt, f_simuind};
  for (int k = 0; k < 3; k++)
    CK(cudaMemcpyAsync(&iv[k], h->d_ist + (size_t)fld[k] * h->pitch + chain, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  int rows = std::min(iv[0], cap), cnt = iv[1], simuind = iv[2];
  if (nrows) *nrows = rows;
  if (ld < rows || (s2chain_out && ld < simuind)) return MCMCB_EINVAL;
  std::vector<double> r((size_t)rows * (D + NY)), cn((size_t)rows), s2((size_t)std::max(simuind, 1) * NY);
  if (rows > 0) {
    CK(cudaMemcpyAsync(r.data(), h->d_store_rows + (size_t)chain * cap * (D + NY), sizeof(double) * r.size(),
                       cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(cn.data(), h->d_store_cnt + (size_t)chain * cap, sizeof(double) * cn.size(), cudaMemcpyDeviceToHost,
                       h->stream));
  }
  if (s2chain_out && simuind > 0)
    CK(cudaMemcpyAsync(s2.data(), h->d_store_s2 + (size_t)chain * cap * NY, sizeof(double) * (size_t)simuind * NY,
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (rows > 0) cn[rows - 1] = (double)cnt;  // the current row's count lives in the chain state
  for (int i = 0; i < rows; i++) {
    if (chain_out) {
      for (int k = 0; k < D; k++) chain_out[(size_t)k * ld + i] = r[(size_t)i * (D + NY) + k];
      chain_out[(size_t)D * ld + i] = cn[i];
    }
    if (sschain_out) {
      for (int k = 0; k < NY; k++) sschain_out[(size_t)k * ld + i] = r[(size_t)i * (D + NY) + D + k];
      sschain_out[(size_t)NY * ld + i] = cn[i];
    }
  }
  if (s2chain_out)
    for (int i = 0; i < simuind; i++)
      for (int k = 0; k < NY; k++) s2chain_out[(size_t)k * ld + i] = s2[(size_t)i * NY + k];
  return MCMCB_OK;
}

static int k1_fetch_chain(mcmcb_handle h, long long chain, int ld, double* chain_out, double* sschain_out,
                          double* s2chain_out, int* nrows) {
  const K1Layout Lo = k1_layout(h->npar, h->nycol);
  return store_fetch_chain(h, chain, ld, chain_out, sschain_out, s2chain_out, nrows, Lo.i_chainind, Lo.i_cnt,
                           Lo.i_simuind);
}


template <class M>
struct K1 {
  static constexpr int D = M::NPAR, NY = M::NY, T = D * (D + 1) / 2;

  static K1Params params(mcmcb_handle h, int nsteps) {
    K1Params p{};
    p.c = h->dc;
    p.nchains = h->cfg.nchains;
    p.pitch = h->pitch;
    p.chain_offset = h->cfg.chain_offset;
    p.seed = h->cfg.seed;
    p.nsteps = nsteps;
    p.st = h->d_st;
    p.ist = h->d_ist;
    p.par0 = h->d_par0;
    p.cmat0 = h->d_cmat0;
    p.sigma2_0 = h->d_sigma2;
    p.nobs = h->d_nobs;
    p.blob = h->d_blob;
    p.blob_n = h->blob_n;
    p.blob_bytes = (unsigned)h->blob_bytes;
    p.prior = h->d_prior;
    p.inj = h->d_inj;
    p.inj_per_chain = h->inj_per_chain;
    p.store_chains = h->store_chains;
    p.store_rows = h->cfg.nsimu;
    p.store_rows_p = h->d_store_rows;
    p.store_cnt_p = h->d_store_cnt;
    p.store_s2_p = h->d_store_s2;
    p.tile_counter = h->d_tile;
    p.hist = h->d_hist;
    p.hist_rows = h->hist_rows;
    p.exp_c1 = MCMCB_EXP_C1L;
    p.exp_c2 = MCMCB_EXP_C2L;
    return p;
  }

  static int alloc(mcmcb_handle h) {
    constexpr K1Layout Lo = k1_layout(D, NY);
    if (h->doscam || h->usesvd) return MCMCB_EUNSUPPORTED;  // SVD factor paths live in the warp-per-chain kernels
    h->nf = Lo.nf;
    h->inf = Lo.i_nf;
    h->pitch = ((h->cfg.nchains + 31) / 32) * 32;
    CK(cudaMalloc(&h->d_st, sizeof(double) * (size_t)Lo.nf * h->pitch));
    CK(cudaMalloc(&h->d_ist, sizeof(int) * (size_t)Lo.i_nf * h->pitch));
    if (h->cfg.method != MCMCB_RAM && h->cfg.doadapt && h->cfg.adapthist > 1) {  // AP window ring (MCMC_adapt.F90:116-136)
      h->hist_rows = h->cfg.adapthist;
      CK(cudaMalloc(&h->d_hist, sizeof(double) * (size_t)h->hist_rows * (D + 1) * h->pitch));
      CK(cudaMemsetAsync(h->d_hist, 0, sizeof(double) * (size_t)h->hist_rows * (D + 1) * h->pitch, h->stream));
    }
    if (h->store_chains > 0) {
      size_t rows = (size_t)h->store_chains * h->cfg.nsimu;
      CK(cudaMalloc(&h->d_store_rows, sizeof(double) * rows * (D + NY)));
      CK(cudaMalloc(&h->d_store_cnt, sizeof(double) * rows));
      CK(cudaMalloc(&h->d_store_s2, sizeof(double) * rows * NY));
      CK(cudaMemsetAsync(h->d_store_rows, 0, sizeof(double) * rows * (D + NY), h->stream));
      CK(cudaMemsetAsync(h->d_store_cnt, 0, sizeof(double) * rows, h->stream));
      CK(cudaMemsetAsync(h->d_store_s2, 0, sizeof(double) * rows * NY, h->stream));
    }
    return 0;
  }

  static int init(mcmcb_handle h) {
    K1Params p = params(h, 0);
    int threads = 256;
    long long blocks = (h->cfg.nchains + threads - 1) / threads;
    k1_init_kernel<M><<<(unsigned)blocks, threads, 0, h->stream>>>(p);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }

  template <int L, bool SMEM, bool EREXIT = false, int B = 1>
  static int launch_LS(mcmcb_handle h, const K1Params& p) {
    auto kern = k1_step_kernel<M, L, SMEM, EREXIT, B>;
    size_t smem = MCMCB_EXP_TAB_DOUBLES * sizeof(double) + (SMEM ? h->blob_bytes : 0);
    // the shared memory left over holds the direct exp table (mcmcb_expmul_direct): up to 8192 entries = arguments down
    // to -2.77; models whose arguments leave its range use the masked 2048-entry table
    int exp_dn = 0;
    if (h->k1_exp_direct && smem + 2048 < h->max_smem) {
      exp_dn = (int)std::min<size_t>(8192, (h->max_smem - 1024 - smem) / sizeof(double)) & ~255;
      if (exp_dn < 2048) exp_dn = 0;
    }
    smem += sizeof(double) * (size_t)exp_dn;
    // the attribute belongs to the kernel function, not to the handle: another handle with a different blob may have
    // lowered it since this handle's last launch, so it is set before every launch (cheap)
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!h->attr_set) {
      int occ = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, h->k1_threads, smem));
      if (occ < 1) occ = 1;
      h->occ = occ;
      h->attr_set = true;
    }
    const long long subs = (h->cfg.nchains * L + 31) / 32;  // sub-tiles: the chains one warp owns with one chain per thread
    // tile plan for W resident warps: full rounds of B-sub-tile tiles while every warp still gets one (as one super-tile
    // per warp), then one last round of smaller tiles sized to what is left per warp.  Returns the plan's length in
    // units of one sub-tile's run time (a tile of b sub-tiles takes ~b units; a round lasts as long as its tiles).
    struct Plan { long long n4, n2, n1, wtiles, units; int rounds; };
    auto plan = [&](long long W) {
      Plan P{0, 0, 0, 0, 0, 0};
      long long rem = subs;
      if (B >= 4) {
        const long long full = rem / (4 * W);
        rem -= 4 * W * full;
        P.units += 4 * full;
        if (h->k1_supertile) { P.rounds = (int)full; P.wtiles = full > 0 ? W : 0; }
        else P.n4 = full * W;
      }
      if (B >= 4 && rem > 3 * W) { P.n4 += (rem + 3) / 4; rem = 0; P.units += 4; }
      if (B >= 4 && rem > 2 * W) { P.n2 = W; rem -= 2 * W; P.units += 2; }           // 2 + 1 per warp
      else if (B >= 2 && B < 4) { P.n2 = (rem / (2 * W)) * W; P.units += 2 * (rem / (2 * W)); rem -= 2 * P.n2; }
      if (B >= 2 && rem > W) { P.n2 += (rem + 1) / 2; rem = 0; P.units += 2; }
      P.n1 = rem;
      if (rem > 0) P.units += 1;
      return P;
    };
    // CTA size.  The default (384 threads with four chains per thread) is right when the population is many rounds
    // deep; a population of only a few sub-tiles per warp can leave most warps idle in the last round (2^20 chains
    // sharded over 8 GPUs = 131 072 per GPU: 3 units of which the last keeps 31 % of the warps busy).  With four chains
    // per thread the CTA size is therefore chosen among 6..12 warps for the best estimated use of the resident warps
    // (fewer warps per SM cost ~0.6 % each, profiles/r02_ab_k1_block.txt); MCMCB_K1_BLOCK fixes it.
    int threads = h->k1_threads;
    if (B >= 4 && !h->k1_threads_fixed) {
      double best = -1.0;
      for (int w = h->k1_threads / 32; w >= 6; w--) {
        const long long Wc = std::min<long long>((long long)h->num_sms * h->occ, (subs + w - 1) / w) * w;
        const Plan P = plan(Wc);
        const double use = (double)subs / ((double)P.units * (double)Wc) * (1.0 - 0.006 * (h->k1_threads / 32 - w));
        if (use > best + 1e-9) { best = use; threads = 32 * w; }
      }
    }
    const long long wpb = threads / 32;
    const long long need = (subs + wpb - 1) / wpb;
    long long blocks = std::min<long long>((long long)h->num_sms * h->occ, need);
    if (blocks < 1) blocks = 1;
    h->blocks = (int)blocks;
    h->smem = smem;
    h->k1_threads_used = threads;
    K1Params q = p;
    q.exp_dn = exp_dn;
    const Plan P = plan(blocks * wpb);
    q.rounds = P.rounds; q.wtiles = P.wtiles;
    q.tier[0] = P.n4; q.tier[1] = P.n2; q.tier[2] = P.n1;
    CK(cudaMemsetAsync(h->d_tile, 0, sizeof(unsigned), h->stream));
    kern<<<(unsigned)blocks, threads, smem, h->stream>>>(q);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }

  template <int L>
  static int launch_L(mcmcb_handle h, const K1Params& p) {
    if (h->smem_blob) return launch_LS<L, true>(h, p);
    return launch_LS<L, false>(h, p);
  }

  static int step(mcmcb_handle h, int nsteps) {
    K1Params p = params(h, nsteps);
    if constexpr (has_ssfunction_er<M>::value) {
      // method 'er', thread per chain, on request (MCMCB_ER_EXIT=1): the model's ssfunction_er with the warp-vote early
      // exit of the data loop.  Not the default: a warp holds 32 independent chains and leaves the loop only when all of
      // them are past their critical value, which measured no gain on the regression model (profiles/r01_summary.md M),
      // while the batched kernel below is 13 % faster.  Chains are identical either way.
      if (h->cfg.method == MCMCB_ER && h->L == 1 && h->k1_batch == 1 && h->er_exit)
        return h->smem_blob ? launch_LS<1, true, true>(h, p) : launch_LS<1, false, true>(h, p);
    }
    if constexpr (has_ssfunction_batch<M>::value) {  // thread per chain, B chains per thread (k1_step_kernel)
      if (h->L == 1 && h->k1_batch == 2) return h->smem_blob ? launch_LS<1, true, false, 2>(h, p) : launch_LS<1, false, false, 2>(h, p);
      if (h->L == 1 && h->k1_batch == 4) return h->smem_blob ? launch_LS<1, true, false, 4>(h, p) : launch_LS<1, false, false, 4>(h, p);
    }
    switch (h->L) {
      case 1: return launch_L<1>(h, p);
      case 2: return launch_L<2>(h, p);
      case 4: return launch_L<4>(h, p);
      case 8: return launch_L<8>(h, p);
      case 16: return launch_L<16>(h, p);
      default: return launch_L<32>(h, p);
    }
  }

  // pooled adaptation, pool.cuh
  static int pool(mcmcb_handle h, int phase) {
    K1Params p = params(h, 0);
    if (phase == 3) {
      const int threads = 256;
      k1_pool_apply_kernel<D, NY><<<(unsigned)((h->cfg.nchains + threads - 1) / threads), threads, 0, h->stream>>>(p, h->d_pool);
      h->launches++;
    } else {
      const int nv = phase == 1 ? 1 + D : D * D;
      double* out = phase == 1 ? h->d_pool : h->d_pool + 1 + D;
      k1_pool_moments_kernel<D, NY><<<POOL_BLOCKS, POOL_THREADS, 0, h->stream>>>(p, phase, h->d_pool, h->d_pool_partial);
      pool_final_kernel<<<(nv + 255) / 256, 256, 0, h->stream>>>(h->d_pool_partial, POOL_BLOCKS, nv, out);
      h->launches += 2;
    }
    CK(cudaGetLastError());
    return 0;
  }

  static ModelEntry entry() {
    ModelEntry e{};
    e.pool = &pool;
    e.name = M::name();
    e.kernel = 1;
    e.npar = D;
    e.ny = NY;
    e.alloc = &alloc;
    e.init = &init;
    e.step = &step;
    e.fetch = &k1_fetch;
    e.fetch_chain = &k1_fetch_chain;
    return e;
  }
};

// ------------------------------------------------------------------ K2 launcher (large npar)
static int k2_fetch(mcmcb_handle h, const char* what, void* out, size_t out_bytes) {
  const long long N = h->cfg.nchains;
  const int D = h->npar, NY = h->nycol, dp = h->dp;
  const K2Layout Lo = k2_layout(NY);
  std::string w(what);
  if (w == "counters") {
    if (out_bytes < sizeof(long long) * 8 * (size_t)N) return MCMCB_EINVAL;
    std::vector<int> ib((size_t)Lo.i_nf * h->pitch);
    CK(cudaMemcpyAsync(ib.data(), h->d_ist, sizeof(int) * ib.size(), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    long long* o = (long long*)out;
    const int src[7] = {Lo.i_stayed, Lo.i_bnd, Lo.i_dracc, Lo.i_drtry, Lo.i_chainind, Lo.i_simuind, Lo.i_status};
    for (long long c = 0; c < N; c++) {
      for (int k = 0; k < 7; k++) o[c * 8 + k] = ib[(size_t)src[k] * h->pitch + c];
      unsigned lo = (unsigned)ib[(size_t)Lo.i_ndlo * h->pitch + c], hi = (unsigned)ib[(size_t)Lo.i_ndhi * h->pitch + c];
      o[c * 8 + 7] = (long long)(((unsigned long long)hi << 32) | lo);
    }
    return MCMCB_OK;
  }
  if (w == "erstayed") {
    if (out_bytes < sizeof(long long) * (size_t)N) return MCMCB_EINVAL;
    std::vector<int> ib((size_t)N);
    CK(cudaMemcpyAsync(ib.data(), h->d_ist + (size_t)Lo.i_er * h->pitch, sizeof(int) * ib.size(), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (long long c = 0; c < N; c++) ((long long*)out)[c] = ib[(size_t)c];
    return MCMCB_OK;
  }
  double* o = (double*)out;
  std::vector<double> buf;
  if (w == "par" || w == "mean" || w == "qcovstd") {
    if (out_bytes < sizeof(double) * (size_t)D * N) return MCMCB_EINVAL;
    buf.resize((size_t)N * dp);
    const bool shared = (w == "qcovstd") && h->q_stride == 0;
    CK(cudaMemcpyAsync(buf.data(), w == "par" ? h->d_theta : (w == "mean" ? h->d_mean : h->d_qstd),
                       sizeof(double) * (shared ? (size_t)dp : buf.size()), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (shared)
      for (long long c = 1; c < N; c++) std::copy(buf.begin(), buf.begin() + dp, buf.begin() + (size_t)c * dp);
    for (long long c = 0; c < N; c++)
      for (int k = 0; k < D; k++) o[(size_t)c * D + k] = buf[(size_t)c * dp + k];
    return MCMCB_OK;
  }
  if (w == "R2" && h->factor_mode != FACTOR_CHOL) return MCMCB_EINVAL;
  if (w == "cmat" || w == "R" || w == "R2") {
    if (out_bytes < sizeof(double) * (size_t)D * D * N) return MCMCB_EINVAL;
    buf.resize((size_t)N * D * D);
    const bool shared = (w != "cmat") && h->r_stride == 0;  // pooled adaptation: one factor for every chain
    CK(cudaMemcpyAsync(buf.data(), w == "cmat" ? h->d_cmat : h->d_Rm, sizeof(double) * (shared ? (size_t)D * D : buf.size()),
                       cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (shared)
      for (long long c = 1; c < N; c++) std::copy(buf.begin(), buf.begin() + (size_t)D * D, buf.begin() + (size_t)c * D * D);
    const double sc = (w == "R2") ? 1.0 / h->dc.drscale : 1.0;
    for (long long c = 0; c < N; c++)
      for (int j = 0; j < D; j++)
        for (int i = 0; i < D; i++) {
          // cmat is symmetric; the factor is stored row-major: element (i,j) at i*D+j -> column-major output
          double v = (w == "R" && h->factor_mode != FACTOR_CHOL) ? buf[(size_t)c * D * D + (size_t)j * D + i]  // SVD factor: column-major
                                                                 : buf[(size_t)c * D * D + (size_t)i * D + j];
          o[(size_t)c * D * D + (size_t)j * D + i] = (w == "cmat") ? v : v * sc;
        }
    return MCMCB_OK;
  }
  int f0 = -1, width = 0;
  if (w == "ss") { f0 = Lo.ss; width = NY; }
  else if (w == "sspri") { f0 = Lo.pri; width = 1; }
  else if (w == "sigma2") { f0 = Lo.s2; width = NY; }
  else if (w == "wsum") { f0 = Lo.wsum; width = 1; }
  else return MCMCB_EINVAL;  // "iC" is never formed by this kernel (matrix-free DR ratio)
  if (out_bytes < sizeof(double) * (size_t)width * N) return MCMCB_EINVAL;
  int rc = fetch_fields(h, f0, width, buf);
  if (rc) return rc;
  for (int k = 0; k < width; k++)
    for (long long c = 0; c < N; c++) o[(size_t)c * width + k] = buf[(size_t)k * h->pitch + c];
  return MCMCB_OK;
}

static int k2_fetch_chain(mcmcb_handle h, long long chain, int ld, double* chain_out, double* sschain_out,
                          double* s2chain_out, int* nrows) {
  const K2Layout Lo = k2_layout(h->nycol);
  return store_fetch_chain(h, chain, ld, chain_out, sschain_out, s2chain_out, nrows, Lo.i_chainind, Lo.i_cnt,
                           Lo.i_simuind);
}

template <class M>
struct K2 {
  static constexpr int NY = M::NY;

  static K2Params params(mcmcb_handle h, int nsteps) {
    K2Params p{};
    p.c = h->dc;
    p.nchains = h->cfg.nchains;
    p.pitch = h->pitch;
    p.chain_offset = h->cfg.chain_offset;
    p.seed = h->cfg.seed;
    p.nsteps = nsteps;
    p.d = h->npar;
    p.dp = h->dp;
    p.st = h->d_st;
    p.ist = h->d_ist;
    p.theta = h->d_theta;
    p.mean = h->d_mean;
    p.Rm = h->d_Rm;
    p.cmat = h->d_cmat;
    p.rowbuf = h->d_rowbuf;
    p.rowcap = h->rowcap;
    p.coef = h->d_coef;
    p.par0 = h->d_par0;
    p.cmat0 = h->d_cmat0_full;
    p.sigma2_0 = h->d_sigma2;
    p.nobs = h->d_nobs;
    p.blob = h->d_blob;
    p.blob_n = h->blob_n;
    p.blob_bytes = (unsigned)h->blob_bytes;
    p.prior = h->d_prior;
    p.inj = h->d_inj;
    p.inj_per_chain = h->inj_per_chain;
    p.store_chains = h->store_chains;
    p.store_rows = h->cfg.nsimu;
    p.store_rows_p = h->d_store_rows;
    p.store_cnt_p = h->d_store_cnt;
    p.store_s2_p = h->d_store_s2;
    p.tile_counter = h->d_tile;
    p.tick_i = 0;
    p.absorbed = 0;
    p.qstd = h->d_qstd;
    p.factor_mode = h->factor_mode;
    p.r_stride = h->r_stride;
    p.q_stride = h->q_stride;
    p.r_resident = h->r_resident ? 1 : 0;
    p.gcm = h->d_gcm;
    p.gmean = h->d_gmean;
    p.gw = h->d_gw;
    return p;
  }

  static int alloc(mcmcb_handle h) {
    static_assert(NY == 1, "the large-npar kernel supports nycol = 1");
    const K2Layout Lo = k2_layout(NY);
    const int d = h->npar;
    if (d > 32 * K2_MAXM) return MCMCB_EUNSUPPORTED;
    const long long N = h->cfg.nchains;
    const mcmcb_config& c = h->cfg;
    h->factor_mode = h->doscam ? FACTOR_SCAM : (h->usesvd ? FACTOR_SVD : FACTOR_CHOL);
    // the SVD square root is a general matrix: no rank-1 Cholesky updates (RAM) on it, and the reference's
    // second-stage ratio with usesvd inverts its upper triangle as if it were a Cholesky factor
    // (MCMC_adapt.F90:216-219) -- not reproduced
    if (h->factor_mode == FACTOR_SVD && (h->cfg.method == MCMCB_RAM || h->dodr)) return MCMCB_EUNSUPPORTED;
    h->nf = Lo.nf;
    h->inf = Lo.i_nf;
    h->pitch = ((N + 31) / 32) * 32;
    h->dp = ((d + 31) / 32) * 32;
    h->rowcap = c.burnintime + 2 * std::max(c.adaptint, 1) + c.adapthist + 2;
    const bool greedy = c.greedy && c.doburnin && c.method != MCMCB_RAM && h->factor_mode == FACTOR_CHOL;
    // plain AM (no burn-in branch, no AP window): every adaptint-th step is a tick that empties the buffer, so at most
    // adaptint rows are logged between two ticks (at d = 200 a row is 1.6 KB: 2^18 SCAM chains keep 43 GB instead of 85)
    if (!c.doburnin && c.burnintime == 0 && c.adapthist <= 1 && c.adaptint > 0 && (c.badaptint <= 0 || c.badaptint == c.adaptint))
      h->rowcap = c.adaptint + 2;
    if (c.method == MCMCB_RAM || (!c.doadapt && !greedy)) h->rowcap = 1;
    CK(cudaMalloc(&h->d_st, sizeof(double) * (size_t)Lo.nf * h->pitch));
    CK(cudaMalloc(&h->d_ist, sizeof(int) * (size_t)Lo.i_nf * h->pitch));
    CK(cudaMalloc(&h->d_theta, sizeof(double) * (size_t)N * h->dp));
    CK(cudaMalloc(&h->d_mean, sizeof(double) * (size_t)N * h->dp));
    // pooled adaptation: every chain proposes from ONE shared factor (stride 0) -- except RAM, whose chains
    // keep private factors between the averaging ticks
    const bool shared_factor = c.pool_adapt && c.method != MCMCB_RAM;
    h->r_stride = shared_factor ? 0 : (long long)d * d;
    h->q_stride = shared_factor ? 0 : h->dp;
    const size_t NR = shared_factor ? 1 : (size_t)N;
    CK(cudaMalloc(&h->d_Rm, sizeof(double) * NR * d * d));
    CK(cudaMalloc(&h->d_cmat, sizeof(double) * (size_t)N * d * d));
    CK(cudaMalloc(&h->d_scratch, sizeof(double) * NR * d * d * (h->factor_mode == FACTOR_CHOL ? 1 : 2)));
    CK(cudaMalloc(&h->d_qstd, sizeof(double) * NR * h->dp));
    CK(cudaMemsetAsync(h->d_qstd, 0, sizeof(double) * NR * h->dp, h->stream));
    CK(cudaMalloc(&h->d_rowbuf, sizeof(double) * (size_t)N * (h->rowcap + 1) * (d + 1)));
    CK(cudaMalloc(&h->d_coef, sizeof(double) * (size_t)N * 2 * (h->rowcap + 1)));
    if (greedy) {  // unit-weight accumulators of the greedy burn-in (MCMC_adapt.F90:83-101)
      CK(cudaMalloc(&h->d_gcm, sizeof(double) * (size_t)N * d * d));
      CK(cudaMalloc(&h->d_gmean, sizeof(double) * (size_t)N * h->dp));
      CK(cudaMalloc(&h->d_gw, sizeof(double) * (size_t)N));
    }
    if (h->store_chains > 0) {
      size_t rows = (size_t)h->store_chains * h->cfg.nsimu;
      CK(cudaMalloc(&h->d_store_rows, sizeof(double) * rows * (d + NY)));
      CK(cudaMalloc(&h->d_store_cnt, sizeof(double) * rows));
      CK(cudaMalloc(&h->d_store_s2, sizeof(double) * rows * NY));
      CK(cudaMemsetAsync(h->d_store_rows, 0, sizeof(double) * rows * (d + NY), h->stream));
      CK(cudaMemsetAsync(h->d_store_cnt, 0, sizeof(double) * rows, h->stream));
      CK(cudaMemsetAsync(h->d_store_s2, 0, sizeof(double) * rows * NY, h->stream));
    }
    return 0;
  }

  static int init(mcmcb_handle h) {
    K2Params p = params(h, 0);
    k2_init_kernel<M><<<(unsigned)h->cfg.nchains, 128, 0, h->stream>>>(p);
    const unsigned nfac = h->r_stride == 0 ? 1u : (unsigned)h->cfg.nchains;  // shared factor: chain 0's cmat == cmat0
    if (h->factor_mode == FACTOR_CHOL)
      k2_initR_kernel<<<nfac, K2_ADAPT_THREADS, 0, h->stream>>>(p, h->d_scratch);
    else
      k3_initR_kernel<<<nfac, K2_ADAPT_THREADS, sizeof(double) * 2 * h->npar, h->stream>>>(p, h->d_scratch,
                                                                                            h->factor_mode);
    h->launches += 2;
    h->k2_i = 1;
    CK(cudaGetLastError());
    return 0;
  }

  template <bool SMEM>
  static int launch_step(mcmcb_handle h, const K2Params& p) {
    auto kern = (h->factor_mode == FACTOR_SCAM) ? k3_scam_step_kernel<M, SMEM> : k2_step_kernel<M, SMEM>;
    const int W = h->k2_warps;
    size_t smem = sizeof(double) * (size_t)W * K2_NVEC * h->dp + (SMEM ? h->blob_bytes : 0) +
                  (h->r_resident ? sizeof(double) * (size_t)W * h->npar * h->npar : 0);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!h->attr_set) {
      int occ = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, W * 32, smem));
      h->occ = std::max(occ, 1);
      h->attr_set = true;
    }
    long long need = (h->cfg.nchains + W - 1) / W;
    long long blocks = std::max<long long>(1, std::min<long long>((long long)h->num_sms * h->occ, need));
    h->blocks = (int)blocks;
    h->smem = smem;
    CK(cudaMemsetAsync(h->d_tile, 0, sizeof(unsigned), h->stream));
    kern<<<(unsigned)blocks, W * 32, smem, h->stream>>>(p);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }

  // group-of-warps-per-chain kernel (k2g_group.cuh): DRAM/AM/ER with a Cholesky factor, npar large enough that a
  // chain keeps several warps busy.  Returns false when it does not apply or does not fit.
  static bool plan_group(mcmcb_handle h, int& GT, int& ngroups, bool& smem_blob) {
    const mcmcb_config& c = h->cfg;
    const char* force = getenv("MCMCB_K2_GROUP");  // opt-in (see the status note in k2g_group.cuh): 1 = whenever it fits
    if (!(force && force[0] == '1')) return false;
    if (h->factor_mode != FACTOR_CHOL || c.method == MCMCB_RAM) return false;
    const int d = h->npar;
    GT = std::min(256, ((d + 31) / 32) * 32);
    if (d > GT * K2G_MAXM) return false;
    const size_t T = (size_t)d * (d + 1) / 2, Tp = (T + 1) & ~(size_t)1;
    const size_t per = sizeof(double) * ((size_t)K2_NVEC * h->dp + Tp + 2 * K2G_RED + 4);
    const size_t room = h->max_smem - 1024;
    const int cap = std::min(15, K2G_MAX_THREADS / GT);
    int with_blob = h->blob_bytes + per <= room ? (int)std::min<size_t>(cap, (room - h->blob_bytes) / per) : 0;
    int without = (int)std::min<size_t>(cap, room / per);
    if (with_blob >= 1 && 2 * with_blob >= without) { ngroups = with_blob; smem_blob = true; }
    else if (without >= 1) { ngroups = without; smem_blob = false; }
    else return false;
    return true;
  }

  template <bool SMEM>
  static int launch_group(mcmcb_handle h, const K2Params& p, int GT, int ngroups) {
    auto kern = k2g_step_kernel<M, SMEM>;
    const int d = h->npar;
    const size_t T = (size_t)d * (d + 1) / 2, Tp = (T + 1) & ~(size_t)1;
    const size_t smem = sizeof(double) * (size_t)ngroups * ((size_t)K2_NVEC * h->dp + Tp + 2 * K2G_RED + 4) +
                        (SMEM ? h->blob_bytes : 0);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!h->attr_set) {
      h->occ = 1;
      h->attr_set = true;
    }
    const long long need = (h->cfg.nchains + ngroups - 1) / ngroups;
    const long long blocks = std::max<long long>(1, std::min<long long>(h->num_sms, need));
    h->blocks = (int)blocks;
    h->smem = smem;
    CK(cudaMemsetAsync(h->d_tile, 0, sizeof(unsigned), h->stream));
    kern<<<(unsigned)blocks, GT * ngroups, smem, h->stream>>>(p, GT);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
  }

  // thread-per-chain RAM kernel (k4_ram.cuh): when the population gives every lane of every resident warp a chain
  static bool use_k4(mcmcb_handle h) {
    const mcmcb_config& c = h->cfg;
    if (c.method != MCMCB_RAM || !c.doadapt || h->factor_mode != FACTOR_CHOL || h->npar > K4_DM) return false;
    if (const char* e = getenv("MCMCB_K4")) return e[0] == '1';  // tuning / tests: force on or off
    return c.nchains >= (long long)h->num_sms * 64;
  }

  static int launch_k4(mcmcb_handle h, const K2Params& p) {
    const int d = h->npar;
    const size_t T = (size_t)d * (d + 1) / 2;
    const long long N = h->cfg.nchains;
    if (!h->d_Rp) CK(cudaMalloc(&h->d_Rp, sizeof(double) * T * (size_t)h->pitch));
    // CTA size: 32, 64 and 128 threads measured identical on BASELINE C4 (5.71e7 chain-steps/s each,
    // profiles/r02_summary.md); MCMCB_K4_BLOCK overrides for experiments
    int threads = K4_THREADS;
    if (const char* e = getenv("MCMCB_K4_BLOCK")) threads = std::max(32, std::min(K4_THREADS, (atoi(e) / 32) * 32));
    const unsigned blocks = (unsigned)((N + threads - 1) / threads);
    k4_pack_kernel<<<blocks, threads, 0, h->stream>>>(h->d_Rm, h->r_stride, h->d_Rp, h->pitch, N, d);
    k4_ram_step_kernel<M><<<blocks, threads, 0, h->stream>>>(p, h->d_Rp);
    k4_unpack_kernel<<<blocks, threads, 0, h->stream>>>(h->d_Rm, h->r_stride, h->d_Rp, h->pitch, N, d);
    h->launches += 3;
    h->blocks = (int)blocks;
    h->smem = 0;
    h->k2_warps = threads / 32;
    h->k4 = true;
    CK(cudaGetLastError());
    return 0;
  }

  // thread-per-chain SCAM kernel (k5_scam.cuh): a large population that shares ONE rotation (pooled adaptation)
  static bool use_k5(mcmcb_handle h) {
    if (h->factor_mode != FACTOR_SCAM || h->r_stride != 0 || h->q_stride != 0 || h->npar > K4_DM) return false;
    if (const char* e = getenv("MCMCB_K5")) return e[0] == '1';  // tuning / tests: force on or off
    return h->cfg.nchains >= (long long)h->num_sms * 64;
  }

  template <bool SMEM>
  static int launch_k5(mcmcb_handle h, const K2Params& p) {
    auto kern = k5_scam_step_kernel<M, SMEM>;
    const size_t smem = SMEM ? h->blob_bytes : 0;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)((h->cfg.nchains + K5_THREADS - 1) / K5_THREADS);
    kern<<<blocks, K5_THREADS, smem, h->stream>>>(p);
    h->launches++;
    h->blocks = (int)blocks;
    h->smem = smem;
    h->k2_warps = K5_THREADS / 32;
    h->k4 = true;  // one thread per chain (mcmcb_info)
    h->k5s_lanes = 0;
    CK(cudaGetLastError());
    return 0;
  }

  // theta-in-shared-memory variant (k5s_scam.cuh): the model must evaluate views, and >= 64 chains' theta must fit beside
  // the blob.  Returns the chains per CTA, 0 = not applicable.
  static int k5s_chains(mcmcb_handle h) {
    if constexpr (!has_ssfunction_view<M>::value) {
      return 0;
    } else {
      if (const char* e = getenv("MCMCB_K5S")) { if (e[0] == '0') return 0; }  // tuning / tests
      for (int nc = K5S_CHAINS; nc >= 64; nc -= 32)
        if (k5s_smem_bytes(h->npar, nc, h->blob_bytes) + 1024 <= h->max_smem) return nc;
      return 0;
    }
  }

  template <int W, int L, int C>
  static int launch_k5s_as(mcmcb_handle h, const K2Params& p, int chains) {
    if constexpr (has_ssfunction_view<M>::value) {
      auto kern = k5s_scam_step_kernel<M, W, L, C>;
      const size_t smem = k5s_smem_bytes(h->npar, chains, h->blob_bytes);
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const unsigned blocks = (unsigned)((h->cfg.nchains + chains - 1) / chains);
      kern<<<blocks, chains * L / C, smem, h->stream>>>(p);
      h->launches++;
      h->blocks = (int)blocks;
      h->smem = smem;
      h->k2_warps = chains * L / C / 32;
      h->k5s_lanes = L;
      h->k4 = L == 1;  // one thread per chain (mcmcb_info)
      CK(cudaGetLastError());
    }
    return 0;
  }

  // lanes per chain x chains per thread: 4 x 2 by default (profiles/r02_summary.md); the others for tuning / tests
  static int launch_k5s(mcmcb_handle h, const K2Params& p, int chains) {
    int lanes = 4, cpt = 2;
    if (const char* e = getenv("MCMCB_K5S_LANES")) lanes = atoi(e);
    if (const char* e = getenv("MCMCB_K5S_CPT")) cpt = atoi(e);
    if (lanes == 1) return launch_k5s_as<8, 1, 1>(h, p, chains);
    if (lanes == 2) return launch_k5s_as<4, 2, 1>(h, p, chains);
    if (lanes == 8) return cpt == 4 ? launch_k5s_as<2, 8, 4>(h, p, chains) : launch_k5s_as<4, 8, 2>(h, p, chains);
    if (cpt == 1) return launch_k5s_as<4, 4, 1>(h, p, chains);
    return launch_k5s_as<4, 4, 2>(h, p, chains);
  }

  static bool is_tick(const mcmcb_config& c, long long i) {
    if (c.method == MCMCB_RAM) return false;
    if (!c.doadapt && !c.doburnin) return false;
    if (c.adaptend > 0 && i > c.adaptend) return false;
    const bool ta = c.adaptint > 0 && i % c.adaptint == 0, tb = c.badaptint > 0 && i % c.badaptint == 0;
    return ta || tb;
  }

  static int step(mcmcb_handle h, int nsteps) {
    const mcmcb_config& c = h->cfg;
    if (use_k4(h)) {  // RAM has no adaptation ticks: one launch for the whole call
      if (nsteps == 0 && h->k2_i > 1) return 0;
      const int rc = launch_k4(h, params(h, nsteps));
      h->k2_i += nsteps;
      return rc;
    }
    // the blob shares shared memory with the per-warp vectors
    // shared-memory plan: K2_MAX_WARPS warps per CTA (latency hiding: these kernels are issue/latency bound),
    // the model blob beside the per-warp vectors when it fits.  RAM rewrites its factor every step, so for RAM
    // a per-warp resident copy of the factor comes first, with as many warps as still fit.
    const size_t d2 = sizeof(double) * (size_t)h->npar * h->npar, vec1 = sizeof(double) * (size_t)K2_NVEC * h->dp;
    const bool ram = c.method == MCMCB_RAM && c.doadapt;
    int W = K2_MAX_WARPS;
    bool resident = false, smem_blob = false;
    const char* force = getenv("MCMCB_K2_RESIDENT");  // tuning experiments only
    if (h->factor_mode != FACTOR_SCAM && (ram || (force && force[0] == '1')) && !(force && force[0] == '0')) {
      int w = (int)std::min<size_t>(K2_MAX_WARPS, (h->max_smem - 1024) / (vec1 + d2));
      if (w >= 4) { resident = true; W = w; }
    }
    // private Cholesky factors read from HBM/L2 for every proposal: keep the factors of the chains in flight inside
    // L2 (about half of the 126 MB is usable from one side).  At npar = 100, 16 warps per CTA re-read 109 MB per step
    // (L2 hit 24 %, 25 GB of DRAM reads per 200 steps); 8 warps fit (hit 97 %, 0.23 GB) and are as fast
    // (profiles/r01_summary.md K)
    if (!resident && h->factor_mode == FACTOR_CHOL && h->r_stride > 0 && c.method != MCMCB_RAM &&
        (size_t)h->num_sms * W * (d2 / 2) > ((size_t)64 << 20))
      W = std::max(8, W / 2);
    if (const char* e = getenv("MCMCB_K2_WARPS")) {  // tuning experiments only
      const int w = atoi(e);
      if (w >= 1 && w <= K2_MAX_WARPS && !resident) W = w;
    }
    const size_t used = (size_t)W * (vec1 + (resident ? d2 : 0));
    if (used + 1024 > h->max_smem) W = (int)std::max<size_t>(1, (h->max_smem - 1024) / vec1);
    smem_blob = (size_t)W * (vec1 + (resident ? d2 : 0)) + h->blob_bytes + 1024 <= h->max_smem;
    if (!smem_blob && !resident && W > 8 && (size_t)8 * vec1 + h->blob_bytes + 1024 <= h->max_smem) {
      W = 8;  // a blob in shared memory beats the extra warps
      smem_blob = true;
    }
    int GT = 0, ngroups = 0;
    bool gblob = false;
    const bool group = plan_group(h, GT, ngroups, gblob);
    if (group) { resident = false; W = GT * ngroups / 32; h->k2_group_threads = GT; }
    if (resident != h->r_resident || W != h->k2_warps) { h->r_resident = resident; h->k2_warps = W; h->attr_set = false; }
    const bool k5 = use_k5(h);
    const int k5s = k5 ? k5s_chains(h) : 0;
    int left = nsteps;
    bool first = true;
    while (left > 0 || first) {
      first = false;
      int seg = left;
      for (int k = 1; k <= left; k++)
        if (is_tick(c, h->k2_i + k)) { seg = k; break; }
      K2Params p = params(h, seg);
      int rc;
      if (k5 && k5s) rc = launch_k5s(h, p, k5s);
      else if (k5) rc = (h->blob_bytes + 2048 <= h->max_smem) ? launch_k5<true>(h, p) : launch_k5<false>(h, p);
      else rc = group ? (gblob ? launch_group<true>(h, p, GT, ngroups) : launch_group<false>(h, p, GT, ngroups))
                      : (smem_blob ? launch_step<true>(h, p) : launch_step<false>(h, p));
      if (rc) return rc;
      h->k2_i += seg;
      left -= seg;
      if (seg > 0 && is_tick(c, h->k2_i)) {
        p.tick_i = (int)h->k2_i;
        // plain AM ticks whose logged rows fit in shared memory: one CTA per SM folds them into the covariance at the
        // FP64 rate (k2_absorb_resident_kernel); the factorisation, if the chains keep private factors, follows in
        // the many-CTAs-per-SM kernel that hides its barrier latency
        if (tick_is_plain(c, h->k2_i) && absorb_resident_ok(h)) {
          const size_t sm = absorb_resident_smem_bytes(h->npar, p.rowcap);
          const int ntiles = ((h->npar + 3) / 4) * ((h->npar + 3) / 4 + 1) / 2;
          const int threads = std::min(K2_ABSR_THREADS, std::max(64, (ntiles + 31) / 32 * 32));
          CK(cudaFuncSetAttribute(k2_absorb_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
          k2_absorb_resident_kernel<<<(unsigned)c.nchains, threads, sm, h->stream>>>(p);
          h->launches++;
          CK(cudaGetLastError());
          p.absorbed = 1;
        }
        if (p.absorbed && c.pool_adapt) {
          // nothing else happens per chain: the pooled factor is built once for everybody (pool())
        } else if (h->factor_mode == FACTOR_CHOL) {
          k2_adapt_kernel<<<(unsigned)c.nchains, K2_ADAPT_THREADS,
                            sizeof(double) * absorb_smem_doubles(h->npar), h->stream>>>(p, h->d_scratch);
          h->launches++;
        } else {
          k3_adapt_kernel<<<(unsigned)c.nchains, K2_ADAPT_THREADS,
                            sizeof(double) * (2 * h->npar + absorb_smem_doubles(h->npar)), h->stream>>>(
              p, h->d_scratch, h->factor_mode);
          h->launches++;
        }
        CK(cudaGetLastError());
      }
    }
    return 0;
  }

  // the tick at step i takes the plain branch of MCMC_adapt.F90:105-159 (not burn-in scaling, not the AP window)
  static bool tick_is_plain(const mcmcb_config& c, long long i) {
    return c.doadapt && c.adapthist <= 1 && i >= (long long)c.burnintime + c.adaptint + c.adapthist;
  }

  static bool absorb_resident_ok(mcmcb_handle h) {
    if (const char* e = getenv("MCMCB_TICK_RESIDENT")) { if (e[0] == '0') return false; }  // tuning / tests
    const K2Params p = params(h, 0);
    return h->npar >= 16 && absorb_resident_smem_bytes(h->npar, p.rowcap) + 1024 <= h->max_smem;
  }

  // pooled adaptation, pool.cuh
  static int pool(mcmcb_handle h, int phase) {
    K2Params p = params(h, 0);
    const int d = h->npar;
    if (phase == 3) {
      k2_pool_factor_kernel<<<1, K2_ADAPT_THREADS, sizeof(double) * 3 * d, h->stream>>>(p, h->d_pool, h->d_scratch,
                                                                                       h->d_Rpool, h->d_fail);
      if (h->cfg.method == MCMCB_RAM)
        k2_pool_broadcast_kernel<<<h->num_sms * 4, 256, 0, h->stream>>>(p, h->d_Rpool, h->d_fail);
      const K2Layout Lo = k2_layout(NY);
      pool_flag_kernel<<<h->num_sms, 256, 0, h->stream>>>(h->d_ist + (size_t)Lo.i_status * h->pitch, h->pitch,
                                                          h->cfg.nchains, h->d_fail);
      h->launches += 3;
    } else {
      const int nv = phase == 1 ? 1 + d : d * d;
      double* out = phase == 1 ? h->d_pool : h->d_pool + 1 + d;
      if (phase == 2 && h->cfg.method != MCMCB_RAM && d >= 64) {  // (tile, slice) grid, four chains in flight per thread
        const int S = pool_cov_slices(d), ntile = (d + POOL_COV_TB - 1) / POOL_COV_TB;
        k2_pool_cov_kernel<<<ntile * S, POOL_THREADS, 0, h->stream>>>(p, h->d_pool, h->d_pool_partial);
        pool_final_kernel<<<(nv + 255) / 256, 256, 0, h->stream>>>(h->d_pool_partial, S, nv, out);
      } else {
        k2_pool_moments_kernel<<<POOL_BLOCKS, POOL_THREADS, 0, h->stream>>>(p, phase, h->d_pool, h->d_pool_partial);
        pool_final_kernel<<<(nv + 255) / 256, 256, 0, h->stream>>>(h->d_pool_partial, POOL_BLOCKS, nv, out);
      }
      h->launches += 2;
    }
    CK(cudaGetLastError());
    return 0;
  }

  static ModelEntry entry() {
    ModelEntry e{};
    e.pool = &pool;
    e.name = M::name();
    e.kernel = 2;
    e.npar = 0;
    e.ny = NY;
    e.alloc = &alloc;
    e.init = &init;
    e.step = &step;
    e.fetch = &k2_fetch;
    e.fetch_chain = &k2_fetch_chain;
    return e;
  }
};

}  // namespace launch
}  // namespace mcmcb
