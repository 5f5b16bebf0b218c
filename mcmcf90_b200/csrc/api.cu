// C ABI of the batched adaptive-MH path (include/mcmcb200.h): handle, model registry,
// kernel launchers, state fetch, streamed dumps.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
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

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                    \
      return MCMCB_ECUDA;                                                                             \
    }                                                                                                 \
  } while (0)

// ------------------------------------------------------------------ registry
namespace mcmcb {
std::vector<ModelEntry>& registry() {
  static std::vector<ModelEntry> r;
  return r;
}
int register_model(const ModelEntry& e) {
  for (auto& x : registry())
    if (std::strcmp(x.name, e.name) == 0 && x.kernel == e.kernel) {
      x = e;
      return 0;
    }
  registry().push_back(e);
  return 0;
}
}  // namespace mcmcb

// ------------------------------------------------------------------ K1 launcher
namespace {

static int fetch_fields(mcmcb_handle h, int f0, int nf, std::vector<double>& buf) {
  buf.resize((size_t)nf * h->pitch);
  CK(cudaMemcpyAsync(buf.data(), h->d_st + (size_t)f0 * h->pitch, sizeof(double) * buf.size(), cudaMemcpyDeviceToHost,
                     h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// SoA state -> the chain-major host layouts of mcmcb_fetch, transposed on the device so that the
// device->host copy is one contiguous transfer straight into the caller's buffer (pinned or not)
__global__ void k1_gather_fields_kernel(const double* st, long long pitch, long long n, int f0, int width, double* out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  const long long c = t / width;
  const int k = (int)(t - c * width);
  out[t] = st[(size_t)(f0 + k) * pitch + c];
}
// packed upper triangle (column-packed) -> full D x D column-major per chain; sym mirrors, else zero below
__global__ void k1_gather_tri_kernel(const double* st, long long pitch, long long n, int f0, int D, int sym, double* out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * D * D) return;
  const long long c = t / (D * D);
  const int e = (int)(t - c * D * D), j = e / D, i = e - j * D;
  double v = 0.0;
  if (i <= j) v = st[(size_t)(f0 + j * (j + 1) / 2 + i) * pitch + c];
  else if (sym) v = st[(size_t)(f0 + i * (i + 1) / 2 + j) * pitch + c];
  out[t] = v;
}
__global__ void k1_gather_counters_kernel(const int* ist, long long pitch, long long n, K1Layout Lo, long long* out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int src[7] = {Lo.i_stayed, Lo.i_bnd, Lo.i_dracc, Lo.i_drtry, Lo.i_chainind, Lo.i_simuind, Lo.i_status};
#pragma unroll
  for (int k = 0; k < 7; k++) out[c * 8 + k] = ist[(size_t)src[k] * pitch + c];
  const unsigned lo = (unsigned)ist[(size_t)Lo.i_ndlo * pitch + c], hi = (unsigned)ist[(size_t)Lo.i_ndhi * pitch + c];
  out[c * 8 + 7] = (long long)(((unsigned long long)hi << 32) | lo);
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

  template <int L, bool SMEM>
  static int launch_LS(mcmcb_handle h, const K1Params& p) {
    auto kern = k1_step_kernel<M, L, SMEM>;
    size_t smem = MCMCB_EXP_TAB_DOUBLES * sizeof(double) + (SMEM ? h->blob_bytes : 0);
    if (!h->attr_set) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int occ = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, K1_THREADS, smem));
      if (occ < 1) occ = 1;
      h->occ = occ;
      h->attr_set = true;
    }
    long long tiles = (h->cfg.nchains * L + 31) / 32;
    long long wpb = K1_THREADS / 32;
    long long need = (tiles + wpb - 1) / wpb;
    long long blocks = std::min<long long>((long long)h->num_sms * h->occ, need);
    if (blocks < 1) blocks = 1;
    h->blocks = (int)blocks;
    h->smem = smem;
    CK(cudaMemsetAsync(h->d_tile, 0, sizeof(unsigned), h->stream));
    kern<<<(unsigned)blocks, K1_THREADS, smem, h->stream>>>(p);
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
    p.qstd = h->d_qstd;
    p.factor_mode = h->factor_mode;
    p.r_stride = h->r_stride;
    p.q_stride = h->q_stride;
    p.r_resident = h->r_resident ? 1 : 0;
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
    if (c.method == MCMCB_RAM || !c.doadapt) h->rowcap = 1;
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
    if (!h->attr_set) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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

  static bool is_tick(const mcmcb_config& c, long long i) {
    if (c.method == MCMCB_RAM) return false;
    if (!c.doadapt && !c.doburnin) return false;
    if (c.adaptend > 0 && i > c.adaptend) return false;
    const bool ta = c.adaptint > 0 && i % c.adaptint == 0, tb = c.badaptint > 0 && i % c.badaptint == 0;
    return ta || tb;
  }

  static int step(mcmcb_handle h, int nsteps) {
    const mcmcb_config& c = h->cfg;
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
    const size_t used = (size_t)W * (vec1 + (resident ? d2 : 0));
    if (used + 1024 > h->max_smem) W = (int)std::max<size_t>(1, (h->max_smem - 1024) / vec1);
    smem_blob = (size_t)W * (vec1 + (resident ? d2 : 0)) + h->blob_bytes + 1024 <= h->max_smem;
    if (!smem_blob && !resident && W > 8 && (size_t)8 * vec1 + h->blob_bytes + 1024 <= h->max_smem) {
      W = 8;  // a blob in shared memory beats the extra warps
      smem_blob = true;
    }
    if (resident != h->r_resident || W != h->k2_warps) { h->r_resident = resident; h->k2_warps = W; h->attr_set = false; }
    int left = nsteps;
    bool first = true;
    while (left > 0 || first) {
      first = false;
      int seg = left;
      for (int k = 1; k <= left; k++)
        if (is_tick(c, h->k2_i + k)) { seg = k; break; }
      K2Params p = params(h, seg);
      int rc = smem_blob ? launch_step<true>(h, p) : launch_step<false>(h, p);
      if (rc) return rc;
      h->k2_i += seg;
      left -= seg;
      if (seg > 0 && is_tick(c, h->k2_i)) {
        p.tick_i = (int)h->k2_i;
        if (h->factor_mode == FACTOR_CHOL)
          k2_adapt_kernel<<<(unsigned)c.nchains, K2_ADAPT_THREADS,
                            sizeof(double) * absorb_smem_doubles(h->rowcap, h->npar), h->stream>>>(p, h->d_scratch);
        else
          k3_adapt_kernel<<<(unsigned)c.nchains, K2_ADAPT_THREADS,
                            sizeof(double) * (2 * h->npar + absorb_smem_doubles(h->rowcap, h->npar)), h->stream>>>(
              p, h->d_scratch, h->factor_mode);
        h->launches++;
        CK(cudaGetLastError());
      }
    }
    return 0;
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
      k2_pool_moments_kernel<<<POOL_BLOCKS, POOL_THREADS, 0, h->stream>>>(p, phase, h->d_pool, h->d_pool_partial);
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

struct BuiltinRegistrar {
  BuiltinRegistrar() {
    register_model(K1<ExpReg>::entry());
    register_model(K2<GaussN>::entry());
    register_model(K2<BananaN>::entry());
    register_model(K2<HierN>::entry());
  }
} builtin_registrar;

const ModelEntry* find_model(const char* name, int kernel) {
  for (auto& e : registry())
    if (std::strcmp(e.name, name) == 0 && (kernel == 0 || e.kernel == kernel)) return &e;
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
  if (dodr) *dodr = c->drscale > 0.0;
  if (doscam) *doscam = sc;
  if (usesvd) *usesvd = c->condmax > 0.0;
  return MCMCB_OK;
}

// ------------------------------------------------------------------ lifecycle
extern "C" int mcmcb_create(const mcmcb_config* cfg, mcmcb_handle* out) {
  if (!cfg || !out || cfg->abi_version != MCMCB_ABI_VERSION || cfg->nchains < 1 || cfg->nsimu < 1) return MCMCB_EINVAL;
  mcmcb_handle h = new mcmcb_handle_s();
  h->cfg = *cfg;
  int rc = mcmcb_check_config(&h->cfg, &h->dodr, &h->doscam, &h->usesvd);
  if (rc) { delete h; return rc; }
  const mcmcb_config& c = h->cfg;
  // configurations that need the stored row history of the reference (AP)
  if (c.method != MCMCB_RAM && (c.adapthist > 1 || (c.greedy && c.doburnin)) ) { delete h; return MCMCB_EUNSUPPORTED; }
  h->model = find_model(c.model, c.kernel);
  if (!h->model) { delete h; return MCMCB_ENOMODEL; }
  // SVD factor paths (SCAM, condmax > 0) live in the warp-per-chain kernels only
  if (h->model->kernel == 1 && (h->doscam || h->usesvd)) { delete h; return MCMCB_EUNSUPPORTED; }
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
  h->npar = h->model->npar;
  h->nycol = h->model->ny;
  *out = h;
  return MCMCB_OK;
}

static void free_dev(mcmcb_handle h) {
  void* ptrs[] = {h->d_st, h->d_ist, h->d_par0, h->d_cmat0, h->d_sigma2, h->d_nobs, h->d_blob, h->d_prior,
                  h->d_inj, h->d_store_rows, h->d_store_cnt, h->d_store_s2, h->d_tile, h->d_theta, h->d_mean, h->d_Rm,
                  h->d_cmat, h->d_rowbuf, h->d_scratch, h->d_cmat0_full, h->d_qstd,
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
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->copy_stream);
  free_dev(h);
  cudaStreamDestroy(h->stream);
  cudaStreamDestroy(h->copy_stream);
  delete h;
  return MCMCB_OK;
}

extern "C" const char* mcmcb_last_error(mcmcb_handle h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int mcmcb_set_data(mcmcb_handle h, const double* blob, size_t n) {
  if (!h || !blob || n == 0) return MCMCB_EINVAL;
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
  int rc = h->model->init(h);
  if (rc) return rc;
  h->initial_set = true;
  h->steps_done = 0;
  return MCMCB_OK;
}

extern "C" int mcmcb_inject_uniforms(mcmcb_handle h, const double* u, size_t per_chain) {
  if (!h) return MCMCB_EINVAL;
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
static int dump_enqueue(mcmcb_handle h) {
  // snapshot theta (field-major, npar x pitch) device->device on the compute stream, then
  // device->pinned host on the copy stream, so the next launch overlaps the PCIe copy
  const bool k2 = h->model->kernel == 2;
  const size_t bytes = k2 ? sizeof(double) * (size_t)h->cfg.nchains * h->dp : sizeof(double) * (size_t)h->npar * h->pitch;
  if (h->dump_slots.empty()) {
    h->dump_slots.resize(4);
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
  CK(cudaMemcpyAsync(s.dev, k2 ? h->d_theta : h->d_st /* theta is field 0 */, bytes, cudaMemcpyDeviceToDevice, h->stream));
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

extern "C" int mcmcb_dump_pop(mcmcb_handle h, double* out, size_t out_bytes, int* step) {
  if (!h || !out) return MCMCB_EINVAL;
  if (h->dump_fifo.empty()) return 0;
  int k = h->dump_fifo.front();
  auto& s = h->dump_slots[k];
  if (cudaEventQuery(s.ev) != cudaSuccess) return 0;
  const long long N = h->cfg.nchains;
  if (out_bytes < sizeof(double) * (size_t)N * h->npar) return MCMCB_EINVAL;
  const bool k2 = h->model->kernel == 2;
  for (int f = 0; f < h->npar; f++)
    for (long long c = 0; c < N; c++)
      out[(size_t)c * h->npar + f] = k2 ? s.host[(size_t)c * h->dp + f] : s.host[(size_t)f * h->pitch + c];
  if (step) *step = s.step;
  s.state = 0;
  h->dump_fifo.pop_front();
  return 1;
}

// ------------------------------------------------------------------ pooled adaptation / diagnostics
static int do_allreduce(mcmcb_handle h, double* dev, size_t n) {
  if (!h->ar_fn) return 0;  // single handle: the local sums are the global sums
  const int rc = h->ar_fn(h->ar_user, dev, n, (void*)h->stream);
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

static int pool_tick(mcmcb_handle h) {
  const size_t d = (size_t)h->npar;
  int rc = h->model->pool(h, 1);
  if (!rc) rc = do_allreduce(h, h->d_pool, 1 + d);
  if (!rc) rc = h->model->pool(h, 2);
  if (!rc) rc = do_allreduce(h, h->d_pool + 1 + d, d * d);
  if (!rc) rc = h->model->pool(h, 3);
  h->pool_ticks++;
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
extern "C" int mcmcb_run(mcmcb_handle h, int nsteps) {
  if (!h || nsteps < 0) return MCMCB_EINVAL;
  if (!h->initial_set || !h->d_blob) return MCMCB_EINVAL;
  if (h->cfg.rng_mode == MCMCB_RNG_INJECTED && !h->d_inj) return MCMCB_EINVAL;
  CK(cudaSetDevice(h->cfg.device));
  const mcmcb_config& c = h->cfg;
  int left = nsteps;
  do {
    // a launch ends where the host has something to do: streamed dump, diagnostics snapshot, pooled tick.
    // simuind after k more steps = 1 + steps_done + k (the first launch also evaluates the initial point)
    int n = left;
    if (c.dump_stride > 0) n = std::min<long long>(n, c.dump_stride - h->steps_done % c.dump_stride);
    if (c.diag_stride > 0) n = std::min<long long>(n, c.diag_stride - h->steps_done % c.diag_stride);
    if (c.pool_adapt) n = std::min<long long>(n, c.adaptint - (1 + h->steps_done) % c.adaptint);
    int rc = h->model->step(h, n);
    if (rc) return rc;
    h->steps_done += n;
    left -= n;
    if (n > 0 && is_pool_tick(c, 1 + h->steps_done)) {
      rc = pool_tick(h);
      if (rc) return rc;
    }
    if (n > 0 && c.diag_stride > 0 && h->steps_done % c.diag_stride == 0) {
      rc = diag_snapshot(h);
      if (rc) return rc;
    }
    if (n > 0 && c.dump_stride > 0 && h->steps_done % c.dump_stride == 0) {
      rc = dump_enqueue(h);
      if (rc) return rc;
    }
  } while (left > 0);
  return MCMCB_OK;
}

extern "C" int mcmcb_set_allreduce(mcmcb_handle h, mcmcb_allreduce_fn fn, void* user) {
  if (!h) return MCMCB_EINVAL;
  h->ar_fn = fn;
  h->ar_user = user;
  return MCMCB_OK;
}

extern "C" int mcmcb_pool_fetch(mcmcb_handle h, double* wsum, double* mean, double* cov) {
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
  h->diag_n = 0;
  return MCMCB_OK;
}

// R-hat and ESS from the chain-summed moments (Gelman et al., BDA3 11.4-11.5; Stan's multi-chain ESS with
// Geyer's initial positive sequence, lags limited to diag_lags)
extern "C" int mcmcb_diagnostics(mcmcb_handle h, double* rhat, double* ess, double* mean, double* var, long long* nsnap,
                                 long long* nchains_total) {
  if (!h || !h->d_diag || h->diag_n < 2) return MCMCB_EINVAL;
  CK(cudaSetDevice(h->cfg.device));
  const int d = h->npar, K = h->diag_K, nv = 2 + K;
  DiagParams p = diag_params(h);
  double* buf = h->d_diag_buf;
  dim3 grid(POOL_BLOCKS, d);
  diag_reduce_kernel<<<grid, POOL_THREADS, 0, h->stream>>>(p, 1, buf, h->d_diag_partial);
  diag_final_kernel<<<(d * 2 + 255) / 256, 256, 0, h->stream>>>(h->d_diag_partial, POOL_BLOCKS, 2, d, 1, buf);
  CK(cudaGetLastError());
  int rc = do_allreduce(h, buf, 1 + (size_t)d);
  if (rc) return rc;
  diag_reduce_kernel<<<grid, POOL_THREADS, 0, h->stream>>>(p, 2, buf, h->d_diag_partial);
  diag_final_kernel<<<(d * nv + 255) / 256, 256, 0, h->stream>>>(h->d_diag_partial, POOL_BLOCKS, nv, d, 2, buf + 1 + d);
  CK(cudaGetLastError());
  rc = do_allreduce(h, buf + 1 + d, (size_t)nv * d);
  if (rc) return rc;
  h->launches += 4;
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
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->copy_stream));
  return MCMCB_OK;
}

// ------------------------------------------------------------------ fetch
extern "C" int mcmcb_fetch(mcmcb_handle h, const char* what, void* out, size_t out_bytes) {
  if (!h || !what || !out || !h->initial_set) return MCMCB_EINVAL;
  CK(cudaSetDevice(h->cfg.device));
  return h->model->fetch(h, what, out, out_bytes);
}

extern "C" int mcmcb_fetch_chain(mcmcb_handle h, long long chain, int ld, double* chain_out, double* sschain_out,
                                 double* s2chain_out, int* nrows) {
  if (!h || !h->initial_set || chain < 0 || chain >= h->store_chains) return MCMCB_EINVAL;
  CK(cudaSetDevice(h->cfg.device));
  return h->model->fetch_chain(h, chain, ld, chain_out, sschain_out, s2chain_out, nrows);
}

// ------------------------------------------------------------------ introspection
extern "C" void* mcmcb_stream(mcmcb_handle h) { return h ? (void*)h->stream : nullptr; }
extern "C" long long mcmcb_launch_count(mcmcb_handle h) { return h ? h->launches : 0; }
extern "C" int mcmcb_info(mcmcb_handle h, int* npar, int* nycol, int* lanes, int* kernel, int* tpb, int* blocks,
                          size_t* smem) {
  if (!h) return MCMCB_EINVAL;
  if (npar) *npar = h->npar;
  if (nycol) *nycol = h->nycol;
  if (lanes) *lanes = h->L;
  if (kernel) *kernel = h->model ? h->model->kernel : 0;
  if (tpb) *tpb = (h->model && h->model->kernel == 2) ? h->k2_warps * 32 : K1_THREADS;
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
