// Pooled cross-chain adaptation (an extension -- the reference runs one chain and has
// nothing to pool).  At an adaptation tick every chain has folded its rows into its own
// (wsum, mean, cmat) accumulators exactly as MCMC_adapt does (MCMC_adapt.F90:141-157 through the
// covmat recursion, matutils.F90:283-310).  Pooling merges those accumulators over ALL chains:
//
//   W  = sum_c w_c                 S1 = sum_c w_c m_c                mu = S1 / W
//   S2 = sum_c [ (w_c - 1) C_c + w_c (m_c - mu)(m_c - mu)' ]         cov = S2 / (W - 1)
//
// which is the weighted covariance of every row of every chain about the global mean (the pairwise
// merge of Chan, Golub & LeVeque 1979 summed over chains).  The sums are plain sums, so across GPUs they
// are ONE sum-allreduce each; mu is needed before S2 (centred second moments, no cancellation), hence two
// phases.  Every chain then proposes from the factor of `cov`, built by the same MCMC_calculate_R
// arithmetic as the per-chain path.  RAM has no covariance accumulators: its pooled quantity is the
// average of the chains' shape matrices, S2 = sum_c R_c'R_c, W = number of chains, R = chol(S2 / W).
//
// Determinism: no floating-point atomics.  A fixed grid accumulates per-thread partials over a static
// chain -> thread map, reduces them in a fixed order inside the block, and one final block adds the
// per-block partials in index order; results are bit-reproducible for a given chain count per GPU.
//
// Buffer layout (doubles): [0] = W, [1 .. d] = S1, [1+d .. 1+d+d*d) = S2, full symmetric d x d.
#pragma once
#include "k1_small.cuh"
#include "k3_scam.cuh"

namespace mcmcb {

constexpr int POOL_BLOCKS = 296;   // 2 per SM
constexpr int POOL_THREADS = 256;

// fixed-order block sum; every thread gets the total.  red: POOL_THREADS/32 doubles of shared memory.
__device__ __forceinline__ double pool_block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
  return s;
}

// partial[b][v] -> out[v] = sum_b partial[b][v], b in index order
static __global__ void pool_final_kernel(const double* partial, int nblocks, int nv, double* out) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblocks; b++) s += partial[(size_t)b * nv + v];
    out[v] = s;
  }
}

// ------------------------------------------------------------------ K1 (packed SoA state, compile-time D)
// phase 1: partial[b] = [sum w, sum w m_k];  phase 2: partial[b][a + D b] = S2 contributions (full matrix);
// ram: phase 1 counts chains, phase 2 sums R'R.
template <int D, int NY>
__global__ void __launch_bounds__(POOL_THREADS) k1_pool_moments_kernel(K1Params p, int phase, const double* buf,
                                                                      double* partial) {
  constexpr int T = D * (D + 1) / 2;
  constexpr K1Layout Lo = k1_layout(D, NY);
  __shared__ double red[POOL_THREADS / 32];
  const bool ram = p.c.method == MCMCB_RAM;
  double acc[D * D > 1 + D ? D * D : 1 + D];
#pragma unroll
  for (int k = 0; k < (D * D > 1 + D ? D * D : 1 + D); k++) acc[k] = 0.0;
  double mu[D];
#pragma unroll
  for (int k = 0; k < D; k++) mu[k] = (phase == 2 && !ram) ? buf[1 + k] / buf[0] : 0.0;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < p.nchains; c += (long long)gridDim.x * blockDim.x) {
    const double* st = p.st + c;
    if (ram) {
      if (phase == 1) {
        acc[0] += 1.0;
      } else {
        double R[T];
#pragma unroll
        for (int k = 0; k < T; k++) R[k] = st[(Lo.r + k) * p.pitch];
#pragma unroll
        for (int b = 0; b < D; b++)
#pragma unroll
          for (int a = 0; a <= b; a++) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i <= a; i++) s = fma(R[pk(i, a)], R[pk(i, b)], s);
            acc[a + D * b] += s;
          }
      }
      continue;
    }
    const double w = st[Lo.wsum * p.pitch];
    if (!(w > 0.0)) continue;
    if (phase == 1) {
      acc[0] += w;
#pragma unroll
      for (int k = 0; k < D; k++) acc[1 + k] = fma(w, st[(Lo.mean + k) * p.pitch], acc[1 + k]);
    } else {
      double dm[D];
#pragma unroll
      for (int k = 0; k < D; k++) dm[k] = st[(Lo.mean + k) * p.pitch] - mu[k];
#pragma unroll
      for (int b = 0; b < D; b++)
#pragma unroll
        for (int a = 0; a <= b; a++)
          acc[a + D * b] += (w - 1.0) * st[(Lo.cm + pk(a, b)) * p.pitch] + w * (dm[a] * dm[b]);
    }
  }
  const int nv = (phase == 1) ? 1 + D : D * D;
  if (phase == 2) {
#pragma unroll
    for (int b = 0; b < D; b++)
#pragma unroll
      for (int a = b + 1; a < D; a++) acc[a + D * b] = acc[b + D * a];  // mirror the upper triangle
  }
  for (int v = 0; v < nv; v++) {
    const double s = pool_block_sum(acc[v], red);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * nv + v] = s;
  }
}

// every chain takes the factor of the pooled covariance (redundantly per thread: D is tiny)
template <int D, int NY>
__global__ void k1_pool_apply_kernel(K1Params p, const double* buf) {
  constexpr int T = D * (D + 1) / 2;
  constexpr K1Layout Lo = k1_layout(D, NY);
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.nchains) return;
  const bool ram = p.c.method == MCMCB_RAM;
  const double W = buf[0];
  const double* S2 = buf + 1 + D;
  double cm[T], R[T], R2[T], iC[T];
  const double den = ram ? W : W - 1.0;
#pragma unroll
  for (int b = 0; b < D; b++)
#pragma unroll
    for (int a = 0; a <= b; a++) cm[pk(a, b)] = S2[a + D * b] / den;
  double* st = p.st + c;
  bool ok;
  if (ram) {  // R = chol(mean of R'R): plain dpotf2, no 2.4/sqrt(d) scaling (R is the RAM shape factor itself)
    ok = chol_packed<D>(cm);
    if (ok) {
#pragma unroll
      for (int k = 0; k < T; k++) st[(Lo.r + k) * p.pitch] = cm[k];
    }
  } else {
    ok = den > 0.0 && calculate_R<D>(cm, R, R2, iC, p.c);
    if (ok) {
#pragma unroll
      for (int k = 0; k < T; k++) {
        st[(Lo.r + k) * p.pitch] = R[k];
        if (p.c.dodr) { st[(Lo.r2 + k) * p.pitch] = R2[k]; st[(Lo.ic + k) * p.pitch] = iC[k]; }
      }
    }
  }
  if (!ok) p.ist[Lo.i_status * p.pitch + c] |= MCMCB_ST_CHOLFAIL;  // old factor kept, MCMC_adapt.F90:169-171
}

// ------------------------------------------------------------------ K2 / K3 (run-time d, per-chain full matrices)
// CTA b owns chains b, b + gridDim.x, ...; threads run over the output entries.
static __global__ void __launch_bounds__(POOL_THREADS) k2_pool_moments_kernel(K2Params p, int phase, const double* buf,
                                                                      double* partial) {
  constexpr K2Layout Lo = k2_layout(1);
  const int d = p.d;
  const bool ram = p.c.method == MCMCB_RAM;
  if (phase == 1) {
    const int nv = 1 + d;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
      double acc = 0.0;
      for (long long c = blockIdx.x; c < p.nchains; c += gridDim.x) {
        const double w = ram ? 1.0 : p.st[Lo.wsum * p.pitch + c];
        if (!(w > 0.0)) continue;
        if (v == 0) acc += w;
        else if (!ram) acc = fma(w, p.mean[c * p.dp + (v - 1)], acc);
      }
      partial[(size_t)blockIdx.x * nv + v] = acc;
    }
    return;
  }
  // phase 2: RAM (sum of R'R), and the covariance modes at small npar (from npar = 64 up they run k2_pool_cov_kernel)
  const int nv = d * d;
  const double W = buf[0];
  for (int v = threadIdx.x; v < nv; v += blockDim.x) {
    const int b = v / d, a = v - b * d;  // entry (a, b), column-major; symmetric result
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    double acc = 0.0;
    for (long long c = blockIdx.x; c < p.nchains; c += gridDim.x) {
      if (ram) {  // (R'R)(a,b) = sum_{i <= min(a,b)} R(i,a) R(i,b); R row-major upper
        const double* R = p.Rm + (size_t)c * p.r_stride;
        double s = 0.0;
        for (int i = 0; i <= lo; i++) s = fma(R[(size_t)i * d + lo], R[(size_t)i * d + hi], s);
        acc += s;
      } else {
        const double w = p.st[Lo.wsum * p.pitch + c];
        if (!(w > 0.0)) continue;
        const double da = p.mean[c * p.dp + a] - buf[1 + a] / W, db = p.mean[c * p.dp + b] - buf[1 + b] / W;
        // cmat is kept symmetric by the recursion (cta_absorb updates every entry)
        acc += (w - 1.0) * p.cmat[(size_t)c * d * d + (size_t)lo * d + hi] + w * (da * db);
      }
    }
    partial[(size_t)blockIdx.x * nv + v] = acc;
  }
}

// Phase 2 of the covariance modes from npar = 64 up: CTA (tile, slice) owns rows b0 .. b0+3 of the result and the
// chains slice, slice + S, ...; thread a owns the four entries (a, b0 + q).  A chain costs a thread ONE mean element of its
// own (the four mean(b) are uniform), four consecutive-address covariance reads and four independent accumulators; four
// chains are in flight, so 16 covariance loads per thread are outstanding (84 GB at BASELINE C5: the layout of
// k2_pool_moments_kernel, entries outer / chains inner with one dependent load chain per thread, ran at 0.6 TB/s).
// cmat is exactly symmetric after a tick (both images of an entry are written from one register by cta_absorb_rows /
// k2_absorb_resident_kernel / cta_ap_window), so image (b, a) -- consecutive a, consecutive addresses -- is read.
// partial[slice][v]; pool_final_kernel adds the slices in index order: bit-reproducible for a given chain count per GPU.
constexpr int POOL_COV_TB = 4;  // rows of the result per CTA
__host__ __device__ __forceinline__ int pool_cov_slices(int d) { return max(1, POOL_BLOCKS / ((d + POOL_COV_TB - 1) / POOL_COV_TB)); }

static __global__ void __launch_bounds__(POOL_THREADS) k2_pool_cov_kernel(K2Params p, const double* buf, double* partial) {
  constexpr K2Layout Lo = k2_layout(1);
  constexpr int U = 4, TB = POOL_COV_TB;
  const int d = p.d, nv = d * d;
  const int ntile = (d + TB - 1) / TB, S = pool_cov_slices(d);
  const int tile = blockIdx.x % ntile, slice = blockIdx.x / ntile, b0 = TB * tile;
  const double W = buf[0];
  int bq[TB];
  double mub[TB];
#pragma unroll
  for (int q = 0; q < TB; q++) { bq[q] = min(b0 + q, d - 1); mub[q] = buf[1 + bq[q]] / W; }
  for (int a = threadIdx.x; a < d; a += blockDim.x) {
    const double mua = buf[1 + a] / W;
    double acc[TB];
#pragma unroll
    for (int q = 0; q < TB; q++) acc[q] = 0.0;
    for (long long c = slice; c < p.nchains; c += (long long)U * S) {
      double w[U], t[U][TB];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const long long cc = c + (long long)u * S;
        w[u] = cc < p.nchains ? p.st[Lo.wsum * p.pitch + cc] : 0.0;
        if (w[u] > 0.0) {
          const double* m = p.mean + cc * p.dp;
          const double* cm = p.cmat + (size_t)cc * nv + a;
          const double da = m[a] - mua;
#pragma unroll
          for (int q = 0; q < TB; q++) t[u][q] = fma(w[u] - 1.0, cm[(size_t)bq[q] * d], w[u] * (da * (m[bq[q]] - mub[q])));
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++)
        if (w[u] > 0.0) {
#pragma unroll
          for (int q = 0; q < TB; q++) acc[q] += t[u][q];
        }
    }
#pragma unroll
    for (int q = 0; q < TB; q++)
      if (b0 + q < d) partial[(size_t)slice * nv + (size_t)(b0 + q) * d + a] = acc[q];
  }
}

// One CTA: pooled covariance -> factor.  Shared factor (r_stride == 0) is written in place; for RAM the
// new factor goes to `Rpool` and k2_pool_broadcast_kernel copies it into every chain's private factor.
static __global__ void __launch_bounds__(K2_ADAPT_THREADS) k2_pool_factor_kernel(K2Params p, double* buf, double* scratch,
                                                                           double* Rpool, int* fail) {
  extern __shared__ double sh[];  // 2 d doubles + d ints
  __shared__ double red[K2_ADAPT_THREADS / 32];
  const int d = p.d;
  const bool ram = p.c.method == MCMCB_RAM;
  const double W = buf[0];
  double* cov = buf + 1 + d;
  const double den = ram ? W : W - 1.0;
  for (int k = threadIdx.x; k < d * d; k += blockDim.x) cov[k] = cov[k] / den;
  __syncthreads();
  int status = 0;
  if (p.factor_mode == FACTOR_CHOL && ram) {  // the RAM shape factor is chol(.) itself: no 2.4/sqrt(d)
    for (int k = threadIdx.x; k < d * d; k += blockDim.x) scratch[k] = cov[k];
    __syncthreads();
    if (cta_cholesky_rowmajor(scratch, d, red)) {
      status = MCMCB_ST_CHOLFAIL;
    } else {
      for (int k = threadIdx.x; k < d * d; k += blockDim.x) Rpool[k] = (k % d >= k / d) ? scratch[k] : 0.0;
    }
  } else if (p.factor_mode == FACTOR_CHOL) {
    if (!(den > 0.0) || !cta_calculate_R(cov, p.Rm, scratch, d, red)) status = MCMCB_ST_CHOLFAIL;
  } else {
    double* sv = sh + d;
    int* perm = reinterpret_cast<int*>(sh + 2 * d);
    status = cta_calculate_R_svd(cov, p.Rm, p.qstd, scratch, scratch + (size_t)d * d, d, p.factor_mode, p.c.condmax, sv,
                                 perm, red);
  }
  __syncthreads();
  if (threadIdx.x == 0) *fail = status;
}

static __global__ void k2_pool_broadcast_kernel(K2Params p, const double* Rpool, const int* fail) {
  if (*fail) return;
  const size_t n = (size_t)p.d * p.d;
  const size_t total = n * (size_t)p.nchains;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
    p.Rm[(k / n) * (size_t)p.r_stride + (k % n)] = Rpool[k % n];
}

static __global__ void pool_flag_kernel(int* ist_status, long long pitch, long long nchains, const int* fail) {
  const int f = *fail;
  if (!f) return;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < nchains; c += (long long)gridDim.x * blockDim.x)
    ist_status[c] |= f;
  (void)pitch;
}

}  // namespace mcmcb
