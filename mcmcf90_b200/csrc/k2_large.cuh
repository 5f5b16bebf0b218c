// K2: large-npar sampler, one warp per chain, run-time npar (<= 256).
//
// The proposal factor no longer fits in registers (d=100: 5050 doubles), so it lives in HBM,
// one private d x d matrix per chain, and is streamed through the warp once per proposal:
// row-major upper triangle, lanes own output columns j = lane + 32 m, the loop runs over rows
// i, so every row segment R(i, i..d-1) is one coalesced read and no reduction is needed
// (p_j = sum_{i<=j} R(i,j) z_i accumulates in registers).  Vectors (theta, proposal, z) sit in
// shared memory, one slab per warp; the user-model blob is TMA-staged once per CTA.
//
// What differs from the reference's arithmetic (all rounding-level):
//  * second-stage proposal uses (R'z)/drscale instead of a stored R2 = R/drscale;
//  * the DR ratio needs no inverse covariance: with y1 = x + R'z1, y2 = x + R'z2/drscale,
//    (y2-y1)' iC (y2-y1) = |z2/drscale - z1|^2 and (x-y1)' iC (x-y1) = |z1|^2 exactly
//    (MCMC_DRAM.F90:181-183 with iC = inv(R'R), MCMC_adapt.F90:216-224) -- no dpotri, no iC
//    traffic, and more accurate than forming iC;
//  * normals are generated 32 candidate pairs at a time (Philox is counter based) and compacted
//    in stream order, which consumes exactly the draws the sequential polar loop would.
//
// Adaptation runs in a separate kernel (k2_adapt_kernel, one CTA per chain) at the ticks of
// MCMC_adapt.F90:45-46: accepted rows are logged with their weights into a per-chain row buffer
// by the step kernel and replayed through the covmat recursion (matutils.F90:283-310) in
// order, then dpotf2-ordered Cholesky writes the new factor.
#pragma once
#include "common.cuh"

namespace mcmcb {

#ifndef MCMCB_K2_MAX_WARPS
#define MCMCB_K2_MAX_WARPS 16
#endif
constexpr int K2_MAX_WARPS = MCMCB_K2_MAX_WARPS;       // warps (= chains in flight) per CTA: 16, or 8 when the resident factor needs the room
constexpr int K2_MAX_THREADS = K2_MAX_WARPS * 32;
constexpr int K2_MAXM = 8;             // npar <= 32 * K2_MAXM
constexpr int K2_NVEC = 6;             // per-warp shared vectors
constexpr int K2_ADAPT_THREADS = 256;

struct K2Layout {  // scalar SoA fields
  int ss, pri, s2, wsum, spare, rama, nf;
  int i_stayed, i_bnd, i_dracc, i_drtry, i_chainind, i_simuind, i_status, i_hasspare, i_cnt, i_pend, i_ndlo, i_ndhi,
      i_nbuf, i_er, i_nf;
};
__host__ __device__ constexpr K2Layout k2_layout(int NY) {
  K2Layout l{};
  int o = 0;
  l.ss = o; o += NY;
  l.pri = o; o += 1;
  l.s2 = o; o += NY;
  l.wsum = o; o += 1;
  l.spare = o; o += 1;
  l.rama = o; o += 1;
  l.nf = o;
  int k = 0;
  l.i_stayed = k++; l.i_bnd = k++; l.i_dracc = k++; l.i_drtry = k++; l.i_chainind = k++; l.i_simuind = k++;
  l.i_status = k++; l.i_hasspare = k++; l.i_cnt = k++; l.i_pend = k++; l.i_ndlo = k++; l.i_ndhi = k++;
  l.i_nbuf = k++;
  l.i_er = k++;  // erstayed, mcmc.F90:49
  l.i_nf = k;
  return l;
}

struct K2Params {
  DevCfg c;
  long long nchains, pitch, chain_offset;
  unsigned long long seed;
  int nsteps, d, dp;  // dp = d rounded up to a multiple of 32
  double* st;
  int* ist;
  double* theta;   // [chain][dp]
  double* mean;    // [chain][dp]
  double* Rm;      // [chain][d*d]  row-major upper factor (Cholesky mode)
  double* cmat;    // [chain][d*d]  symmetric, full
  double* rowbuf;  // [chain][cap][d+1]  completed rows + weights since the last adaptation
  double* coef;    // [chain][2 * (cap + 1)] covmat recursion coefficients of the logged rows (tick kernels' scratch)
  int rowcap;
  const double* par0;      // [chain][d]
  const double* cmat0;     // [d*d] column-major full
  const double* sigma2_0;
  const int* nobs;
  const double* blob;
  unsigned long long blob_n;
  unsigned blob_bytes;
  const double* prior;
  const double* inj;
  unsigned long long inj_per_chain;
  int store_chains, store_rows;
  double* store_rows_p;
  double* store_cnt_p;
  double* store_s2_p;
  unsigned int* tile_counter;
  int tick_i;      // adaptation kernel: the step index i of this tick
  double* qstd;    // [chain][dp]  SCAM proposal standard deviations (k3_scam.cuh)
  int factor_mode; // 0 = row-major upper Cholesky factor, 1 = column-major SVD factor (usesvd), 2 = SCAM
  long long r_stride, q_stride;  // per-chain strides of Rm / qstd: d*d / dp, or 0 when every chain shares the pooled factor
  double* gcm;     // [chain][d*d]  greedy burn-in: unit-weight covariance of the CLOSED rows so far (MCMC_adapt.F90:83-101)
  double* gmean;   // [chain][dp]
  double* gw;      // [chain]
  int r_resident;  // 1: the warp keeps its chain's factor in shared memory for the whole launch (d*d doubles per warp)
  int absorbed;    // tick kernels: 1 = k2_absorb_resident_kernel has already folded the logged rows into (wsum, mean, cmat)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
  return v;
}

// n normals in the reference's draw order (mcmcrand.F90:60-83,166-190), produced by the whole
// warp: lane l examines candidate pair l of the stream; accepted pairs are compacted in order.
__device__ __forceinline__ void warp_normals(Rng& g, double* zs, int n, int lane) {
  int filled = 0;
  if (g.has_spare && n > 0) {
    if (lane == 0) zs[0] = g.spare;
    g.has_spare = false;
    filled = 1;
  }
  while (filled < n) {
    // lane l looks 2l draws ahead of the stream position.  With an injected stream the look-ahead may pass its end
    // although the pairs actually consumed do not: unavailable pairs are never read, and exhaustion is declared --
    // by the whole warp at once -- only if a pair that the sequential polar loop would have consumed is missing
    const bool avail = g.inj == nullptr || g.nd + 2ull * lane + 1ull < g.inj_n;
    double u1 = 0.5, u2 = 0.5;
    if (avail) g.uniform_pair_at(g.nd + 2ull * lane, u1, u2);
    const double x1 = 2.0 * u1 - 1.0, x2 = 2.0 * u2 - 1.0;
    const double xx = x1 * x1 + x2 * x2;
    const bool ok = avail && (xx < 1.0 && xx != 0.0);
    const unsigned m = __ballot_sync(FULL, ok);
    const unsigned mav = __ballot_sync(FULL, avail);
    const int need = (n - filled + 1) >> 1;  // accepted pairs still needed
    const int have = __popc(m);
    const int use = have < need ? have : need;
    const int span = have >= need ? (int)__fns(m, 0, need) + 1 : 32;  // candidate pairs consumed by this round
    if (mav != FULL && __ffs(~mav) - 1 < span) {  // injected stream ran out: fill with zeros, the caller reports the flag
      g.exhausted = 1;
      for (int k = filled + lane; k < n; k += 32) zs[k] = 0.0;
      filled = n;
      break;
    }
    const int rank = __popc(m & ((1u << lane) - 1u));
    double sp = 0.0;
    bool made_spare = false;
    if (ok && rank < use) {
      const double z = sqrt(-2.0 * log(xx) / xx);
      const int idx = filled + 2 * rank;
      zs[idx] = z * x2;                 // second of the pair is returned first (mcmcrand.F90:183-185)
      if (idx + 1 < n) zs[idx + 1] = z * x1;
      else { sp = z * x1; made_spare = true; }
    }
    const unsigned ms = __ballot_sync(FULL, made_spare);
    if (ms) {
      g.spare = __shfl_sync(FULL, sp, __ffs(ms) - 1);
      g.has_spare = true;
    }
    g.nd += 2ull * span;  // stop right after the last pair used
    filled += 2 * use;
  }
  __syncwarp();
}

// p_j = sum_{i<=j} R(i,j) v_i for the columns this lane owns; R row-major upper.
// Column group m (columns 32m .. 32m+31) is a rectangle of rows 0 .. 32m, where every lane is above the
// diagonal, on top of a 31-row triangle.  The rectangle -- most of the factor -- is two instructions per
// element (load, DFMA) with U row loads in flight per lane before the first FMA consumes one; predicated
// loads would be fused with their FMAs and issue one at a time, so out-of-range lanes of the last group
// load a clamped (valid) column and their sums are simply never used.  Every column still accumulates its
// rows in ascending order: results are bit-identical to the rolled reference loop.
template <int U, bool TRI>
__device__ __forceinline__ void tmv_load(double (&r)[U], const double* R, int d, int i, int jc) {
  if (TRI) {
#pragma unroll
    for (int u = 0; u < U; u++) r[u] = R[(size_t)(i + u) * d + max(jc, i + u)];
  } else {
    // rectangle: one pointer for the batch, rows d apart (32-bit element offsets: one multiply-add per row instead of a
    // 64-bit index computation)
    const double* p = R + ((size_t)i * d + jc);
#pragma unroll
    for (int u = 0; u < U; u++) r[u] = p[(unsigned)(u * d)];
  }
}
template <int U, bool TRI>
__device__ __forceinline__ void tmv_use(const double (&r)[U], const double* vs, int i, int j, double& acc) {
#pragma unroll
  for (int u = 0; u < U; u++) {
    const double t = fma(r[u], vs[i + u], acc);
    acc = (!TRI || i + u <= j) ? t : acc;
  }
}
// nb batches of U rows starting at row i, software pipelined: the loads of the next batch are issued before
// the FMAs of the current one, so a lane keeps U..2U row loads in flight across the whole column
template <int U, bool TRI>
__device__ __forceinline__ void tmv_batches(const double* R, const double* vs, int d, int j, int jc, int i, int nb,
                                            double& acc) {
  if (nb <= 0) return;
  double r0[U], r1[U];
  tmv_load<U, TRI>(r0, R, d, i, jc);
  int b = 0;
  while (b + 2 <= nb) {
    tmv_load<U, TRI>(r1, R, d, i + (b + 1) * U, jc);
    tmv_use<U, TRI>(r0, vs, i + b * U, j, acc);
    if (b + 2 < nb) tmv_load<U, TRI>(r0, R, d, i + (b + 2) * U, jc);
    tmv_use<U, TRI>(r1, vs, i + (b + 1) * U, j, acc);
    b += 2;
  }
  if (b < nb) tmv_use<U, TRI>(r0, vs, i + b * U, j, acc);
}

template <int U>
__device__ __forceinline__ double tri_matvec_col(const double* R, const double* vs, int d, int j, int jc) {
  // j = this lane's column (may be >= d), jc = min(j, d - 1); returns sum_{i <= min(j, d-1)} R(i, jc) v_i
  const int m32 = j & ~31;                      // first column of the group = last fully rectangular row
  const int nrect = min(m32 + 1, d);            // rows 0 .. nrect-1: all 32 lanes active
  const int iend = min(m32 + 32, d);            // rows nrect .. iend-1: the triangle, lane active while row <= column
  double acc = 0.0;
  const int nb = nrect / U;
  tmv_batches<U, false>(R, vs, d, j, jc, 0, nb, acc);
  int i = nb * U;
  // the rows left over from the rectangle join the triangle: same arithmetic, the select is a no-op for them
  const int nbt = (iend - i) / U;
  tmv_batches<U, true>(R, vs, d, j, jc, i, nbt, acc);
  for (i += nbt * U; i < iend; i++) {
    const double t = fma(R[(size_t)i * d + max(jc, i)], vs[i], acc);
    acc = (i <= j) ? t : acc;
  }
  return acc;
}

// not inlined: the pipelined loops get their own register allocation instead of competing with the step
// kernel's long-lived state (the call happens once per proposal)
static __device__ __noinline__ void tri_matvec_t(const double* R, const double* vs, int d, int lane,
                                          double (&acc)[K2_MAXM]) {
  const int mm = (d + 31) >> 5;
#pragma unroll
  for (int m = 0; m < K2_MAXM; m++) {
    acc[m] = 0.0;
    if (m < mm) {
      const int j = lane + 32 * m;
      acc[m] = tri_matvec_col<8>(R, vs, d, j, min(j, d - 1));
    }
  }
}

template <class M>
__global__ void k2_init_kernel(K2Params p) {
  constexpr int NY = M::NY;
  constexpr K2Layout Lo = k2_layout(NY);
  const long long c = blockIdx.x;
  if (c >= p.nchains) return;
  const int d = p.d;
  for (int k = threadIdx.x; k < p.dp; k += blockDim.x) {
    double v = k < d ? p.par0[c * d + k] : 0.0;
    p.theta[c * p.dp + k] = v;
    p.mean[c * p.dp + k] = v;
    if (p.gmean) p.gmean[c * p.dp + k] = v;
  }
  for (int k = threadIdx.x; k < d * d; k += blockDim.x) {
    p.cmat[(size_t)c * d * d + k] = p.cmat0[k];
    if (p.gcm) p.gcm[(size_t)c * d * d + k] = p.cmat0[k];
    p.Rm[(size_t)c * p.r_stride + k] = 0.0;
  }
  if (threadIdx.x == 0) {
    double* st = p.st + c;
    int* ist = p.ist + c;
    for (int k = 0; k < NY; k++) { st[(Lo.ss + k) * p.pitch] = 0.0; st[(Lo.s2 + k) * p.pitch] = p.sigma2_0[k]; }
    st[Lo.pri * p.pitch] = 0.0;
    st[Lo.wsum * p.pitch] = (double)p.c.initcmatn;
    if (p.gw) p.gw[c] = (double)p.c.initcmatn;
    st[Lo.spare * p.pitch] = 0.0;
    st[Lo.rama * p.pitch] = 0.0;
    for (int k = 0; k < Lo.i_nf; k++) ist[k * p.pitch] = 0;
  }
}

// dpotf2('U') order on a row-major copy (A(i,k) at i*d+k); whole CTA; returns 0 or failing column+1.
// Only the upper triangle is read or written.
__device__ __forceinline__ int cta_cholesky_rowmajor(double* A, int d, double* red /* >= blockDim/32 doubles */) {
  __shared__ int fail;
  if (threadIdx.x == 0) fail = 0;
  __syncthreads();
  for (int j = 0; j < d; j++) {
    double t = 0.0;
    for (int i = threadIdx.x; i < j; i += blockDim.x) { double a = A[(size_t)i * d + j]; t = fma(a, a, t); }
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
      double ajj = A[(size_t)j * d + j] - s;
      if (!(ajj > 0.0)) fail = j + 1;
      red[0] = sqrt(ajj);
    }
    __syncthreads();
    if (fail) return fail;
    const double ajj = red[0];
    const double rajj = 1.0 / ajj;
    for (int k = j + threadIdx.x; k < d; k += blockDim.x) {
      if (k == j) { A[(size_t)j * d + j] = ajj; continue; }
      double tt = 0.0;
      for (int i = 0; i < j; i++) tt = fma(A[(size_t)i * d + k], A[(size_t)i * d + j], tt);
      A[(size_t)j * d + k] = (A[(size_t)j * d + k] - tt) * rajj;
    }
    __syncthreads();
  }
  return 0;
}

// R = chol(cm) * 2.4/sqrt(d) into row-major Rm; cm is full symmetric (either order).  Scratch = Rm itself:
// on failure the old factor must survive (MCMC_adapt.F90:169-171), so work in `tmp` and copy on success.
__device__ __forceinline__ bool cta_calculate_R(const double* cm, double* Rm, double* tmp, int d, double* red) {
  for (int k = threadIdx.x; k < d * d; k += blockDim.x) tmp[k] = cm[k];
  __syncthreads();
  const int info = cta_cholesky_rowmajor(tmp, d, red);
  if (info) return false;
  const double sq = sqrt((double)d);
  for (int k = threadIdx.x; k < d * d; k += blockDim.x) {
    const int i = k / d, j = k - i * d;
    Rm[k] = (j >= i) ? tmp[k] * 2.4 / sq : 0.0;
  }
  __syncthreads();
  return true;
}

// Initial factor, MCMC_init.F90:108-110.  One CTA per chain; tmp = that chain's slice of a scratch buffer.
static __global__ void k2_initR_kernel(K2Params p, double* scratch) {
  __shared__ double red[K2_ADAPT_THREADS / 32];
  const long long c = blockIdx.x;
  const int d = p.d;
  constexpr K2Layout Lo = k2_layout(1);
  bool ok = cta_calculate_R(p.cmat + (size_t)c * d * d, p.Rm + (size_t)c * p.r_stride, scratch + (size_t)c * d * d, d, red);
  if (!ok && threadIdx.x == 0) p.ist[Lo.i_status * p.pitch + c] |= MCMCB_ST_CHOLFAIL;
}

// The covmat recursion (matutils.F90:283-310) over all logged rows of one chain by the whole CTA.
//
// The reference applies one rank-1 update per row to the whole matrix: cm <- cm + f1 (f2 dv dv' - cm) with
// dv = x - mean, then mean <- mean + f3 dv.  Doing that row by row is one read+write pass over cm per row
// (d*d*16 bytes x ~adaptint rows: the tick then costs more than the sampling between ticks).  Here:
//  phase 1 runs the cheap O(d) part of every row in order -- dv (stored over the row in the row buffer),
//          the mean, wsum and the per-row coefficients (shared memory) -- thread k owns mean[k];
//  phase 2 keeps a tile of cm entries in registers and applies the rows' updates to it in order, reading
//          the dv rows from shared-memory chunks, so cm is read and written ONCE per tick.
// Every entry sees exactly the reference's sequence of operations, so the result is bit-identical to the
// row-by-row loop.  Only the upper triangle is computed; the mirror image is written beside it.
constexpr int ABS_E = 8;    // cm entries per thread and tile
constexpr int ABS_RC = 8;   // dv rows per shared-memory chunk

// shared scratch: chunk[ABS_RC * d] -- O(d), independent of the row capacity (the per-row coefficients live in global
// memory, K2Params::coef: rowcap grows with burnintime and would otherwise pass the 48 KB / 227 KB launch limits)
__host__ __device__ __forceinline__ size_t absorb_smem_doubles(int d) { return (size_t)ABS_RC * d; }

__device__ __forceinline__ void cta_absorb_rows(double* rb, int nrows, double* cm, double* mean, double& wsum, int d,
                                                double* coef, double* chunk, bool unit_weights = false) {
  // coef: (f1, f2) per row, global; f2 = -1: no-op row, f2 = -2: reset (first row, wsum == 0).  chunk: ABS_RC x d, shared
  const int tid = threadIdx.x, nt = blockDim.x;
  // ---- phase 1: rows in order; every thread tracks wsum, thread k owns component k
  double ws = wsum;
  for (int r = 0; r < nrows; r++) {
    double* x = rb + (size_t)r * (d + 1);
    const double w = unit_weights ? 1.0 : x[d];
    if (ws > 0.0) {
      const double f3 = w / (ws + w);
      for (int k = tid; k < d; k += nt) {
        const double dv = x[k] - mean[k];
        mean[k] = mean[k] + f3 * dv;
        x[k] = dv;
      }
      if (tid == 0) { coef[2 * r] = w / (ws + w - 1.0); coef[2 * r + 1] = ws / (ws + w); }
      ws = w + ws;
    } else if (w > 0.0) {
      for (int k = tid; k < d; k += nt) mean[k] = x[k];
      if (tid == 0) { coef[2 * r] = 0.0; coef[2 * r + 1] = -2.0; }
      ws = w;
    } else {
      if (tid == 0) { coef[2 * r] = 0.0; coef[2 * r + 1] = -1.0; }
    }
  }
  wsum = ws;
  __syncthreads();
  // ---- phase 2: the upper triangle in registers, rows streamed through shared memory.  A thread owns runs
  // of ABS_E consecutive rows a0 .. a0+7 of one column b, so one row of the recursion costs it one dv[b] and
  // eight consecutive dv[a] from shared memory for eight entry updates.
  int nchunk = 0;
  for (int b = 0; b < d; b++) nchunk += (b + ABS_E) / ABS_E;  // ceil((b + 1) / ABS_E) runs in column b
  for (int c0 = 0; c0 < nchunk; c0 += nt) {
    const int cidx = c0 + tid;
    int b = 0, a0 = 0;
    bool live = cidx < nchunk;
    if (live) {  // run index -> (column, first row): walk the columns (d <= 256 steps, once per tick)
      int left = cidx;
      for (;;) {
        const int nb = (b + ABS_E) / ABS_E;
        if (left < nb) break;
        left -= nb;
        b++;
      }
      a0 = left * ABS_E;
    }
    double v[ABS_E];
#pragma unroll
    for (int q = 0; q < ABS_E; q++) v[q] = (live && a0 + q <= b) ? cm[(size_t)b * d + a0 + q] : 0.0;  // symmetric: (b,a) == (a,b)
    for (int r0 = 0; r0 < nrows; r0 += ABS_RC) {
      const int nr = min(ABS_RC, nrows - r0);
      __syncthreads();
      for (int k = tid; k < nr * d; k += nt) chunk[k] = rb[(size_t)(r0 + k / d) * (d + 1) + (k % d)];
      __syncthreads();
      if (live) {
        for (int r = 0; r < nr; r++) {
          const double f1 = coef[2 * (r0 + r)], f2 = coef[2 * (r0 + r) + 1];
          const double* dv = chunk + (size_t)r * d;
          if (f2 >= 0.0) {
            const double db = dv[b];
            // rows past the diagonal (a0 + q > b) of the last run read in-bounds garbage that is never stored
#pragma unroll
            for (int q = 0; q < ABS_E; q++) v[q] = v[q] + f1 * (f2 * (dv[min(a0 + q, d - 1)] * db) - v[q]);
          } else if (f2 == -2.0) {
#pragma unroll
            for (int q = 0; q < ABS_E; q++) v[q] = 0.0;
          }
        }
      }
    }
    if (live) {
#pragma unroll
      for (int q = 0; q < ABS_E; q++)
        if (a0 + q <= b) {
          cm[(size_t)(a0 + q) * d + b] = v[q];
          cm[(size_t)b * d + a0 + q] = v[q];
        }
    }
  }
  __syncthreads();
}

// AP (MCMC_adapt.F90:116-136): covariance of the last adapthist steps by the batch formula of covmat
// (matutils.F90:312-337), whole CTA.  The row buffer holds the closed rows (theta, repeat count) still inside the
// window; the open row is appended as row nbuf.  The first row's weight is trimmed so that the weights add up to
// adapthist.  Afterwards the closed rows that can still fall inside the next window sit at the front of the buffer.
__device__ __forceinline__ void cta_ap_window(double* rb, int nbuf, const double* theta, double* cm, double* mean,
                                              double* st, int* ist, long long pitch, int d, int adapthist) {
  constexpr K2Layout Lo = k2_layout(1);
  __shared__ int s_first, s_hist;
  for (int k = threadIdx.x; k < d; k += blockDim.x) rb[(size_t)nbuf * (d + 1) + k] = theta[k];
  if (threadIdx.x == 0) {
    rb[(size_t)nbuf * (d + 1) + d] = (double)ist[Lo.i_cnt * pitch];
    int first = nbuf, histsum = ist[Lo.i_cnt * pitch];
    while (histsum < adapthist && first > 0) { first--; histsum += (int)rb[(size_t)first * (d + 1) + d]; }
    s_first = first; s_hist = histsum;
  }
  __syncthreads();
  const int first = s_first, nrow = nbuf - first + 1;
  const double wfirst = (double)((int)rb[(size_t)first * (d + 1) + d] - s_hist + adapthist);
  double wsum2 = 0.0;
  for (int r = 0; r < nrow; r++) wsum2 += (r == 0) ? wfirst : rb[(size_t)(first + r) * (d + 1) + d];
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < nrow; r++) acc = acc + rb[(size_t)(first + r) * (d + 1) + k] * ((r == 0) ? wfirst : rb[(size_t)(first + r) * (d + 1) + d]);
    mean[k] = acc / wsum2;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int a = e / d, b = e - a * d;  // entry (a, b), b <= a computed, mirrored
    if (b > a) continue;
    const double ma = mean[a], mb2 = mean[b];
    double acc = 0.0;
    for (int r = 0; r < nrow; r++) {
      const double* x = rb + (size_t)(first + r) * (d + 1);
      acc = acc + (x[a] - ma) * ((x[b] - mb2) * ((r == 0) ? wfirst : x[d]));
    }
    const double v = acc / (wsum2 - 1.0);
    cm[(size_t)a * d + b] = v;
    cm[(size_t)b * d + a] = v;
  }
  __syncthreads();
  const int keep = nbuf - first;  // closed rows first .. nbuf-1
  if (first > 0) {
    for (int r = 0; r < keep; r++) {
      for (int k = threadIdx.x; k <= d; k += blockDim.x) rb[(size_t)r * (d + 1) + k] = rb[(size_t)(first + r) * (d + 1) + k];
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    st[Lo.wsum * pitch] = wsum2;
    ist[Lo.i_nbuf * pitch] = keep;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// The same recursion with EVERY logged row of the chain resident in shared memory (one CTA per SM).
//
// cta_absorb_rows streams the rows through an 8-row window: at npar = 200 that is 11 passes over the row buffer, 286
// barriers and as many exposed L2 round trips per chain, and the tick of 262 144 SCAM chains took 1.02 s where the FP64
// work is ~0.09 s (profiles/r02_summary.md).  Here the CTA loads the chain's whole row buffer once ((adaptint + 1) x npar
// doubles: 162 KB at BASELINE C5, 161 KB at C2), runs phase 1 out of shared memory with no barrier (thread k owns
// component k of every row), and phase 2 on 4 x 4 register tiles of the upper triangle: 8 shared-memory doubles per
// row feed 16 entry updates (48 FP64 instructions), so the FP64 pipe and not the shared-memory port bounds it, there is no
// barrier inside, and both images of a tile are written as 32-byte runs.  Every entry sees the reference's sequence of
// operations (matutils.F90:283-310): bit-identical to cta_absorb_rows, checked by tests/test_r02_coverage.py.
constexpr int K2_ABSR_THREADS = 640;  // npar = 200: 1275 tiles = two passes of 640 threads at 99.6 %

// shared memory: rows[nrows * d] | weight[nrows] | coef[2 * nrows], nrows = rowcap + 1
__host__ __device__ __forceinline__ size_t absorb_resident_smem_bytes(int d, int rowcap) {
  const size_t nrows = (size_t)rowcap + 1;
  return sizeof(double) * (nrows * d + 3 * nrows + 2);
}

template <bool VEC>
__device__ __forceinline__ void absr_tile(const double* __restrict__ X, const double* __restrict__ coef, int nrows, double* cm,
                                          int d, int a0, int b0) {
  int ai[4], bj[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { ai[k] = VEC ? a0 + k : min(a0 + k, d - 1); bj[k] = VEC ? b0 + k : min(b0 + k, d - 1); }
  double v[4][4];  // v[i][j] = entry (a0 + i, b0 + j), read from its mirror image (b, a): contiguous in i
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) v[i][j] = cm[(size_t)bj[j] * d + ai[i]];
#pragma unroll 2
  for (int r = 0; r < nrows; r++) {
    const double f1 = coef[2 * r], f2 = coef[2 * r + 1];
    if (f2 >= 0.0) {
      const double* x = X + (size_t)r * d;
      double da[4], db[4];
      if (VEC) {  // d % 4 == 0: whole tiles, 16-byte aligned pairs
        const double2 p0 = *reinterpret_cast<const double2*>(x + a0), p1 = *reinterpret_cast<const double2*>(x + a0 + 2);
        const double2 q0 = *reinterpret_cast<const double2*>(x + b0), q1 = *reinterpret_cast<const double2*>(x + b0 + 2);
        da[0] = p0.x; da[1] = p0.y; da[2] = p1.x; da[3] = p1.y;
        db[0] = q0.x; db[1] = q0.y; db[2] = q1.x; db[3] = q1.y;
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++) { da[k] = x[ai[k]]; db[k] = x[bj[k]]; }
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) v[i][j] = v[i][j] + f1 * (f2 * (da[i] * db[j]) - v[i][j]);
    } else if (f2 == -2.0) {
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) v[i][j] = 0.0;
    }
  }
  // entries below the diagonal of a diagonal tile and past the edge of a ragged one hold garbage: never stored
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (a0 + i <= b0 + j && b0 + j < d) cm[(size_t)(a0 + i) * d + b0 + j] = v[i][j];
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (a0 + i <= b0 + j && b0 + j < d) cm[(size_t)(b0 + j) * d + a0 + i] = v[i][j];
}

// One CTA per chain: the plain adaptation branch (MCMC_adapt.F90:105-159 with adapthist <= 1) up to, not including,
// the factorisation.  The host launches it only at such ticks (K2Launcher::tick_is_plain) and only when the rows fit.
static __global__ void __launch_bounds__(K2_ABSR_THREADS) k2_absorb_resident_kernel(K2Params p) {
  extern __shared__ __align__(16) double sm_absr[];
  constexpr K2Layout Lo = k2_layout(1);
  const long long c = blockIdx.x;
  const int d = p.d, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  double* st = p.st + c;
  int* ist = p.ist + c;
  double* cm = p.cmat + (size_t)c * d * d;
  double* gmean = p.mean + c * p.dp;
  const double* theta = p.theta + c * p.dp;
  const double* rb = p.rowbuf + (size_t)c * (p.rowcap + 1) * (d + 1);
  const int nbuf = ist[Lo.i_nbuf * p.pitch], nrows = nbuf + 1;
  double* X = sm_absr;
  double* wt = X + (size_t)(p.rowcap + 1) * d;
  double* coef = wt + (p.rowcap + 1) + ((p.rowcap + 1) & 1);  // 16-byte aligned
  // ---- load: the logged rows (cp.async, 8 bytes each: row r starts at an odd multiple of 8 bytes when npar is even, so
  // no bulk copy; nothing waits until every copy is in flight), then the open row with its pending weight
  for (int r = warp; r < nbuf; r += nwarps) {
    const double* x = rb + (size_t)r * (d + 1);
    for (int k = lane; k <= d; k += 32) cp_async_8(k < d ? X + (size_t)r * d + k : wt + r, x + k);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int k = tid; k < d; k += nt) X[(size_t)nbuf * d + k] = theta[k];
  if (tid == 0) wt[nbuf] = (double)ist[Lo.i_pend * p.pitch];
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- phase 1a: the weight recursion.  One thread runs the additions in row order (wsum before row r -> coef[2r]);
  // then thread r turns row r's pair (w, wsum) into the recursion's three quotients, all rows at once:
  // coef[2r] = f1 = w / (wsum + w - 1), coef[2r+1] = f2 = wsum / (wsum + w) (-2: reset row, -1: no-op row), wt[r] = f3 =
  // w / (wsum + w).  Same expressions as cta_absorb_rows, evaluated once per row instead of once per row and thread.
  double ws = st[Lo.wsum * p.pitch];
  if (tid == 0) {
    for (int r = 0; r < nrows; r++) {
      const double w = wt[r];
      coef[2 * r] = ws;
      if (ws > 0.0) ws = w + ws;
      else if (w > 0.0) ws = w;
    }
    st[Lo.wsum * p.pitch] = ws;
    ist[Lo.i_pend * p.pitch] = 0;
    ist[Lo.i_nbuf * p.pitch] = 0;
  }
  __syncthreads();
  for (int r = tid; r < nrows; r += nt) {
    const double w = wt[r], w0 = coef[2 * r];
    if (w0 > 0.0) {
      coef[2 * r] = w / (w0 + w - 1.0);
      coef[2 * r + 1] = w0 / (w0 + w);
      wt[r] = w / (w0 + w);
    } else {
      coef[2 * r] = 0.0;
      coef[2 * r + 1] = w > 0.0 ? -2.0 : -1.0;
    }
  }
  __syncthreads();
  // ---- phase 1b: rows in order, thread k owns component k (mean in a register); no barrier
  for (int k = tid; k < d; k += nt) {
    double m = gmean[k];
    for (int r = 0; r < nrows; r++) {
      const double f2 = coef[2 * r + 1];
      double* x = X + (size_t)r * d + k;
      if (f2 >= 0.0) {
        const double f3 = wt[r];
        const double dv = *x - m;
        m = m + f3 * dv;
        *x = dv;
      } else if (f2 == -2.0) {
        m = *x;
      }
    }
    gmean[k] = m;
  }
  __syncthreads();
  // ---- phase 2: 4 x 4 tiles of the upper triangle, tile t = jb (jb + 1) / 2 + ib, ib <= jb
  const int nb = (d + 3) >> 2, ntiles = nb * (nb + 1) / 2;
  const bool vec = (d & 3) == 0;
  for (int t = tid; t < ntiles; t += nt) {
    int jb = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while (jb * (jb + 1) / 2 > t) jb--;
    while ((jb + 1) * (jb + 2) / 2 <= t) jb++;
    const int ib = t - jb * (jb + 1) / 2;
    if (vec) absr_tile<true>(X, coef, nrows, cm, d, 4 * ib, 4 * jb);
    else absr_tile<false>(X, coef, nrows, cm, d, 4 * ib, 4 * jb);
  }
}

// MCMC_adapt.F90:12-174 at step index p.tick_i, one CTA per chain.
static __global__ void k2_adapt_kernel(K2Params p, double* scratch) {
  extern __shared__ double sh[];  // absorb_smem_doubles(d)
  double* coef = p.coef + (size_t)blockIdx.x * 2 * (p.rowcap + 1);
  __shared__ double red[K2_ADAPT_THREADS / 32];
  constexpr K2Layout Lo = k2_layout(1);
  const long long c = blockIdx.x;
  const DevCfg& cf = p.c;
  const int d = p.d, i = p.tick_i;
  double* st = p.st + c;
  int* ist = p.ist + c;
  double* cm = p.cmat + (size_t)c * d * d;
  double* Rm = p.Rm + (size_t)c * p.r_stride;
  double* mean = p.mean + c * p.dp;
  double* theta = p.theta + c * p.dp;
  double* rb = p.rowbuf + (size_t)c * (p.rowcap + 1) * (d + 1);
  // the factorisation's work matrix stays in this chain's slice of the global scratch (L2): a d x d copy in shared memory
  // (80 KB at d = 100) was measured SLOWER on BASELINE C2 -- 10.5 ms per tick against 7.5 -- because it leaves two CTAs per
  // SM where eight hide each other's barrier latency (profiles/r02_summary.md)
  double* tmp = scratch + (size_t)c * d * d;
  const int ma = cf.adaptint > 0 ? i % cf.adaptint : 1;
  const int mb = cf.badaptint > 0 ? i % cf.badaptint : 1;
  if (ma != 0 && mb != 0) return;
  double wsum = st[Lo.wsum * p.pitch];
  const int nbuf = ist[Lo.i_nbuf * p.pitch];
  if (i < cf.burnintime && cf.doburnin && mb == 0) {  // MCMC_adapt.F90:60-102
    const double staypc = (double)ist[Lo.i_stayed * p.pitch] / (double)i;
    if (staypc > 1.0 - cf.scalelimit) {
      for (int k = threadIdx.x; k < d * d; k += blockDim.x) Rm[k] = Rm[k] / cf.scalefactor;
    } else if (staypc < cf.scalelimit) {
      for (int k = threadIdx.x; k < d * d; k += blockDim.x) Rm[k] = Rm[k] * cf.scalefactor;
    } else if (cf.greedy) {
      // MCMC_adapt.F90:83-102: chaincmat = covariance of rows 1..chainind with unit weights on top of (cmat0, par0,
      // initcmatn).  The greedy accumulators hold the rows closed before the last greedy tick; the rows closed since
      // are in the row buffer, and the open row joins for this tick only (it is fed again, once, when it closes).
      double* gcm = p.gcm + (size_t)c * d * d;
      double* gmean = p.gmean + c * p.dp;
      double gw = p.gw[c];
      __syncthreads();
      cta_absorb_rows(rb, nbuf, gcm, gmean, gw, d, coef, sh, true);
      for (int k = threadIdx.x; k < d * d; k += blockDim.x) cm[k] = gcm[k];
      for (int k = threadIdx.x; k < d; k += blockDim.x) { mean[k] = gmean[k]; rb[k] = theta[k]; }
      if (threadIdx.x == 0) rb[d] = 1.0;
      __syncthreads();
      wsum = gw;
      cta_absorb_rows(rb, 1, cm, mean, wsum, d, coef, sh, true);
      const bool ok = cta_calculate_R(cm, Rm, tmp, d, red);
      // what the AM branch starts from: see k1_finish (the reference resets at simuind == burnintime+adaptint+adapthist
      // only if that step is a tick, otherwise it carries on from the greedy covariance)
      const int t0 = cf.burnintime + cf.adaptint + cf.adapthist;
      const bool will_reset = cf.doadapt && ((cf.adaptint > 0 && t0 % cf.adaptint == 0) || (cf.badaptint > 0 && t0 % cf.badaptint == 0)) &&
                              !(cf.adaptend > 0 && t0 > cf.adaptend);
      __syncthreads();
      if (will_reset) {
        for (int k = threadIdx.x; k < d * d; k += blockDim.x) cm[k] = p.cmat0[k];
        for (int k = threadIdx.x; k < d; k += blockDim.x) mean[k] = p.par0[c * d + k];
        wsum = (double)cf.initcmatn;
      }
      if (threadIdx.x == 0) {
        p.gw[c] = gw;
        st[Lo.wsum * p.pitch] = wsum;
        ist[Lo.i_pend * p.pitch] = 0;  // lastfreq = count of the open row
        ist[Lo.i_nbuf * p.pitch] = 0;
        if (!ok) ist[Lo.i_status * p.pitch] |= MCMCB_ST_CHOLFAIL;
      }
    } else {
      // restart the accumulation at the open row with its full count; R = chol(cmat0) (see K1)
      for (int k = threadIdx.x; k < d * d; k += blockDim.x) cm[k] = p.cmat0[k];
      for (int k = threadIdx.x; k < d; k += blockDim.x) mean[k] = p.par0[c * d + k];
      __syncthreads();
      if (threadIdx.x == 0) {
        st[Lo.wsum * p.pitch] = (double)cf.initcmatn;
        ist[Lo.i_pend * p.pitch] = ist[Lo.i_cnt * p.pitch];
        ist[Lo.i_nbuf * p.pitch] = 0;
      }
      if (!cta_calculate_R(cm, Rm, tmp, d, red) && threadIdx.x == 0) ist[Lo.i_status * p.pitch] |= MCMCB_ST_CHOLFAIL;
    }
  } else if (i >= cf.burnintime + cf.adaptint + cf.adapthist && cf.doadapt && cf.adapthist > 1) {
    cta_ap_window(rb, nbuf, theta, cm, mean, st, ist, p.pitch, d, cf.adapthist);
    if (!cf.pool && !cta_calculate_R(cm, Rm, tmp, d, red) && threadIdx.x == 0) ist[Lo.i_status * p.pitch] |= MCMCB_ST_CHOLFAIL;
  } else if (i >= cf.burnintime + cf.adaptint + cf.adapthist && cf.doadapt) {  // MCMC_adapt.F90:105-159
    // the open row joins the logged rows with its pending weight (slot nbuf always exists: rowcap + 1 rows)
    if (!p.absorbed) {
      for (int k = threadIdx.x; k < d; k += blockDim.x) rb[(size_t)nbuf * (d + 1) + k] = theta[k];
      if (threadIdx.x == 0) rb[(size_t)nbuf * (d + 1) + d] = (double)ist[Lo.i_pend * p.pitch];
      __syncthreads();
      cta_absorb_rows(rb, nbuf + 1, cm, mean, wsum, d, coef, sh);
      if (threadIdx.x == 0) {
        st[Lo.wsum * p.pitch] = wsum;
        ist[Lo.i_pend * p.pitch] = 0;
        ist[Lo.i_nbuf * p.pitch] = 0;
      }
    }
    if (!cf.pool && !cta_calculate_R(cm, Rm, tmp, d, red) && threadIdx.x == 0) ist[Lo.i_status * p.pitch] |= MCMCB_ST_CHOLFAIL;
  }
}

// rank-1 update / downdate of the row-major upper factor by one warp (dchud.f:122-139, dchdd.f:141-179).
// x in shared (overwritten); wk = 2 more shared vectors.
__device__ __forceinline__ void warp_chud(double* R, double* x, int d, int lane) {
  for (int i = 0; i < d; i++) {
    double* row = R + (size_t)i * d;
    double c = 0.0, s = 0.0;
    if (lane == 0) {
      double rii = row[i], xi = x[i];
      drotg(rii, xi, c, s);
      row[i] = rii;
    }
    c = __shfl_sync(FULL, c, 0);
    s = __shfl_sync(FULL, s, 0);
    for (int j = i + 1 + lane; j < d; j += 32) {
      const double rij = row[j], xj = x[j];
      row[j] = c * rij + s * xj;
      x[j] = c * xj - s * rij;
    }
    __syncwarp();
  }
}

__device__ __forceinline__ bool warp_chdd(double* R, double* x, double* sv, double* cv, int d, int lane) {
  // solve R' a = x (forward substitution, row sweep), result in sv
  for (int k = lane; k < d; k += 32) sv[k] = x[k];
  __syncwarp();
  for (int i = 0; i < d; i++) {
    const double* row = R + (size_t)i * d;
    double ai = 0.0;
    if (lane == 0) { ai = sv[i] / row[i]; sv[i] = ai; }
    ai = __shfl_sync(FULL, ai, 0);
    for (int j = i + 1 + lane; j < d; j += 32) sv[j] -= row[j] * ai;
    __syncwarp();
  }
  // classic dnrm2 + rotation recurrences: short scalar loops, done redundantly by every lane
  double norm;
  if (d == 1) {
    norm = fabs(sv[0]);
  } else {
    double scale = 0.0, ssq = 1.0;
    for (int k = 0; k < d; k++) {
      const double v = sv[k];
      if (v != 0.0) {
        const double a = fabs(v);
        if (scale < a) { const double t = scale / a; ssq = 1.0 + ssq * t * t; scale = a; }
        else { const double t = a / scale; ssq = ssq + t * t; }
      }
    }
    norm = scale * sqrt(ssq);
  }
  if (!(norm < 1.0)) return false;
  double alpha = sqrt(1.0 - norm * norm);
  __syncwarp();
  // dchdd.f:157-165.  Only alpha is carried from row to row (one division and one square root on the
  // serial path); lane 0 runs that chain and leaves (alpha_i, scale_i * nr_i) behind, then every lane forms
  // c_i, s_i of its own rows with the reference's operations -- same values, 3 of the 5 divisions per row
  // off the serial path.
  if (lane == 0) {
    for (int i = d - 1; i >= 0; i--) {
      const double scale = alpha + fabs(sv[i]);
      const double a = alpha / scale, b = sv[i] / scale;
      const double nr = sqrt(a * a + b * b);
      cv[i] = alpha;  // alpha entering row i
      alpha = scale * nr;
    }
  }
  __syncwarp();
  for (int i = lane; i < d; i += 32) {
    const double al = cv[i];
    const double scale = al + fabs(sv[i]);
    const double a = al / scale, b = sv[i] / scale;
    const double nr = sqrt(a * a + b * b);
    cv[i] = a / nr;
    sv[i] = b / nr;
  }
  __syncwarp();
  // apply: per column j, xx runs from row j down to row 0; lanes own columns
  double xx[K2_MAXM];
#pragma unroll
  for (int m = 0; m < K2_MAXM; m++) xx[m] = 0.0;
  for (int i = d - 1; i >= 0; i--) {
    double* row = R + (size_t)i * d;
    const double c = cv[i], s = sv[i];
#pragma unroll
    for (int m = 0; m < K2_MAXM; m++) {
      const int j = lane + 32 * m;
      if (j >= i && j < d) {
        const double rij = row[j];
        const double t = c * xx[m] + s * rij;
        row[j] = c * rij - s * xx[m];
        xx[m] = t;
      }
    }
  }
  __syncwarp();
  return true;
}

// y_i = sum_j R(i,j) v_j for the rows this lane owns; R column-major general (dgemv 'N', matutils.F90:161,
// the usesvd proposal of MCMC_DRAM.F90:27).  U columns of loads in flight, same summation order.
__device__ __forceinline__ void gen_matvec_n(const double* R, const double* vs, int d, int lane,
                                             double (&acc)[K2_MAXM]) {
  constexpr int U = 8;
  const int mm = (d + 31) >> 5;
#pragma unroll
  for (int m = 0; m < K2_MAXM; m++) {
    acc[m] = 0.0;
    if (m < mm) {
      const int ic = min(lane + 32 * m, d - 1);  // clamped row: lanes past d compute a sum nobody reads
      double a = 0.0;
      int j = 0;
      for (; j + U <= d; j += U) {
        double r[U];
#pragma unroll
        for (int u = 0; u < U; u++) r[u] = R[(size_t)(j + u) * d + ic];
#pragma unroll
        for (int u = 0; u < U; u++) a = fma(r[u], vs[j + u], a);
      }
      for (; j < d; j++) a = fma(R[(size_t)j * d + ic], vs[j], a);
      acc[m] = a;
    }
  }
}

template <class M, bool SMEM>
__global__ void __launch_bounds__(K2_MAX_THREADS, 1) k2_step_kernel(const __grid_constant__ K2Params p) {
  constexpr int NY = M::NY;
  constexpr K2Layout Lo = k2_layout(NY);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  const int d = p.d, dp = p.dp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const DevCfg& c = p.c;

  // dynamic shared memory: [per-warp vectors][model blob]
  double* vecs = reinterpret_cast<double*>(smem_raw) + (size_t)warp * K2_NVEC * dp;
  double *th = vecs, *prop = vecs + dp, *z1 = vecs + 2 * dp, *z2 = vecs + 3 * dp, *w1 = vecs + 4 * dp,
         *w2 = vecs + 5 * dp;
  // dynamic shared memory: [per-warp vectors][per-warp resident factors (optional)][model blob (optional)]
  double* Rs = reinterpret_cast<double*>(smem_raw) + (size_t)nwarps * K2_NVEC * dp + (size_t)warp * d * d;
  const double* data = p.blob;
  if (SMEM) {
    unsigned char* blob_s = smem_raw + sizeof(double) * ((size_t)nwarps * K2_NVEC * dp +
                                                        (p.r_resident ? (size_t)nwarps * d * d : 0));
    tma_stage_blob(blob_s, p.blob, p.blob_bytes, &mbar);
    data = reinterpret_cast<const double*>(blob_s);
  }
  mcmcb_ctx ctx;
  ctx.data = data; ctx.ndata = p.blob_n; ctx.prior = p.prior; ctx.lane = lane; ctx.nlanes = 32;
  ctx.exp_tl = 0u; ctx.exp_c1 = MCMCB_EXP_C1L; ctx.exp_c2 = MCMCB_EXP_C2L;
  ctx.scratch = vecs + 5 * dp;  // w2: free while the model runs

  for (;;) {
    unsigned tile = 0;
    if (lane == 0) tile = atomicAdd(p.tile_counter, 1u);
    tile = __shfl_sync(FULL, tile, 0);
    if ((long long)tile >= p.nchains) break;
    const long long cc = tile;
    double* st = p.st + cc;
    int* ist = p.ist + cc;
    double* Rg = p.Rm + (size_t)cc * p.r_stride;
    double* Rm = Rg;
    double* gth = p.theta + cc * dp;
    double* rb = p.rowbuf + (size_t)cc * (p.rowcap + 1) * (d + 1);

    if (p.r_resident) {
      // the factor is read once per launch instead of once per proposal, and the RAM rank-1 updates run at
      // shared-memory latency; written back once at the end when RAM changed it
      for (int k = lane; k < d * d; k += 32) Rs[k] = Rg[k];
      Rm = Rs;
    }
    for (int k = lane; k < dp; k += 32) { th[k] = gth[k]; prop[k] = gth[k]; }
    double ss1[NY], s2[NY];
#pragma unroll
    for (int k = 0; k < NY; k++) { ss1[k] = st[(Lo.ss + k) * p.pitch]; s2[k] = st[(Lo.s2 + k) * p.pitch]; }
    double pri1 = st[Lo.pri * p.pitch], rama = st[Lo.rama * p.pitch];
    int stayed = ist[Lo.i_stayed * p.pitch], bnd = ist[Lo.i_bnd * p.pitch], dracc = ist[Lo.i_dracc * p.pitch];
    int drtry = ist[Lo.i_drtry * p.pitch], chainind = ist[Lo.i_chainind * p.pitch];
    int simuind = ist[Lo.i_simuind * p.pitch], status = ist[Lo.i_status * p.pitch];
    int cnt = ist[Lo.i_cnt * p.pitch], pend = ist[Lo.i_pend * p.pitch], nbuf = ist[Lo.i_nbuf * p.pitch];
    int erst = ist[Lo.i_er * p.pitch];
    double sscrit = 0.0;
    Rng g;
    g.nd = ((unsigned long long)(unsigned)ist[Lo.i_ndhi * p.pitch] << 32) | (unsigned)ist[Lo.i_ndlo * p.pitch];
    g.seed = p.seed; g.chain = (unsigned long long)(p.chain_offset + cc);
    g.inj = p.inj ? p.inj + (unsigned long long)cc * p.inj_per_chain : nullptr;
    g.inj_n = p.inj_per_chain;
    g.cache_valid = false; g.cache_lo = g.cache_hi = 0; g.cache_blk = 0;
    g.has_spare = ist[Lo.i_hasspare * p.pitch] != 0;
    g.spare = st[Lo.spare * p.pitch];
    g.exhausted = 0;
    const bool stored = (cc < p.store_chains);
    double* srow = p.store_rows_p + (size_t)cc * p.store_rows * (d + NY);
    double* scnt = p.store_cnt_p + (size_t)cc * p.store_rows;
    double* ss2st = p.store_s2_p + (size_t)cc * p.store_rows * NY;
    __syncwarp();

    int phase = (simuind == 0) ? -1 : 0;
    int done = 0;
    double ss2[NY], pri2 = 0.0, a12 = 0.0, z1sq = 0.0;
#pragma unroll
    for (int k = 0; k < NY; k++) ss2[k] = 0.0;

    while (phase < 0 || done < p.nsteps) {
      // ---------------- proposal
      bool inb = true;
      if (phase >= 0) {
        double* zs = (phase == 0) ? z1 : z2;
        warp_normals(g, zs, d, lane);
        double acc[K2_MAXM];
        if (p.factor_mode == 0) tri_matvec_t(Rm, zs, d, lane, acc);
        else gen_matvec_n(Rm, zs, d, lane, acc);
        const double sc = (phase == 0) ? 1.0 : 1.0 / c.drscale;
#pragma unroll
        for (int m = 0; m < K2_MAXM; m++) {
          const int j = lane + 32 * m;
          if (j < d) prop[j] = th[j] + (phase == 0 ? acc[m] : acc[m] * sc);
        }
        __syncwarp();
        inb = M::checkbounds(prop, d, ctx);
        if (c.method == MCMCB_ER && inb) {  // MCMC_sscrit, MCMC_DRAM.F90:124-135 (every lane keeps the same stream)
          const double u = g.uniform();
          double sum = 0.0;
#pragma unroll
          for (int k = 0; k < NY; k++) sum += ss1[k] / s2[k];
          sscrit = -2.0 * log(u) + sum + pri1;
        }
      }
      // ---------------- user model (cooperative over the 32 lanes)
      double ssn[NY];
      M::ssfunction(prop, d, NY, ctx, ssn);
#pragma unroll
      for (int k = 0; k < NY; k++) ssn[k] = warp_sum(ssn[k]);
      const double prn = M::priorfun(prop, d, ctx);
      // ---------------- accept / reject (warp uniform)
      bool reject = false;
      if (phase < 0) {
#pragma unroll
        for (int k = 0; k < NY; k++) ss1[k] = ssn[k];
        pri1 = prn;
        chainind = 1; simuind = 1; cnt = 1; pend = 1;
        if (stored) {
          for (int k = lane; k < d; k += 32) srow[k] = th[k];
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NY; k++) { srow[d + k] = ss1[k]; if (c.updatesigma) ss2st[k] = s2[k]; }
          }
        }
        phase = 0;
        continue;
      }
      if (c.method == MCMCB_ER) {  // MCMC_run_er.F90:50-83
        if (!inb) {
          bnd++;
          reject = true;
        } else if (prn >= sscrit) {
          erst++;
          reject = true;
        } else {
          const double crit = s2[0] * (sscrit - prn);
          double sum = 0.0;
#pragma unroll
          for (int k = 0; k < NY; k++) sum += ssn[k];
          reject = sum >= crit;
        }
      } else if (phase == 0) {
        if (!inb) {
          if (!c.dodr || c.method == MCMCB_RAM) bnd++;
#pragma unroll
          for (int k = 0; k < NY; k++) ssn[k] = DBL_HUGE;
          a12 = (c.method != MCMCB_RAM) ? 0.0 : rama;
          reject = true;
        } else {
          double sum = 0.0;
#pragma unroll
          for (int k = 0; k < NY; k++) sum += (ssn[k] - ss1[k]) / s2[k];
          a12 = alpha_from_tst(-0.5 * (sum + (prn - pri1)));
          reject = mh_reject(a12, g);
        }
        rama = a12;
        if (reject && c.dodr) {
          drtry++;
#pragma unroll
          for (int k = 0; k < NY; k++) ss2[k] = ssn[k];
          pri2 = inb ? prn : DBL_HUGE;
          double t = 0.0;
          for (int k = lane; k < d; k += 32) t = fma(z1[k], z1[k], t);
          z1sq = warp_sum(t);
          phase = 1;
          continue;
        }
      } else {
        if (!inb) {
          bnd++;
          reject = true;
        } else {
          double a32;
          if (a12 == 0.0) {
            a32 = 0.0;
          } else {
            double sum = 0.0;
#pragma unroll
            for (int k = 0; k < NY; k++) sum += (ss2[k] - ssn[k]) / s2[k];
            a32 = fmin(1.0, exp_subnormal_safe(-0.5 * (sum + (pri2 - prn))));
          }
          double sum = 0.0;
#pragma unroll
          for (int k = 0; k < NY; k++) sum += (ssn[k] - ss1[k]) / s2[k];
          const double l2 = -0.5 * (sum + (prn - pri1));
          double t = 0.0;
          const double sc = 1.0 / c.drscale;
          for (int k = lane; k < d; k += 32) { const double v = z2[k] * sc - z1[k]; t = fma(v, v, t); }
          const double q1 = -0.5 * (warp_sum(t) - z1sq);
          double a13 = exp_subnormal_safe(l2 + q1) * (1.0 - a32) / (1.0 - a12);
          if (a13 == a13) a13 = fmin(1.0, a13);
          reject = mh_reject(a13, g);
          if (!reject) dracc++;
        }
        phase = 0;
      }
      // ---------------- end of step
      const int i = simuind + 1;
      simuind = i;
      const bool absorbing = (c.doadapt && c.method != MCMCB_RAM && !(c.adaptend > 0 && i > c.adaptend)) ||
                             (c.greedy && c.doburnin && c.method != MCMCB_RAM && i <= c.burnintime);
      if (reject) {
        stayed++;
        cnt++; pend++;
      } else {
        if (absorbing) {  // log the completed row and its not-yet-counted weight for the adaptation kernel
          if (nbuf < p.rowcap) {
            for (int k = lane; k < d; k += 32) rb[(size_t)nbuf * (d + 1) + k] = th[k];
            // AP windows weigh a row by its repeat count, the streaming recursion by what it has not yet counted
            if (lane == 0) rb[(size_t)nbuf * (d + 1) + d] = (double)((c.doadapt && c.adapthist > 1) ? cnt : pend);
            nbuf++;
          } else {
            status |= MCMCB_ST_STORE_FULL;
          }
        }
        if (stored && chainind - 1 < p.store_rows && lane == 0) scnt[chainind - 1] = (double)cnt;
        for (int k = lane; k < d; k += 32) th[k] = prop[k];
#pragma unroll
        for (int k = 0; k < NY; k++) ss1[k] = ssn[k];
        pri1 = prn;
        chainind++;
        cnt = 1; pend = 1;
        __syncwarp();
      }
      if (c.updatesigma) {
#pragma unroll
        for (int k = 0; k < NY; k++) {
          const double gg = g.gamma(c.N0 / 2.0 + (double)p.nobs[k] / 2.0, 2.0 / (c.N0 * c.S02 + ss1[k]));
          s2[k] = 1.0 / gg;
        }
      }
      if (stored) {
        if (!reject) {
          if (chainind - 1 < p.store_rows) {
            for (int k = lane; k < d; k += 32) srow[(size_t)(chainind - 1) * (d + NY) + k] = th[k];
            if (lane == 0) {
#pragma unroll
              for (int k = 0; k < NY; k++) srow[(size_t)(chainind - 1) * (d + NY) + d + k] = ss1[k];
            }
          } else {
            status |= MCMCB_ST_STORE_FULL;
          }
        }
        if (c.updatesigma && i - 1 < p.store_rows && lane == 0) {
#pragma unroll
          for (int k = 0; k < NY; k++) ss2st[(size_t)(i - 1) * NY + k] = s2[k];
        }
      }
      if (c.method == MCMCB_RAM && c.doadapt && !(i < c.burnintime && c.doburnin)) {  // MCMC_run_ram.F90:104-179
        const double a = 1.0 / pow((double)(float)i, c.nuparam) * (rama - c.alphatarget);
        double t = 0.0;
        for (int k = lane; k < d; k += 32) t = fma(z1[k], z1[k], t);
        // the reference sums u**2 sequentially; a lane-strided tree sum differs by rounding only
        const double su2 = warp_sum(t);
        if (a >= 0.0) {
          for (int k = lane; k < d; k += 32) w1[k] = z1[k] / su2 * a;
          __syncwarp();
          warp_chud(Rm, w1, d, lane);
        } else {
          for (int k = lane; k < d; k += 32) w1[k] = -z1[k] / su2 * a;
          __syncwarp();
          if (!warp_chdd(Rm, w1, w2, z2, d, lane)) status |= MCMCB_ST_DOWNDATE_FAIL;
        }
        __threadfence_block();
        __syncwarp();
      }
      if (g.exhausted) status |= MCMCB_ST_RNG_EXHAUSTED;
      done++;
    }

    // ---- write state back
    if (p.r_resident && c.method == MCMCB_RAM && c.doadapt) {
      __syncwarp();
      for (int k = lane; k < d * d; k += 32) Rg[k] = Rs[k];
    }
    for (int k = lane; k < dp; k += 32) gth[k] = th[k];
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NY; k++) { st[(Lo.ss + k) * p.pitch] = ss1[k]; st[(Lo.s2 + k) * p.pitch] = s2[k]; }
      st[Lo.pri * p.pitch] = pri1; st[Lo.rama * p.pitch] = rama; st[Lo.spare * p.pitch] = g.spare;
      ist[Lo.i_stayed * p.pitch] = stayed; ist[Lo.i_bnd * p.pitch] = bnd; ist[Lo.i_dracc * p.pitch] = dracc;
      ist[Lo.i_drtry * p.pitch] = drtry; ist[Lo.i_chainind * p.pitch] = chainind;
      ist[Lo.i_simuind * p.pitch] = simuind; ist[Lo.i_status * p.pitch] = status;
      ist[Lo.i_hasspare * p.pitch] = g.has_spare ? 1 : 0;
      ist[Lo.i_cnt * p.pitch] = cnt; ist[Lo.i_pend * p.pitch] = pend; ist[Lo.i_nbuf * p.pitch] = nbuf;
      ist[Lo.i_er * p.pitch] = erst;
      ist[Lo.i_ndlo * p.pitch] = (int)(unsigned)(g.nd & 0xffffffffull);
      ist[Lo.i_ndhi * p.pitch] = (int)(unsigned)(g.nd >> 32);
    }
    __syncwarp();
  }
}

}  // namespace mcmcb
