// Convergence diagnostics across chains (an extension -- the reference leaves R-hat / ESS to
// mcmcstat-style post-processing of one chain).  Every diag_stride steps the current theta of every chain
// is folded into per-(chain, component) running sums about the chain's first snapshot o:
//   S = sum y,  Q = sum y^2,  P_t = sum_i y_i y_{i+t}  (t = 1..K),  y = theta - o,
// plus the first and the last K values (head / ring), which turn P_t into the exact lag-t
// autocovariance about the chain mean.  Reduction over chains happens in two phases like the pooled
// adaptation (means first, then centred sums), each followed by one sum-allreduce.
#pragma once
#include "pool.cuh"

namespace mcmcb {

struct DiagParams {
  const double* theta;        // component k of chain c at theta[c * chain_stride + k * comp_stride]
  long long chain_stride, comp_stride;
  long long nchains;
  int d, K;
  long long nsnap;            // snapshots folded so far (before this one, for the update kernel)
  double* ds;                 // [field][nchains * d], fields: o, S, Q, P[K], ring[K], head[K]
};

__device__ __forceinline__ size_t diag_f(const DiagParams& p, int field, long long e) {
  return (size_t)field * (size_t)(p.nchains * p.d) + (size_t)e;
}

// one thread per (chain, component); e = c * d + k
static __global__ void diag_update_kernel(DiagParams p) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.nchains * p.d) return;
  const long long c = e / p.d;
  const int k = (int)(e - c * p.d);
  const double x = p.theta[c * p.chain_stride + k * p.comp_stride];
  const long long n = p.nsnap;
  const int K = p.K;
  double o;
  if (n == 0) {
    o = x;
    p.ds[diag_f(p, 0, e)] = o;
    p.ds[diag_f(p, 1, e)] = 0.0;
    p.ds[diag_f(p, 2, e)] = 0.0;
    for (int t = 0; t < K; t++) p.ds[diag_f(p, 3 + t, e)] = 0.0;
  } else {
    o = p.ds[diag_f(p, 0, e)];
  }
  const double y = x - o;
  p.ds[diag_f(p, 1, e)] += y;
  p.ds[diag_f(p, 2, e)] = fma(y, y, p.ds[diag_f(p, 2, e)]);
  for (int t = 1; t <= K && t <= n; t++) {
    const double prev = p.ds[diag_f(p, 3 + K + (int)((n - t) % K), e)];
    p.ds[diag_f(p, 3 + t - 1, e)] = fma(y, prev, p.ds[diag_f(p, 3 + t - 1, e)]);
  }
  if (K > 0) {
    if (n < K) p.ds[diag_f(p, 3 + 2 * K + (int)n, e)] = y;
    p.ds[diag_f(p, 3 + K + (int)(n % K), e)] = y;
  }
}

// grid (B, d): component k = blockIdx.y.  phase 1: partial = [count, sum_c mean_c];
// phase 2 (mu_k = buf[1 + k] / buf[0]): [sum_c (mean_c - mu)^2, sum_c s2_c, sum_c acov_{t,c} (t = 1..K)].
// partial layout: [k][b][nv]
static __global__ void __launch_bounds__(POOL_THREADS) diag_reduce_kernel(DiagParams p, int phase, const double* buf,
                                                                  double* partial) {
  __shared__ double red[POOL_THREADS / 32];
  const int k = blockIdx.y, K = p.K;
  const long long n = p.nsnap;
  const int nv = (phase == 1) ? 2 : 2 + K;
  double acc[2 + MCMCB_DIAG_MAXLAGS];
  for (int v = 0; v < nv; v++) acc[v] = 0.0;
  const double mu = (phase == 2) ? buf[1 + k] / buf[0] : 0.0;
  const double dn = (double)n;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < p.nchains; c += (long long)gridDim.x * blockDim.x) {
    const long long e = c * p.d + k;
    const double o = p.ds[diag_f(p, 0, e)], S = p.ds[diag_f(p, 1, e)], Q = p.ds[diag_f(p, 2, e)];
    const double m = S / dn;
    if (phase == 1) {
      acc[0] += 1.0;
      acc[1] += o + m;
      continue;
    }
    const double dm = (o + m) - mu;
    acc[0] = fma(dm, dm, acc[0]);
    acc[1] += (n > 1) ? (Q - dn * m * m) / (dn - 1.0) : 0.0;
    // acov_t = (1/n) sum_{i<=n-t} (y_i - m)(y_{i+t} - m)
    //        = (1/n) [ P_t - m (A_t + B_t) + (n - t) m^2 ],  A_t = S - (last t values),  B_t = S - (first t values)
    double tail = 0.0, head = 0.0;
    for (int t = 1; t <= K; t++) {
      if (t >= n) break;
      tail += p.ds[diag_f(p, 3 + K + (int)((n - t) % K), e)];
      head += p.ds[diag_f(p, 3 + 2 * K + (t - 1), e)];
      const double Pt = p.ds[diag_f(p, 3 + t - 1, e)];
      acc[1 + t] += (Pt - m * ((S - tail) + (S - head)) + (double)(n - t) * m * m) / dn;
    }
  }
  for (int v = 0; v < nv; v++) {
    const double s = pool_block_sum(acc[v], red);
    if (threadIdx.x == 0) partial[((size_t)k * gridDim.x + blockIdx.x) * nv + v] = s;
  }
}

// out layout phase 1: [0] = chains, [1 + k] = sum of chain means;  phase 2: [k * (2 + K) + v]
static __global__ void diag_final_kernel(const double* partial, int nblocks, int nv, int d, int phase, double* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d * nv) return;
  const int k = idx / nv, v = idx - k * nv;
  double s = 0.0;
  for (int b = 0; b < nblocks; b++) s += partial[((size_t)k * nblocks + b) * nv + v];
  if (phase == 1) {
    if (v == 0) { if (k == 0) out[0] = s; }
    else out[1 + k] = s;
  } else {
    out[(size_t)k * nv + v] = s;
  }
}

}  // namespace mcmcb
