// K1: small-npar register-resident sampler.
//
// One chain is owned by a group of L lanes (L = 1..32, power of two).  Every lane of the
// group carries the chain's whole state in registers (theta, ss, sigma2, packed upper
// factors R/R2/iC, streaming mean/covariance, counters, RNG position); the lanes only
// split the data loop of ssfunction and add their partial sums with warp shuffles.  The
// user-model blob is staged ONCE per CTA into shared memory by TMA (cp.async.bulk) and
// reused by every chain the CTA processes.  Chain state lives in HBM as structure of
// arrays (field-major, chain contiguous) and is touched only at launch begin / end: a
// launch advances every chain by `nsteps` iterations of MCMC_LOOP.
//
// Lanes run a per-chain state machine whose only warp-converged section is the model
// evaluation: phase -1 = initial point (MCMC_run.F90:27-36), 0 = first-stage proposal,
// 1 = delayed-rejection second stage (MCMC_run.F90:65-91).  A lane whose first stage was
// accepted starts its next step while its neighbour evaluates a DR proposal, so no lane
// idles in the hot loop because of DR divergence.
//
// Reference map: loop MCMC_run.F90:41-107 / MCMC_run_ram.F90:45-80; primitives
// MCMC_DRAM.F90; adaptation MCMC_adapt.F90:12-174,181-230 in streaming form (see
// absorb()); covmat recursion matutils.F90:283-310; dpotf2/dpotri order as unblocked
// LAPACK; dchud.f:122-139, dchdd.f:141-179.
#pragma once
#include "common.cuh"

namespace mcmcb {

#ifndef MCMCB_K1_THREADS
#define MCMCB_K1_THREADS 512
#endif
constexpr int K1_THREADS = MCMCB_K1_THREADS;

// field offsets of the SoA state (doubles and ints), shared by host and device
struct K1Layout {
  int th, ss, pri, s2, r, r2, ic, cm, mean, wsum, spare, rama, gcm, gmean, gw, nf;
  int i_stayed, i_bnd, i_dracc, i_drtry, i_chainind, i_simuind, i_status, i_hasspare, i_cnt, i_pend, i_ndlo, i_ndhi,
      i_er, i_nf;
};
__host__ __device__ constexpr K1Layout k1_layout(int D, int NY) {
  K1Layout l{};
  int T = D * (D + 1) / 2;
  int o = 0;
  l.th = o; o += D;
  l.ss = o; o += NY;
  l.pri = o; o += 1;
  l.s2 = o; o += NY;
  l.r = o; o += T;
  l.r2 = o; o += T;
  l.ic = o; o += T;
  l.cm = o; o += T;
  l.mean = o; o += D;
  l.wsum = o; o += 1;
  l.spare = o; o += 1;
  l.rama = o; o += 1;
  l.gcm = o; o += T;      // greedy burn-in: unit-weight covariance of every row so far (MCMC_adapt.F90:83-101)
  l.gmean = o; o += D;
  l.gw = o; o += 1;
  l.nf = o;
  int k = 0;
  l.i_stayed = k++; l.i_bnd = k++; l.i_dracc = k++; l.i_drtry = k++; l.i_chainind = k++; l.i_simuind = k++;
  l.i_status = k++; l.i_hasspare = k++; l.i_cnt = k++; l.i_pend = k++; l.i_ndlo = k++; l.i_ndhi = k++;
  l.i_er = k++;  // erstayed, mcmc.F90:49
  l.i_nf = k;
  return l;
}

struct K1Params {
  DevCfg c;
  long long nchains, pitch, chain_offset;
  unsigned long long seed;
  int nsteps;
  double* st;  // [field][pitch]
  int* ist;    // [field][pitch]
  const double* par0;       // [nchains][D]
  const double* cmat0;      // packed upper, T doubles
  const double* sigma2_0;   // NY
  const int* nobs;          // NY
  const double* blob;       // global copy of the model blob
  unsigned long long blob_n;  // doubles
  unsigned blob_bytes;      // padded to 16
  const double* prior;      // mu[D], sig[D] or nullptr
  const double* inj;        // [nchains][inj_per_chain] or nullptr
  unsigned long long inj_per_chain;
  // stored run-length chains (first store_chains chains)
  int store_chains, store_rows;
  double* store_rows_p;     // [chain][row][D+NY]
  double* store_cnt_p;      // [chain][row]
  double* store_s2_p;       // [chain][step][NY]
  unsigned int* tile_counter;
  // AP window (adapthist > 1, MCMC_adapt.F90:116-136): ring of the last hist_rows closed run-length rows,
  // [slot][D+1][pitch] (theta then repeat count), slot = (row index - 1) mod hist_rows
  double* hist;
  int hist_rows;
  long long tier[3];  // tiles of 4, 2 and 1 sub-tiles, in this order along the chain range (k1_step_kernel)
  // bulk of a large population: `wtiles` SUPER-tiles, each `rounds` 4-sub-tile tiles run back to back by one warp
  // (super-tile s = tiles s, s + wtiles, ...); they come first along the chain range, the tiers above follow them
  long long wtiles;
  int rounds;
  double exp_c1, exp_c2;    // MCMCB_EXP_C1L / C2L: see mcmcb_expmul_fast for why they travel as parameters
  int exp_dn;               // entries of the direct exp table staged behind the blob (0 = none), mcmcb_expmul_direct
};

__host__ __device__ constexpr int pk(int i, int j) { return j * (j + 1) / 2 + i; }  // i <= j

// dpotf2('U') on a packed upper triangle, in place; false if a pivot is not positive (A is then garbage)
template <int D>
__device__ __forceinline__ bool chol_packed(double (&A)[D * (D + 1) / 2]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < D; j++) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < j; i++) t += A[pk(i, j)] * A[pk(i, j)];
    double ajj = A[pk(j, j)] - t;
    if (!(ajj > 0.0)) ok = false;
    ajj = sqrt(ajj);
    A[pk(j, j)] = ajj;
    double rajj = 1.0 / ajj;
#pragma unroll
    for (int k = j + 1; k < D; k++) {
      double tt = 0.0;
#pragma unroll
      for (int i = 0; i < j; i++) tt += A[pk(i, k)] * A[pk(i, j)];
      A[pk(j, k)] = (A[pk(j, k)] - tt) * rajj;
    }
  }
  return ok;
}

// R = chol(cm)*2.4/sqrt(D); iC = inv(R'R); R2 = R/drscale.  MCMC_adapt.F90:181-230 with
// covtor (matutils.F90:345-374) = dpotf2('U') and dpotri('U') = dtrti2 + dlauu2.
// Returns false (and leaves R,R2,iC untouched) if the factorisation fails.
template <int D>
__device__ __forceinline__ bool calculate_R(const double (&cm)[D * (D + 1) / 2], double (&R)[D * (D + 1) / 2],
                                            double (&R2)[D * (D + 1) / 2], double (&iC)[D * (D + 1) / 2],
                                            const DevCfg& c) {
  constexpr int T = D * (D + 1) / 2;
  double A[T];
#pragma unroll
  for (int k = 0; k < T; k++) A[k] = cm[k];
  const bool ok = chol_packed<D>(A);
  if (!ok) return false;
  const double sq = sqrt((double)D);
#pragma unroll
  for (int k = 0; k < T; k++) R[k] = A[k] * 2.4 / sq;
  if (c.dodr) {
#pragma unroll
    for (int k = 0; k < T; k++) A[k] = R[k];
    // dtrti2 (upper, non-unit)
#pragma unroll
    for (int j = 0; j < D; j++) {
      A[pk(j, j)] = 1.0 / A[pk(j, j)];
      double ajj = -A[pk(j, j)];
#pragma unroll
      for (int jj = 0; jj < j; jj++) {  // dtrmv('U','N','N') on the leading j x j block
        double temp = A[pk(jj, j)];
#pragma unroll
        for (int i = 0; i < jj; i++) A[pk(i, j)] += temp * A[pk(i, jj)];
        A[pk(jj, j)] = temp * A[pk(jj, jj)];
      }
#pragma unroll
      for (int i = 0; i < j; i++) A[pk(i, j)] *= ajj;
    }
    // dlauu2 (upper): A <- U * U'
#pragma unroll
    for (int i = 0; i < D; i++) {
      double aii = A[pk(i, i)];
      if (i < D - 1) {
        double t = 0.0;
#pragma unroll
        for (int k = i; k < D; k++) t += A[pk(i, k)] * A[pk(i, k)];
        A[pk(i, i)] = t;
#pragma unroll
        for (int r = 0; r < i; r++) A[pk(r, i)] *= aii;
#pragma unroll
        for (int k = i + 1; k < D; k++) {
          double temp = A[pk(i, k)];
#pragma unroll
          for (int r = 0; r < i; r++) A[pk(r, i)] += temp * A[pk(r, k)];
        }
      } else {
#pragma unroll
        for (int r = 0; r <= i; r++) A[pk(r, i)] *= aii;
      }
    }
#pragma unroll
    for (int k = 0; k < T; k++) {
      iC[k] = A[k];
      R2[k] = R[k] / c.drscale;
    }
  }
  return true;
}

// v' * S * v for symmetric S stored as packed upper (dsymv('u') + dot, MCMC_DRAM.F90:182-183)
template <int D>
__device__ __forceinline__ double quadform(const double (&S)[D * (D + 1) / 2], const double (&v)[D]) {
  double q = 0.0;
#pragma unroll
  for (int a = 0; a < D; a++) {
    double w = 0.0;
#pragma unroll
    for (int b = 0; b < D; b++) w += S[a <= b ? pk(a, b) : pk(b, a)] * v[b];
    q += w * v[a];
  }
  return q;
}

// One row of the covmat recursion (matutils.F90:283-310) applied to the streaming
// accumulators.  The reference walks stored rows lastind..chainind at every adaptation
// (MCMC_adapt.F90:141-157); feeding each row when it completes, and the partial current
// row at the adaptation tick, performs the same updates in the same order.  With
// wsum == 0 (initcmatn = 0) the reference uses the two-pass batch formula over the
// window (matutils.F90:312-337); starting the recursion from the first row (mean = x,
// cov = 0, wsum = w) is algebraically identical and differs only by rounding.
template <int D>
__device__ __forceinline__ void absorb(const double (&x)[D], double w, double (&cm)[D * (D + 1) / 2],
                                       double (&mean)[D], double& wsum) {
  if (wsum > 0.0) {
    double d[D];
#pragma unroll
    for (int k = 0; k < D; k++) d[k] = x[k] - mean[k];
    double f1 = w / (wsum + w - 1.0);
    double f2 = wsum / (wsum + w);
#pragma unroll
    for (int b = 0; b < D; b++)
#pragma unroll
      for (int a = 0; a <= b; a++) cm[pk(a, b)] = cm[pk(a, b)] + f1 * (f2 * (d[a] * d[b]) - cm[pk(a, b)]);
    double f3 = w / (wsum + w);
#pragma unroll
    for (int k = 0; k < D; k++) mean[k] = mean[k] + f3 * d[k];
    wsum = w + wsum;
  } else if (w > 0.0) {
#pragma unroll
    for (int k = 0; k < D; k++) mean[k] = x[k];
#pragma unroll
    for (int k = 0; k < D * (D + 1) / 2; k++) cm[k] = 0.0;
    wsum = w;
  }
}

// dchud.f:122-139 on packed upper R
template <int D>
__device__ __forceinline__ void chud(double (&R)[D * (D + 1) / 2], const double (&x)[D]) {
  double cs[D], sn[D];
#pragma unroll
  for (int j = 0; j < D; j++) {
    double xj = x[j];
#pragma unroll
    for (int i = 0; i < j; i++) {
      double t = cs[i] * R[pk(i, j)] + sn[i] * xj;
      xj = cs[i] * xj - sn[i] * R[pk(i, j)];
      R[pk(i, j)] = t;
    }
    drotg(R[pk(j, j)], xj, cs[j], sn[j]);
  }
}

// dchdd.f:141-179 on packed upper R; returns false (R untouched) when not positive definite
template <int D>
__device__ __forceinline__ bool chdd(double (&R)[D * (D + 1) / 2], const double (&x)[D]) {
  double s[D], c[D];
  s[0] = x[0] / R[pk(0, 0)];
#pragma unroll
  for (int j = 1; j < D; j++) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < j; i++) t += R[pk(i, j)] * s[i];
    s[j] = (x[j] - t) / R[pk(j, j)];
  }
  // classic dnrm2
  double norm;
  if (D == 1) {
    norm = fabs(s[0]);
  } else {
    double scale = 0.0, ssq = 1.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
      if (s[i] != 0.0) {
        double a = fabs(s[i]);
        if (scale < a) {
          double t = scale / a;
          ssq = 1.0 + ssq * t * t;
          scale = a;
        } else {
          double t = a / scale;
          ssq = ssq + t * t;
        }
      }
    }
    norm = scale * sqrt(ssq);
  }
  if (!(norm < 1.0)) return false;
  double alpha = sqrt(1.0 - norm * norm);
#pragma unroll
  for (int i = D - 1; i >= 0; i--) {
    double scale = alpha + fabs(s[i]);
    double a = alpha / scale, b = s[i] / scale;
    double nr = sqrt(a * a + b * b);
    c[i] = a / nr;
    s[i] = b / nr;
    alpha = scale * nr;
  }
#pragma unroll
  for (int j = 0; j < D; j++) {
    double xx = 0.0;
#pragma unroll
    for (int i = j; i >= 0; i--) {
      double t = c[i] * xx + s[i] * R[pk(i, j)];
      R[pk(i, j)] = c[i] * R[pk(i, j)] - s[i] * xx;
      xx = t;
    }
  }
  return true;
}

// Initial state: what MCMC_init leaves behind (MCMC_init.F90:99-116,147-154)
template <class M>
__global__ void k1_init_kernel(K1Params p) {
  constexpr int D = M::NPAR, NY = M::NY, T = D * (D + 1) / 2;
  constexpr K1Layout Lo = k1_layout(D, NY);
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.nchains) return;
  double cm[T], R[T], R2[T], iC[T];
#pragma unroll
  for (int k = 0; k < T; k++) {
    cm[k] = p.cmat0[k];
    R[k] = 0.0; R2[k] = 0.0; iC[k] = 0.0;
  }
  bool ok = calculate_R<D>(cm, R, R2, iC, p.c);
  double* st = p.st + c;
  int* ist = p.ist + c;
#pragma unroll
  for (int k = 0; k < D; k++) {
    double v = p.par0[c * D + k];
    st[(Lo.th + k) * p.pitch] = v;
    st[(Lo.mean + k) * p.pitch] = v;
  }
#pragma unroll
  for (int k = 0; k < NY; k++) {
    st[(Lo.ss + k) * p.pitch] = 0.0;
    st[(Lo.s2 + k) * p.pitch] = p.sigma2_0[k];
  }
  st[Lo.pri * p.pitch] = 0.0;
#pragma unroll
  for (int k = 0; k < T; k++) {
    st[(Lo.r + k) * p.pitch] = R[k];
    st[(Lo.r2 + k) * p.pitch] = R2[k];
    st[(Lo.ic + k) * p.pitch] = iC[k];
    st[(Lo.cm + k) * p.pitch] = cm[k];
  }
  st[Lo.wsum * p.pitch] = (double)p.c.initcmatn;
  st[Lo.spare * p.pitch] = 0.0;
  st[Lo.rama * p.pitch] = 0.0;
#pragma unroll
  for (int k = 0; k < T; k++) st[(Lo.gcm + k) * p.pitch] = cm[k];
#pragma unroll
  for (int k = 0; k < D; k++) st[(Lo.gmean + k) * p.pitch] = p.par0[c * D + k];
  st[Lo.gw * p.pitch] = (double)p.c.initcmatn;
#pragma unroll
  for (int k = 0; k < Lo.i_nf; k++) ist[k * p.pitch] = 0;
  if (!ok) ist[Lo.i_status * p.pitch] = MCMCB_ST_CHOLFAIL;
}

// Everything one chain owns.  The struct is handed by pointer to the two non-inlined cold
// routines below, so it lives in (L1-cached, lane-interleaved) local memory; the hot model
// loop only keeps the proposal and its accumulators in registers.  That caps the kernel at
// K1_REGS registers per thread and buys the occupancy the FP64 pipe needs to stay busy.
template <int D, int NY>
struct K1State {
  static constexpr int T = D * (D + 1) / 2;
  double th[D], ss1[NY], s2[NY], R[T], R2[T], iC[T], cm[T], mean[D];
  double pri1, wsum, rama;
  double gcm[T], gmean[D], gw;              // greedy burn-in accumulators (unit row weights)
  double y1[D], z1[D], ss2[NY], pri2, a12;  // first-stage proposal kept for DR / RAM
  double sscrit;                            // method 'er': critical value of this proposal (MCMC_DRAM.F90:124-135)
  int stayed, bnd, dracc, drtry, chainind, simuind, status, cnt, pend, er;
  int phase, done;
  bool valid, stored;
  long long cc;
  Rng g;
};

template <class M>
__device__ __forceinline__ void k1_load_state(K1State<M::NPAR, M::NY>& S, const K1Params& p, long long cc) {
  constexpr int D = M::NPAR, NY = M::NY, T = D * (D + 1) / 2;
  constexpr K1Layout Lo = k1_layout(D, NY);
  const double* st = p.st + cc;
  const int* ist = p.ist + cc;
#pragma unroll
  for (int k = 0; k < D; k++) { S.th[k] = st[(Lo.th + k) * p.pitch]; S.mean[k] = st[(Lo.mean + k) * p.pitch]; }
#pragma unroll
  for (int k = 0; k < NY; k++) { S.ss1[k] = st[(Lo.ss + k) * p.pitch]; S.s2[k] = st[(Lo.s2 + k) * p.pitch]; }
#pragma unroll
  for (int k = 0; k < T; k++) {
    S.R[k] = st[(Lo.r + k) * p.pitch]; S.R2[k] = st[(Lo.r2 + k) * p.pitch];
    S.iC[k] = st[(Lo.ic + k) * p.pitch]; S.cm[k] = st[(Lo.cm + k) * p.pitch];
  }
  S.pri1 = st[Lo.pri * p.pitch]; S.wsum = st[Lo.wsum * p.pitch]; S.rama = st[Lo.rama * p.pitch];
  if (p.c.greedy) {
#pragma unroll
    for (int k = 0; k < T; k++) S.gcm[k] = st[(Lo.gcm + k) * p.pitch];
#pragma unroll
    for (int k = 0; k < D; k++) S.gmean[k] = st[(Lo.gmean + k) * p.pitch];
    S.gw = st[Lo.gw * p.pitch];
  }
  S.er = ist[Lo.i_er * p.pitch]; S.sscrit = 0.0;
  S.stayed = ist[Lo.i_stayed * p.pitch]; S.bnd = ist[Lo.i_bnd * p.pitch]; S.dracc = ist[Lo.i_dracc * p.pitch];
  S.drtry = ist[Lo.i_drtry * p.pitch]; S.chainind = ist[Lo.i_chainind * p.pitch];
  S.simuind = ist[Lo.i_simuind * p.pitch]; S.status = ist[Lo.i_status * p.pitch];
  S.cnt = ist[Lo.i_cnt * p.pitch]; S.pend = ist[Lo.i_pend * p.pitch];
  Rng& g = S.g;
  g.nd = ((unsigned long long)(unsigned)ist[Lo.i_ndhi * p.pitch] << 32) | (unsigned)ist[Lo.i_ndlo * p.pitch];
  g.seed = p.seed; g.chain = (unsigned long long)(p.chain_offset + cc);
  g.inj = p.inj ? p.inj + (unsigned long long)cc * p.inj_per_chain : nullptr;
  g.inj_n = p.inj_per_chain;
  g.cache_valid = false; g.cache_lo = g.cache_hi = 0; g.cache_blk = 0;
  g.has_spare = ist[Lo.i_hasspare * p.pitch] != 0;
  g.spare = st[Lo.spare * p.pitch];
  g.exhausted = 0;
  S.phase = (S.simuind == 0) ? -1 : 0;
  S.done = 0;
  S.pri2 = 0.0; S.a12 = 0.0;
#pragma unroll
  for (int k = 0; k < D; k++) { S.y1[k] = S.th[k]; S.z1[k] = 0.0; }
#pragma unroll
  for (int k = 0; k < NY; k++) S.ss2[k] = 0.0;
  S.cc = cc;
}

template <class M>
__device__ __forceinline__ void k1_store_state(const K1State<M::NPAR, M::NY>& S, const K1Params& p) {
  constexpr int D = M::NPAR, NY = M::NY, T = D * (D + 1) / 2;
  constexpr K1Layout Lo = k1_layout(D, NY);
  double* st = p.st + S.cc;
  int* ist = p.ist + S.cc;
#pragma unroll
  for (int k = 0; k < D; k++) { st[(Lo.th + k) * p.pitch] = S.th[k]; st[(Lo.mean + k) * p.pitch] = S.mean[k]; }
#pragma unroll
  for (int k = 0; k < NY; k++) { st[(Lo.ss + k) * p.pitch] = S.ss1[k]; st[(Lo.s2 + k) * p.pitch] = S.s2[k]; }
#pragma unroll
  for (int k = 0; k < T; k++) {
    st[(Lo.r + k) * p.pitch] = S.R[k]; st[(Lo.r2 + k) * p.pitch] = S.R2[k];
    st[(Lo.ic + k) * p.pitch] = S.iC[k]; st[(Lo.cm + k) * p.pitch] = S.cm[k];
  }
  st[Lo.pri * p.pitch] = S.pri1; st[Lo.wsum * p.pitch] = S.wsum; st[Lo.rama * p.pitch] = S.rama;
  st[Lo.spare * p.pitch] = S.g.spare;
  if (p.c.greedy) {
#pragma unroll
    for (int k = 0; k < T; k++) st[(Lo.gcm + k) * p.pitch] = S.gcm[k];
#pragma unroll
    for (int k = 0; k < D; k++) st[(Lo.gmean + k) * p.pitch] = S.gmean[k];
    st[Lo.gw * p.pitch] = S.gw;
  }
  ist[Lo.i_er * p.pitch] = S.er;
  ist[Lo.i_stayed * p.pitch] = S.stayed; ist[Lo.i_bnd * p.pitch] = S.bnd; ist[Lo.i_dracc * p.pitch] = S.dracc;
  ist[Lo.i_drtry * p.pitch] = S.drtry; ist[Lo.i_chainind * p.pitch] = S.chainind;
  ist[Lo.i_simuind * p.pitch] = S.simuind; ist[Lo.i_status * p.pitch] = S.status;
  ist[Lo.i_hasspare * p.pitch] = S.g.has_spare ? 1 : 0;
  ist[Lo.i_cnt * p.pitch] = S.cnt; ist[Lo.i_pend * p.pitch] = S.pend;
  ist[Lo.i_ndlo * p.pitch] = (int)(unsigned)(S.g.nd & 0xffffffffull);
  ist[Lo.i_ndhi * p.pitch] = (int)(unsigned)(S.g.nd >> 32);
}

// Next proposal of this chain (cold, divergent): theta + R'z, MCMC_DRAM.F90:20-31, with the
// first- or second-stage factor.  Returns the bounds verdict (MCMC_DRAM.F90:37-46).
template <class M>
__device__ __noinline__ bool k1_prepare(K1State<M::NPAR, M::NY>* Sp, double* prop_out, const mcmcb_ctx* ctx, int method) {
  constexpr int D = M::NPAR, NY = M::NY;
  K1State<M::NPAR, M::NY>& S = *Sp;
  double prop[D];
  if (S.phase < 0) {
#pragma unroll
    for (int k = 0; k < D; k++) prop_out[k] = S.th[k];
    return true;
  }
  double z[D];
#pragma unroll
  for (int k = 0; k < D; k++) z[k] = S.g.normal();
  const bool first = S.phase == 0;
#pragma unroll
  for (int j = D - 1; j >= 0; j--) {  // dtrmv('u','t','n') order, matutils.F90:108-109
    double acc = z[j] * (first ? S.R[pk(j, j)] : S.R2[pk(j, j)]);
#pragma unroll
    for (int i = j - 1; i >= 0; i--) acc += (first ? S.R[pk(i, j)] : S.R2[pk(i, j)]) * z[i];
    prop[j] = S.th[j] + acc;
  }
  if (first) {
#pragma unroll
    for (int k = 0; k < D; k++) S.z1[k] = z[k];
  }
#pragma unroll
  for (int k = 0; k < D; k++) prop_out[k] = prop[k];
  const bool inb = M::checkbounds(prop, D, *ctx);
  if (method == MCMCB_ER && inb) {  // MCMC_sscrit, MCMC_DRAM.F90:124-135: the uniform is drawn for every in-bounds proposal
    const double u = S.g.uniform();
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < NY; k++) sum += S.ss1[k] / S.s2[k];
    S.sscrit = -2.0 * log(u) + sum + S.pri1;
  }
  return inb;
}

// Accept / reject and, at the end of a step, everything MCMC_run.F90:93-105 does after it
// (cold, divergent).
template <class M>
__device__ __noinline__ void k1_finish(K1State<M::NPAR, M::NY>* Sp, const K1Params* pp, const double* prop_in,
                                       const double* ssn_in, double prn, bool inb) {
  constexpr int D = M::NPAR, NY = M::NY, T = D * (D + 1) / 2;
  K1State<D, NY>& S = *Sp;
  const K1Params& p = *pp;
  const DevCfg& c = p.c;
  Rng& g = S.g;
  double prop[D], ssn[NY];
#pragma unroll
  for (int k = 0; k < D; k++) prop[k] = prop_in[k];
#pragma unroll
  for (int k = 0; k < NY; k++) ssn[k] = ssn_in[k];
  double* srow = p.store_rows_p + (size_t)S.cc * p.store_rows * (D + NY);
  double* scnt = p.store_cnt_p + (size_t)S.cc * p.store_rows;
  double* ss2st = p.store_s2_p + (size_t)S.cc * p.store_rows * NY;
  const bool stored = S.stored;

  bool reject = false;
  if (S.phase < 0) {  // MCMC_run.F90:27-36: initial point, saved as row 1
#pragma unroll
    for (int k = 0; k < NY; k++) S.ss1[k] = ssn[k];
    S.pri1 = prn;
    S.chainind = 1; S.simuind = 1; S.cnt = 1; S.pend = 1;
    if (c.greedy && c.doburnin) absorb<D>(S.th, 1.0, S.gcm, S.gmean, S.gw);
    if (stored) {
#pragma unroll
      for (int k = 0; k < D; k++) srow[k] = S.th[k];
#pragma unroll
      for (int k = 0; k < NY; k++) { srow[D + k] = S.ss1[k]; if (c.updatesigma) ss2st[k] = S.s2[k]; }
    }
    S.phase = 0;
    return;
  }
  if (c.method == MCMCB_ER) {  // MCMC_run_er.F90:50-83
    if (!inb) {
      S.bnd++;
      reject = true;
    } else if (prn >= S.sscrit) {  // rejected by the prior alone
      S.er++;
      reject = true;
    } else {
      const double crit = S.s2[0] * (S.sscrit - prn);  // "problem here if nycol > 1", MCMC_run_er.F90:70-71
      double sum = 0.0;
#pragma unroll
      for (int k = 0; k < NY; k++) sum += ssn[k];
      reject = sum >= crit;
    }
  } else if (S.phase == 0) {  // MCMC_run.F90:46-59
    if (!inb) {
      if (!c.dodr || c.method == MCMCB_RAM) S.bnd++;
#pragma unroll
      for (int k = 0; k < NY; k++) ssn[k] = DBL_HUGE;
      S.a12 = (c.method != MCMCB_RAM) ? 0.0 : S.rama;  // RAM keeps the stale alpha12 (MCMC_run_ram.F90:52-55)
      reject = true;
    } else {
      double sum = 0.0;
#pragma unroll
      for (int k = 0; k < NY; k++) sum += (ssn[k] - S.ss1[k]) / S.s2[k];
      S.a12 = alpha_from_tst(-0.5 * (sum + (prn - S.pri1)));
      reject = mh_reject(S.a12, g);
    }
    S.rama = S.a12;
    if (reject && c.dodr) {  // MCMC_run.F90:65-68: one delayed-rejection try
      S.drtry++;
#pragma unroll
      for (int k = 0; k < D; k++) S.y1[k] = prop[k];
#pragma unroll
      for (int k = 0; k < NY; k++) S.ss2[k] = ssn[k];
      S.pri2 = inb ? prn : DBL_HUGE;
      S.phase = 1;
      return;
    }
  } else {  // MCMC_run.F90:69-91
    if (!inb) {
      S.bnd++;
      reject = true;
    } else {  // MCMC_DR_alpha13, MCMC_DRAM.F90:162-186
      double a32;
      if (S.a12 == 0.0) {
        a32 = 0.0;
      } else {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < NY; k++) sum += (S.ss2[k] - ssn[k]) / S.s2[k];
        a32 = fmin(1.0, exp_subnormal_safe(-0.5 * (sum + (S.pri2 - prn))));
      }
      double sum = 0.0;
#pragma unroll
      for (int k = 0; k < NY; k++) sum += (ssn[k] - S.ss1[k]) / S.s2[k];
      double l2 = -0.5 * (sum + (prn - S.pri1));
      double va[D], vb[D];
#pragma unroll
      for (int k = 0; k < D; k++) { va[k] = prop[k] - S.y1[k]; vb[k] = S.th[k] - S.y1[k]; }
      double q1 = -0.5 * (quadform<D>(S.iC, va) - quadform<D>(S.iC, vb));
      double a13 = exp_subnormal_safe(l2 + q1) * (1.0 - a32) / (1.0 - S.a12);
      if (a13 == a13) a13 = fmin(1.0, a13);  // NaN rejects
      reject = mh_reject(a13, g);
      if (!reject) S.dracc++;
    }
    S.phase = 0;
  }

  // ---------------- end of one MCMC_LOOP iteration, MCMC_run.F90:93-105
  const int i = S.simuind + 1;
  S.simuind = i;
  const bool absorbing = c.doadapt && c.method != MCMCB_RAM && !(c.adaptend > 0 && i > c.adaptend) && c.adapthist <= 1;
  if (reject) {
    S.stayed++;
    S.cnt++; S.pend++;
  } else {
    if (absorbing) absorb<D>(S.th, (double)S.pend, S.cm, S.mean, S.wsum);
    if (stored && S.chainind - 1 < p.store_rows) scnt[S.chainind - 1] = (double)S.cnt;
    if (p.hist != nullptr) {  // the closing row joins the AP window ring
      double* hr = p.hist + ((size_t)((S.chainind - 1) % p.hist_rows) * (D + 1)) * p.pitch + S.cc;
#pragma unroll
      for (int k = 0; k < D; k++) hr[(size_t)k * p.pitch] = S.th[k];
      hr[(size_t)D * p.pitch] = (double)S.cnt;
    }
#pragma unroll
    for (int k = 0; k < D; k++) S.th[k] = prop[k];
#pragma unroll
    for (int k = 0; k < NY; k++) S.ss1[k] = ssn[k];
    S.pri1 = prn;
    S.chainind++;
    S.cnt = 1; S.pend = 1;
    // greedy burn-in: covmat over ALL rows with unit weights (MCMC_adapt.F90:88-93) == every row fed once, when it opens
    if (c.greedy && c.doburnin && i < c.burnintime) absorb<D>(S.th, 1.0, S.gcm, S.gmean, S.gw);
  }
  if (c.updatesigma) {  // MCMC_updatesigma2, MCMC_DRAM.F90:192-206
#pragma unroll
    for (int k = 0; k < NY; k++) {
      double gg = g.gamma(c.N0 / 2.0 + (double)p.nobs[k] / 2.0, 2.0 / (c.N0 * c.S02 + S.ss1[k]));
      S.s2[k] = 1.0 / gg;
    }
  }
  if (stored) {  // MCMC_savechain, MCMC_aux.F90:166-185
    if (!reject) {
      if (S.chainind - 1 < p.store_rows) {
#pragma unroll
        for (int k = 0; k < D; k++) srow[(size_t)(S.chainind - 1) * (D + NY) + k] = S.th[k];
#pragma unroll
        for (int k = 0; k < NY; k++) srow[(size_t)(S.chainind - 1) * (D + NY) + D + k] = S.ss1[k];
      } else {
        S.status |= MCMCB_ST_STORE_FULL;
      }
    }
    if (c.updatesigma && i - 1 < p.store_rows) {
#pragma unroll
      for (int k = 0; k < NY; k++) ss2st[(size_t)(i - 1) * NY + k] = S.s2[k];
    }
  }
  if (c.method == MCMCB_RAM) {  // MCMC_adapt_ram, MCMC_run_ram.F90:104-179
    if (c.doadapt && !(i < c.burnintime && c.doburnin)) {
      double a = 1.0 / pow((double)(float)i, c.nuparam) * (S.rama - c.alphatarget);
      double su2 = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) su2 += S.z1[k] * S.z1[k];
      double xv[D];
      if (a >= 0.0) {
#pragma unroll
        for (int k = 0; k < D; k++) xv[k] = S.z1[k] / su2 * a;
        chud<D>(S.R, xv);
      } else {
#pragma unroll
        for (int k = 0; k < D; k++) xv[k] = -S.z1[k] / su2 * a;
        if (!chdd<D>(S.R, xv)) S.status |= MCMCB_ST_DOWNDATE_FAIL;
      }
    }
  } else if ((c.doadapt || c.doburnin) && !(c.adaptend > 0 && i > c.adaptend)) {  // MCMC_adapt.F90:42-46
    const int ma = c.adaptint > 0 ? i % c.adaptint : 1;
    const int mb = c.badaptint > 0 ? i % c.badaptint : 1;
    if (ma == 0 || mb == 0) {
      if (i < c.burnintime && c.doburnin && mb == 0) {  // MCMC_adapt.F90:60-102
        const double staypc = (double)S.stayed / (double)i;
        if (staypc > 1.0 - c.scalelimit) {
#pragma unroll
          for (int k = 0; k < T; k++) {
            S.R[k] = S.R[k] / c.scalefactor;
            if (c.dodr) { S.R2[k] = S.R2[k] / c.scalefactor; S.iC[k] = S.iC[k] * c.scalefactor * c.scalefactor; }
          }
        } else if (staypc < c.scalelimit) {
#pragma unroll
          for (int k = 0; k < T; k++) {
            S.R[k] = S.R[k] * c.scalefactor;
            if (c.dodr) { S.R2[k] = S.R2[k] * c.scalefactor; S.iC[k] = S.iC[k] / c.scalefactor / c.scalefactor; }
          }
        } else if (c.greedy) {
          // MCMC_adapt.F90:83-102: chaincmat = covariance of rows 1..chainind with unit weights on top of
          // (cmat0, par0, initcmatn) -- the greedy accumulators hold exactly that; lastfreq = count of the
          // open row, lastind = chainind; then MCMC_calculate_R(chaincmat)
          if (!calculate_R<D>(S.gcm, S.R, S.R2, S.iC, c)) S.status |= MCMCB_ST_CHOLFAIL;
          // What the AM branch starts from: the reference resets (chaincmat, chainmean, chainwsum) at
          // simuind == burnintime+adaptint+adapthist (MCMC_adapt.F90:108-114) -- only if that step is a tick --
          // and then feeds rows lastind..chainind; otherwise it carries on from the greedy covariance.
          const int t0 = c.burnintime + c.adaptint + c.adapthist;
          const bool will_reset = c.doadapt && ((c.adaptint > 0 && t0 % c.adaptint == 0) || (c.badaptint > 0 && t0 % c.badaptint == 0)) &&
                                  !(c.adaptend > 0 && t0 > c.adaptend);
          if (will_reset) {
#pragma unroll
            for (int k = 0; k < T; k++) S.cm[k] = p.cmat0[k];
#pragma unroll
            for (int k = 0; k < D; k++) S.mean[k] = p.par0[S.cc * D + k];
            S.wsum = (double)c.initcmatn;
          } else {
#pragma unroll
            for (int k = 0; k < T; k++) S.cm[k] = S.gcm[k];
#pragma unroll
            for (int k = 0; k < D; k++) S.mean[k] = S.gmean[k];
            S.wsum = S.gw;
          }
          S.pend = 0;  // lastfreq = current count: only later repeats of the open row count
        } else {
          // lastind = chainind (MCMC_adapt.F90:102): rows before the current one never enter the
          // covariance, the current row enters with its full count (lastfreq stays 0); then
          // MCMC_calculate_R(chaincmat) with chaincmat still == cmat0 (no AM update happened yet)
          double c0[T];
#pragma unroll
          for (int k = 0; k < T; k++) { c0[k] = p.cmat0[k]; S.cm[k] = c0[k]; }
#pragma unroll
          for (int k = 0; k < D; k++) S.mean[k] = p.par0[S.cc * D + k];
          S.wsum = (double)c.initcmatn;
          S.pend = S.cnt;
          if (!calculate_R<D>(c0, S.R, S.R2, S.iC, c)) S.status |= MCMCB_ST_CHOLFAIL;
        }
      } else if (i >= c.burnintime + c.adaptint + c.adapthist && c.doadapt) {  // MCMC_adapt.F90:105-159
        if (c.adapthist > 1) {
          // AP (MCMC_adapt.F90:116-136): covariance of the last adapthist steps, batch formula of covmat
          // (matutils.F90:312-337) over the run-length rows istart..chainind, first weight trimmed
          const int H = p.hist_rows;
          const double* hb = p.hist + S.cc;
          int istart = S.chainind, histsum = S.cnt;
          while (histsum < c.adapthist && istart > 1) {
            istart--;
            histsum += (int)hb[((size_t)((istart - 1) % H) * (D + 1) + D) * p.pitch];
          }
          const int nrow = S.chainind - istart + 1;
          double wsum2 = 0.0, xs[D], xm[D];
#pragma unroll
          for (int k = 0; k < D; k++) xs[k] = 0.0;
          for (int r = 0; r < nrow; r++) {
            const int row = istart + r;
            const double* hr = hb + ((size_t)((row - 1) % H) * (D + 1)) * p.pitch;
            double w = (row == S.chainind) ? (double)S.cnt : hr[(size_t)D * p.pitch];
            if (r == 0) w = (double)((int)w - histsum + c.adapthist);
            wsum2 += w;
#pragma unroll
            for (int k = 0; k < D; k++) xs[k] = xs[k] + ((row == S.chainind) ? S.th[k] : hr[(size_t)k * p.pitch]) * w;
          }
#pragma unroll
          for (int k = 0; k < D; k++) xm[k] = xs[k] / wsum2;
          double acc[T];
#pragma unroll
          for (int k = 0; k < T; k++) acc[k] = 0.0;
          for (int r = 0; r < nrow; r++) {
            const int row = istart + r;
            const double* hr = hb + ((size_t)((row - 1) % H) * (D + 1)) * p.pitch;
            double w = (row == S.chainind) ? (double)S.cnt : hr[(size_t)D * p.pitch];
            if (r == 0) w = (double)((int)w - histsum + c.adapthist);
            double dv[D];
#pragma unroll
            for (int k = 0; k < D; k++) dv[k] = ((row == S.chainind) ? S.th[k] : hr[(size_t)k * p.pitch]) - xm[k];
#pragma unroll
            for (int b = 0; b < D; b++)
#pragma unroll
              for (int a = 0; a <= b; a++) acc[pk(a, b)] = acc[pk(a, b)] + dv[b] * (dv[a] * w);  // cmat(i,j), j<=i: (x_i-m_i)*((x_j-m_j)*w)
          }
#pragma unroll
          for (int k = 0; k < T; k++) S.cm[k] = acc[k] / (wsum2 - 1.0);
#pragma unroll
          for (int k = 0; k < D; k++) S.mean[k] = xm[k];
          S.wsum = wsum2;
        } else {
          absorb<D>(S.th, (double)S.pend, S.cm, S.mean, S.wsum);
        }
        S.pend = 0;
        // pooled adaptation: the factor comes from the covariance pooled over all chains (pool.cuh)
        if (!c.pool && !calculate_R<D>(S.cm, S.R, S.R2, S.iC, c)) S.status |= MCMCB_ST_CHOLFAIL;
      }
    }
  }
  if (g.exhausted) S.status |= MCMCB_ST_RNG_EXHAUSTED;
  S.done++;
}

// does the model supply the early-rejection form of its ssfunction (ssfunction_er, external_inc.h:20-24)?
template <class M, class = void>
struct has_ssfunction_er { static constexpr bool value = false; };
template <class M>
struct has_ssfunction_er<M, decltype((void)&M::ssfunction_er)> { static constexpr bool value = true; };

// does the model evaluate several parameter vectors in one sweep over its data (ssfunction_batch<B>)?
template <class M, class = void>
struct has_ssfunction_batch { static constexpr bool value = false; };
template <class M>
struct has_ssfunction_batch<M, decltype((void)&M::template ssfunction_batch<2>)> { static constexpr bool value = true; };

// One tile: NB x (32/L) consecutive chains, NB of them per thread, advanced by p.nsteps iterations.
//
// EREXIT: method 'er' with a model that has ssfunction_er -- every lane hands the model its critical value and
// the warp leaves the data loop as soon as every lane's partial sum has reached it (the GPU form of the
// reference's "stop summing once ss >= sscrit").  A rejected proposal's ss is never used, an accepted one's
// loop ran to the end, so chains are identical with and without the early exit.
//
// NB > 1 (L == 1 only): every lane runs NB independent chain state machines and hands the model NB proposals per
// sweep (M::ssfunction_batch<NB>): one shared-memory read of a datum then serves NB chains.  The datum loop of the
// exponential-regression model is bound by shared-memory wavefronts (broadcast data reads + conflicting table
// lookups), so cutting the data reads per chain-datum is what this buys.  The per-chain arithmetic
// and its order are those of NB == 1: results are bit-identical.
//
// K > 1 (super-tile): every chain slot runs K chains one after the other (chain `first + k * kstride + ...`, k = 0..K-1)
// and moves on to its next chain as soon as the current one has done its nsteps steps, without waiting for the other
// slots of the warp.  Under delayed rejection a chain needs nsteps + Binomial(nsteps, q) evaluations; the warp runs
// until its slowest slot is done, and the maximum over 32 NB slots of a SUM of K such counts is relatively
// sqrt(K) closer to the mean than the maximum of single counts (the wait for the slowest chain of every tile was
// ~6 % of the launch on BASELINE C3).  Chains are independent: results do not depend on K.
template <class M, int L, bool EREXIT, int NB>
__device__ __forceinline__ void k1_run_tile(const K1Params& p, const mcmcb_ctx& ctx, long long first, int sub, int gl,
                                            int K = 1, long long kstride = 0) {
  constexpr int D = M::NPAR, NY = M::NY;
  K1State<D, NY> S[NB];
  double prop[NB][D];
  int kk[NB];
#pragma unroll
  for (int b = 0; b < NB; b++) {
    const long long ch = first + b * (32 / L) + sub;  // consecutive lanes, consecutive chains
    S[b].valid = ch < p.nchains;
    k1_load_state<M>(S[b], p, S[b].valid ? ch : p.nchains - 1);
    S[b].stored = S[b].valid && (ch < p.store_chains) && gl == 0;
    kk[b] = 0;
#pragma unroll
    for (int k = 0; k < D; k++) prop[b][k] = S[b].th[k];
  }
  for (;;) {
    bool act[NB], inb[NB], any = false;
#pragma unroll
    for (int b = 0; b < NB; b++) {
      act[b] = S[b].valid && (S[b].phase < 0 || S[b].done < p.nsteps);
      any = any || act[b];
      inb[b] = true;
    }
    if (!__any_sync(FULL, any)) break;
#pragma unroll
    for (int b = 0; b < NB; b++)
      if (act[b]) inb[b] = k1_prepare<M>(&S[b], prop[b], &ctx, p.c.method);
    __syncwarp();
    // ---------------- user model: the hot, warp-converged section
    double ssn[NB][NY];
    if constexpr (NB > 1) {
      M::template ssfunction_batch<NB>(&prop[0][0], D, NY, ctx, &ssn[0][0]);
    } else if constexpr (EREXIT) {
      double crit = -DBL_HUGE;  // lanes whose ss nobody reads vote "done" at once
      if (act[0]) {
        if (S[0].phase < 0) {
          crit = DBL_HUGE;  // initial point: the full sum is needed
        } else if (inb[0]) {
          const double pr = M::priorfun(prop[0], D, ctx);
          if (pr < S[0].sscrit) crit = S[0].s2[0] * (S[0].sscrit - pr);
        }
      }
      M::ssfunction_er(prop[0], D, NY, ctx, crit, ssn[0]);
    } else {
      M::ssfunction(prop[0], D, NY, ctx, ssn[0]);
    }
    if (L > 1) {
#pragma unroll
      for (int k = 0; k < NY; k++) {
#pragma unroll
        for (int off = L / 2; off > 0; off >>= 1) ssn[0][k] += __shfl_xor_sync(FULL, ssn[0][k], off);
      }
    }
#pragma unroll
    for (int b = 0; b < NB; b++) {
      const double prn = M::priorfun(prop[b], D, ctx);
      if (act[b]) {
        k1_finish<M>(&S[b], &p, prop[b], ssn[b], prn, inb[b]);
        if (K > 1 && S[b].phase == 0 && S[b].done >= p.nsteps && kk[b] + 1 < K) {  // this slot's next chain
          if (S[b].valid && gl == 0) k1_store_state<M>(S[b], p);
          kk[b]++;
          const long long ch = first + (long long)kk[b] * kstride + b * (32 / L) + sub;
          S[b].valid = ch < p.nchains;
          k1_load_state<M>(S[b], p, S[b].valid ? ch : p.nchains - 1);
          S[b].stored = S[b].valid && (ch < p.store_chains) && gl == 0;
#pragma unroll
          for (int k = 0; k < D; k++) prop[b][k] = S[b].th[k];
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < NB; b++)
    if (S[b].valid && gl == 0) k1_store_state<M>(S[b], p);
  __syncwarp();
}

// Persistent CTAs; every warp pulls tiles from a global counter.  The chain range is cut into three tiers of
// tiles holding 4, 2 and 1 sub-tiles (a sub-tile = the 32/L chains one warp owns with one chain per thread), laid
// out by the host (K1Params::tier) so that the bulk runs B chains per thread and the last round is made of smaller
// tiles -- with 4-sub-tile tiles alone the last round would leave most warps idle for a quarter of the launch.
template <class M, int L, bool SMEM, bool EREXIT = false, int B = 1>
__global__ void __launch_bounds__(K1_THREADS, 1) k1_step_kernel(const __grid_constant__ K1Params p) {
  static_assert(B == 1 || (L == 1 && !EREXIT), "chains-per-thread batching is for the thread-per-chain mapping");
  constexpr int SUB = 32 / L;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;

  // dynamic shared memory: [exp2 table, 16 KB][model blob (when it fits)][direct exp table, exp_dn entries]
  double* exp_tab = reinterpret_cast<double*>(smem_raw);
  mcmcb_stage_exp_table(exp_tab);
  double* exp_direct = reinterpret_cast<double*>(smem_raw + MCMCB_EXP_TAB_DOUBLES * sizeof(double) + (SMEM ? p.blob_bytes : 0u));
  if (p.exp_dn > 0) mcmcb_stage_exp_direct(exp_direct, p.exp_dn);
  const double* data = p.blob;
  if (SMEM) {
    unsigned char* blob_s = smem_raw + MCMCB_EXP_TAB_DOUBLES * sizeof(double);
    tma_stage_blob(blob_s, p.blob, p.blob_bytes, &mbar);  // contains a __syncthreads()
    data = reinterpret_cast<const double*>(blob_s);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane / L, gl = lane % L;
  mcmcb_ctx ctx;
  ctx.data = data; ctx.ndata = p.blob_n; ctx.prior = p.prior; ctx.lane = gl; ctx.nlanes = L;
  ctx.exp_tl = mcmcb_exp_column(exp_tab); ctx.exp_c1 = p.exp_c1; ctx.exp_c2 = p.exp_c2; ctx.scratch = nullptr;
  ctx.exp_dn = p.exp_dn;
  ctx.exp_td = p.exp_dn > 0 ? mcmcb_exp_direct_base(exp_direct, p.exp_dn) : 0u;

  const long long t4 = p.tier[0], t2 = p.tier[1], t1 = p.tier[2];
  const long long ws = (B >= 4 && p.rounds > 0) ? p.wtiles : 0;  // super-tiles come first in the work queue
  const long long base = ws * p.rounds * 4 * SUB;                // first chain after the super-tiles' range
  for (;;) {
    unsigned tile = 0;
    if (lane == 0) tile = atomicAdd(p.tile_counter, 1u);
    tile = __shfl_sync(FULL, tile, 0);
    long long t = tile;
    if (t >= ws + t4 + t2 + t1) break;
    if constexpr (B >= 4) {
      if (t < ws) { k1_run_tile<M, L, EREXIT, 4>(p, ctx, t * 4 * SUB, sub, gl, p.rounds, ws * 4 * SUB); continue; }
      t -= ws;
      if (t < t4) { k1_run_tile<M, L, EREXIT, 4>(p, ctx, base + t * 4 * SUB, sub, gl); continue; }
    }
    if constexpr (B >= 2) {
      if (t < t4 + t2) { k1_run_tile<M, L, EREXIT, 2>(p, ctx, base + (t4 * 4 + (t - t4) * 2) * SUB, sub, gl); continue; }
    }
    k1_run_tile<M, L, EREXIT, 1>(p, ctx, base + (t4 * 4 + t2 * 2 + (t - t4 - t2)) * SUB, sub, gl);
  }
}

}  // namespace mcmcb
