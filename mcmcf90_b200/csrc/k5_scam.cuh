// K5: thread-per-chain SCAM sweeps for LARGE populations that share one rotation (BASELINE C5: 262 144 chains, npar =
// 200, pooled adaptation).
//
// A SCAM step is a sweep over the npar components; every component move is one normal, the proposal theta + delta
// U(:,j), one full ssfunction and one accept test (MCMC_run_scam.F90:38-88).  A warp that owns ONE chain
// (k3_scam_step_kernel) spends most of its issue slots around the model, not in it: 32 candidate normal pairs are
// produced for the single normal a move needs, the accept logic and its Philox draw run redundantly on 32 lanes, a
// 198-group model leaves 26 of the last 32 lanes idle, every partial sum goes through a shuffle tree -- 1185 warp
// instructions per component move, ~250 of them the model (profiles/r02_summary.md D).  With one chain per THREAD none
// of that exists: the move costs what the model costs.
//
//   * theta lives in the thread's local memory (lane-interleaved, i.e. coalesced); it is the only per-chain vector:
//   * the proposal is never materialised when the model can evaluate a VIEW of its parameter vector (optional model
//     member ssfunction_view<V>, with MCMCB_VIEW_DEFAULTS = checkbounds always true and the default prior): the view
//     composes theta + delta U(:,j) on the fly (and, with MCMCB_K5_PEND > 0, accepted moves that have not been written
//     back yet -- measured slower, see K5_PEND).  Every element goes through the same fma sequence as an eager update:
//     results are bit-identical.  Without ssfunction_view the proposal is formed in a second local vector and the
//     model's ordinary functions are called;
//   * the rotation U and qcovstd are SHARED by all chains (pool_adapt = 1: stride 0): every thread reads the same
//     column, a broadcast out of L1; private rotations stay on the warp-per-chain kernel;
//   * the model blob is TMA-staged once per CTA into shared memory, as everywhere.
//
// Draws, acceptance rule, row logging for the adaptation tick and the state layout are k3_scam_step_kernel's; the tick
// kernels and the pooled merge do not know which step kernel ran.
#pragma once
#include "k3_scam.cuh"
#include "k4_ram.cuh"

namespace mcmcb {

constexpr int K5_THREADS = 128;

// Accepted moves composed on the fly before theta is rewritten.  Measured on BASELINE C5: composing 3 pending moves cut
// the kernel's DRAM traffic by 40 % but cost 40 % more instructions and a third of the occupancy (166 registers):
// 3.0e6 sweeps/s against 4.2e6 with theta rewritten on every accepted move (0), profiles/r02_summary.md.
#ifndef MCMCB_K5_PEND
#define MCMCB_K5_PEND 0
#endif
constexpr int K5_PEND = MCMCB_K5_PEND;
#ifndef MCMCB_K5_MINB
#define MCMCB_K5_MINB 4
#endif

// does the model evaluate a view of its parameter vector (ssfunction_view<V>)?
template <class M, class = void>
struct has_ssfunction_view { static constexpr bool value = false; };
template <class M>
struct has_ssfunction_view<M, decltype((void)&M::template ssfunction_view<mcmcb_view_probe>)> {
  static constexpr bool value = M::MCMCB_VIEW_DEFAULTS;
};

// theta + dl[0] u[0] + ... + dl[n-1] u[n-1], element by element, each term one fma in move order
struct K5View {
  const double* th;
  const double* u[K5_PEND + 1];
  double dl[K5_PEND + 1];
  int n;
  __device__ __forceinline__ double operator[](int k) const {
    double v = th[k];
#pragma unroll
    for (int q = 0; q < K5_PEND + 1; q++)
      if (q < n) v = fma(u[q][k], dl[q], v);
    return v;
  }
};

// default prior (priorfun.f90:97-100) of a view
template <class V>
__device__ __forceinline__ double k5_default_prior_view(const V& theta, int len, const mcmcb_ctx& c) {
  if (c.prior == nullptr) return 0.0;
  double p = 0.0;
  for (int i = 0; i < len; i++) {
    const double sg = c.prior[len + i];
    if (sg > 0.0) {
      const double t = (theta[i] - c.prior[i]) / sg;
      p += t * t;
    }
  }
  return p;
}

template <class M, bool SMEM>
__global__ void __launch_bounds__(K5_THREADS, MCMCB_K5_MINB) k5_scam_step_kernel(const __grid_constant__ K2Params p) {
  constexpr int NY = M::NY;
  constexpr K2Layout Lo = k2_layout(NY);
  constexpr bool AXPY = has_ssfunction_view<M>::value;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  const double* data = p.blob;
  if (SMEM) {
    tma_stage_blob(smem_raw, p.blob, p.blob_bytes, &mbar);  // every thread of the CTA takes part (barrier inside)
    data = reinterpret_cast<const double*>(smem_raw);
  }
  const long long cc = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cc >= p.nchains) return;
  const int d = p.d;
  const DevCfg& c = p.c;
  const size_t P = (size_t)p.pitch;
  double* st = p.st + cc;
  int* ist = p.ist + cc;
  const double* U = p.Rm;    // shared rotation, column-major: column j at U + j d
  const double* gq = p.qstd; // shared qcovstd
  double* gth = p.theta + cc * p.dp;
  double* rb = p.rowbuf + (size_t)cc * (p.rowcap + 1) * (d + 1);

  double th[K4_DM];
  double prop[AXPY ? 1 : K4_DM];
  for (int k = 0; k < d; k++) th[k] = gth[k];
  double ss1[NY], s2[NY];
#pragma unroll
  for (int k = 0; k < NY; k++) { ss1[k] = st[(Lo.ss + k) * P]; s2[k] = st[(Lo.s2 + k) * P]; }
  double pri1 = st[Lo.pri * P];
  int stayed = ist[Lo.i_stayed * P], bnd = ist[Lo.i_bnd * P], chainind = ist[Lo.i_chainind * P];
  int simuind = ist[Lo.i_simuind * P], status = ist[Lo.i_status * P], cnt = ist[Lo.i_cnt * P], pend = ist[Lo.i_pend * P];
  int nbuf = ist[Lo.i_nbuf * P];
  Rng g;
  g.nd = ((unsigned long long)(unsigned)ist[Lo.i_ndhi * P] << 32) | (unsigned)ist[Lo.i_ndlo * P];
  g.seed = p.seed; g.chain = (unsigned long long)(p.chain_offset + cc);
  g.inj = p.inj ? p.inj + (unsigned long long)cc * p.inj_per_chain : nullptr;
  g.inj_n = p.inj_per_chain;
  g.cache_valid = false; g.cache_lo = g.cache_hi = 0; g.cache_blk = 0;
  g.has_spare = ist[Lo.i_hasspare * P] != 0;
  g.spare = st[Lo.spare * P];
  g.exhausted = 0;
  const bool stored = cc < p.store_chains;
  double* srow = p.store_rows_p + (size_t)cc * p.store_rows * (d + NY);
  double* scnt = p.store_cnt_p + (size_t)cc * p.store_rows;
  double* ss2st = p.store_s2_p + (size_t)cc * p.store_rows * NY;

  mcmcb_ctx ctx;
  ctx.data = data; ctx.ndata = p.blob_n; ctx.prior = p.prior; ctx.lane = 0; ctx.nlanes = 1;
  ctx.exp_tl = 0u; ctx.exp_c1 = MCMCB_EXP_C1L; ctx.exp_c2 = MCMCB_EXP_C2L; ctx.scratch = nullptr;

  if (simuind == 0) {  // MCMC_run_scam.F90:26-36: initial point, saved as row 1
    double ssn[NY];
    M::ssfunction(th, d, NY, ctx, ssn);
#pragma unroll
    for (int k = 0; k < NY; k++) ss1[k] = ssn[k];
    pri1 = M::priorfun(th, d, ctx);
    chainind = 1; simuind = 1; cnt = 1; pend = 1;
    if (stored) {
      for (int k = 0; k < d; k++) srow[k] = th[k];
#pragma unroll
      for (int k = 0; k < NY; k++) { srow[d + k] = ss1[k]; if (c.updatesigma) ss2st[k] = s2[k]; }
    }
  }

  K5View vw;  // pending accepted moves (vw.n of them); the trial move is appended for an evaluation
  vw.th = th;
  vw.n = 0;
#pragma unroll
  for (int q = 0; q < K5_PEND + 1; q++) { vw.u[q] = U; vw.dl[q] = 0.0; }

  for (int done = 0; done < p.nsteps; done++) {
    bool rejall = true;
    bool logged = false;  // the row that is about to be replaced has been written to the row buffer
    for (int j = 0; j < d; j++) {
      // MCMC_propose_sc (MCMC_run_scam.F90:94-117) in the O(d) form theta + delta U(:,j)
      const double delta = g.normal() * gq[j];
      const double* col = U + (size_t)j * d;
      bool reject;
      double ssn[NY], prn = 0.0;
      if constexpr (AXPY) {
        K5View tv = vw;
#pragma unroll
        for (int q = 0; q < K5_PEND + 1; q++)
          if (q == vw.n) { tv.u[q] = col; tv.dl[q] = delta; }
        tv.n = vw.n + 1;
        prn = k5_default_prior_view(tv, d, ctx);
        M::ssfunction_view(tv, d, NY, ctx, ssn);
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < NY; k++) sum += (ssn[k] - ss1[k]) / s2[k];
        reject = mh_reject(alpha_from_tst(-0.5 * (sum + (prn - pri1))), g);
      } else {
        for (int k = 0; k < d; k++) prop[k] = fma(col[k], delta, th[k]);
        if (!M::checkbounds(prop, d, ctx)) {
          bnd++;  // dodr is forced off for SCAM (mcmcinit.F90:328-330)
          reject = true;
        } else {
          prn = M::priorfun(prop, d, ctx);
          M::ssfunction(prop, d, NY, ctx, ssn);
          double sum = 0.0;
#pragma unroll
          for (int k = 0; k < NY; k++) sum += (ssn[k] - ss1[k]) / s2[k];
          reject = mh_reject(alpha_from_tst(-0.5 * (sum + (prn - pri1))), g);
        }
      }
      if (!reject) {  // MCMC_run_scam.F90:63-68
        if (!logged) {
          // first acceptance of this sweep: the current row is complete -- log it for the adaptation
          // kernel before theta changes (the reference reads it back from the stored chain)
          const bool absorbing = c.doadapt && !(c.adaptend > 0 && simuind + 1 > c.adaptend);
          if (absorbing) {
            if (nbuf < p.rowcap) {
              for (int k = 0; k < d; k++) rb[(size_t)nbuf * (d + 1) + k] = AXPY ? vw[k] : th[k];
              rb[(size_t)nbuf * (d + 1) + d] = (double)((c.doadapt && c.adapthist > 1) ? cnt : pend);
              nbuf++;
            } else {
              status |= MCMCB_ST_STORE_FULL;
            }
          }
          logged = true;
        }
        if constexpr (AXPY) {
          if (vw.n == K5_PEND) {  // the pending moves and this one are written back together
            for (int k = 0; k < d; k++) th[k] = fma(col[k], delta, vw[k]);
            vw.n = 0;
          } else {
#pragma unroll
            for (int q = 0; q < K5_PEND; q++)
              if (q == vw.n) { vw.u[q] = col; vw.dl[q] = delta; }
            vw.n++;
          }
        } else {
          for (int k = 0; k < d; k++) th[k] = prop[k];
        }
#pragma unroll
        for (int k = 0; k < NY; k++) ss1[k] = ssn[k];
        pri1 = prn;
        rejall = false;
      }
    }
    // ---------------- end of sweep, MCMC_run_scam.F90:74-86
    const int i = simuind + 1;
    simuind = i;
    if (rejall) {
      stayed++;
      cnt++; pend++;
    } else {
      if (stored && chainind - 1 < p.store_rows) scnt[chainind - 1] = (double)cnt;
      chainind++;
      cnt = 1; pend = 1;
    }
    if (c.updatesigma) {
#pragma unroll
      for (int k = 0; k < NY; k++) {
        const double gg = g.gamma(c.N0 / 2.0 + (double)p.nobs[k] / 2.0, 2.0 / (c.N0 * c.S02 + ss1[k]));
        s2[k] = 1.0 / gg;
      }
    }
    if (stored) {
      if (!rejall) {
        if (chainind - 1 < p.store_rows) {
          for (int k = 0; k < d; k++) srow[(size_t)(chainind - 1) * (d + NY) + k] = AXPY ? vw[k] : th[k];
#pragma unroll
          for (int k = 0; k < NY; k++) srow[(size_t)(chainind - 1) * (d + NY) + d + k] = ss1[k];
        } else {
          status |= MCMCB_ST_STORE_FULL;
        }
      }
      if (c.updatesigma && i - 1 < p.store_rows) {
#pragma unroll
        for (int k = 0; k < NY; k++) ss2st[(size_t)(i - 1) * NY + k] = s2[k];
      }
    }
    if (g.exhausted) status |= MCMCB_ST_RNG_EXHAUSTED;
  }

  // ---- write state back (pending moves included)
  for (int k = 0; k < d; k++) gth[k] = AXPY ? vw[k] : th[k];
#pragma unroll
  for (int k = 0; k < NY; k++) { st[(Lo.ss + k) * P] = ss1[k]; st[(Lo.s2 + k) * P] = s2[k]; }
  st[Lo.pri * P] = pri1; st[Lo.spare * P] = g.spare;
  ist[Lo.i_stayed * P] = stayed; ist[Lo.i_bnd * P] = bnd;
  ist[Lo.i_chainind * P] = chainind; ist[Lo.i_simuind * P] = simuind; ist[Lo.i_status * P] = status;
  ist[Lo.i_hasspare * P] = g.has_spare ? 1 : 0;
  ist[Lo.i_cnt * P] = cnt; ist[Lo.i_pend * P] = pend; ist[Lo.i_nbuf * P] = nbuf;
  ist[Lo.i_ndlo * P] = (int)(unsigned)(g.nd & 0xffffffffull);
  ist[Lo.i_ndhi * P] = (int)(unsigned)(g.nd >> 32);
}

}  // namespace mcmcb
