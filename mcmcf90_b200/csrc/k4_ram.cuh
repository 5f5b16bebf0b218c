// K4: thread-per-chain RAM sampler for run-time npar and LARGE populations (BASELINE C4: 65 536 chains, npar = 50).
//
// The RAM loop (MCMC_run_ram.F90:45-80) rewrites its Cholesky factor at every step with a rank-1 update or downdate
// (dchud.f:122-139 / dchdd.f:141-179).  Both are scalar recurrences: every Givens rotation needs the result of the
// previous one, through a division, a square root and another division.  A warp that owns ONE chain
// (k2_step_kernel) spends the step waiting on those latencies with 31 lanes idle -- measured 1.8e7 chain-steps/s
// = 6 % of the HBM roofline.  When there are enough chains to give every lane its own, the recurrences of 32 chains
// run side by side in one warp at full SIMT width and the factor traffic becomes the bound, as SURVEY.md 8d
// predicted for private factors:
//
//   * one thread = one chain; theta, z, the proposal and the rotation vectors c, s live in the thread's local memory
//     (lane-interleaved by the hardware, i.e. coalesced);
//   * the factors of the launch live in HBM as packed upper triangles in structure-of-arrays form
//     Rp[k][chain], k = j (j + 1) / 2 + i, so that the 32 chains of a warp read / write 256 contiguous bytes per
//     element; every pass over a factor (proposal, solve, rotation sweep) streams it in ascending k, the order in
//     which dchud / dchdd walk it (column by column);
//   * the layout of everything else is k2_step_kernel's: the launcher converts the row-major [chain][d*d] factors
//     into Rp before the launch and back after it (k4_pack / k4_unpack: 2 x 30 KB per chain per launch against
//     ~35 KB per chain per STEP), so pooled ticks, fetches and checkpoints do not know this kernel exists.
//
// Algorithmic HBM bytes per chain-step: 16 d(d+1)/2 + state.  This kernel moves 8 T for the proposal,
// 16 T for an update, 24 T for a downdate (T = d(d+1)/2).
//
// Operation order: proposal = dtrmv('u','t','n') (matutils.F90:108-109), dchud / dchdd / drotg / dnrm2 as the
// register kernel's (k1_small.cuh), i.e. the reference's; draws in the reference's order from the chain's own stream.
#pragma once
#include "k2_large.cuh"

namespace mcmcb {

constexpr int K4_THREADS = 128;
constexpr int K4_DM = 32 * K2_MAXM;  // largest npar (local arrays are sized for it; only npar entries are touched)

__host__ __device__ constexpr size_t k4_pk(int i, int j) { return (size_t)j * (j + 1) / 2 + i; }  // i <= j

// row-major upper [chain][d*d] -> packed SoA [T][pitch]
static __global__ void k4_pack_kernel(const double* Rm, long long r_stride, double* Rp, long long pitch, long long n, int d) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double* R = Rm + (size_t)c * r_stride;
  for (int j = 0; j < d; j++)
    for (int i = 0; i <= j; i++) Rp[k4_pk(i, j) * pitch + c] = R[(size_t)i * d + j];
}
static __global__ void k4_unpack_kernel(double* Rm, long long r_stride, const double* Rp, long long pitch, long long n, int d) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double* R = Rm + (size_t)c * r_stride;
  for (int j = 0; j < d; j++)
    for (int i = 0; i <= j; i++) R[(size_t)i * d + j] = Rp[k4_pk(i, j) * pitch + c];
}

template <class M>
__global__ void __launch_bounds__(K4_THREADS) k4_ram_step_kernel(const __grid_constant__ K2Params p, double* __restrict__ Rp) {
  constexpr int NY = M::NY;
  constexpr K2Layout Lo = k2_layout(NY);
  const long long cc = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cc >= p.nchains) return;
  const int d = p.d;
  const DevCfg& c = p.c;
  const size_t P = (size_t)p.pitch;
  double* R = Rp + cc;  // element k of this chain's factor at R[k * P]
  double* st = p.st + cc;
  int* ist = p.ist + cc;
  double* gth = p.theta + cc * p.dp;

  double th[K4_DM], z[K4_DM], prop[K4_DM], cv[K4_DM], sv[K4_DM];
  for (int k = 0; k < d; k++) { th[k] = gth[k]; prop[k] = th[k]; }
  double ss1[NY], s2[NY];
#pragma unroll
  for (int k = 0; k < NY; k++) { ss1[k] = st[(Lo.ss + k) * P]; s2[k] = st[(Lo.s2 + k) * P]; }
  double pri1 = st[Lo.pri * P], rama = st[Lo.rama * P];
  int stayed = ist[Lo.i_stayed * P], bnd = ist[Lo.i_bnd * P], chainind = ist[Lo.i_chainind * P];
  int simuind = ist[Lo.i_simuind * P], status = ist[Lo.i_status * P], cnt = ist[Lo.i_cnt * P], pend = ist[Lo.i_pend * P];
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
  ctx.data = p.blob; ctx.ndata = p.blob_n; ctx.prior = p.prior; ctx.lane = 0; ctx.nlanes = 1;
  ctx.exp_tl = 0u; ctx.exp_c1 = MCMCB_EXP_C1L; ctx.exp_c2 = MCMCB_EXP_C2L; ctx.scratch = nullptr;

  int done = 0;
  if (simuind == 0) {  // MCMC_run_ram.F90:31-40: the initial point is row 1
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
  while (done < p.nsteps) {
    // ---------------- proposal theta + R'z (MCMC_propose_ram, MCMC_run_ram.F90:87-101; dtrmv order)
    double su2 = 0.0;
    for (int k = 0; k < d; k++) { z[k] = g.normal(); }
    for (int k = 0; k < d; k++) su2 += z[k] * z[k];  // sum(u**2), MCMC_run_ram.F90:168-170
    for (int j = d - 1; j >= 0; j--) {
      const double* col = R + k4_pk(0, j) * P;
      double acc = z[j] * col[(size_t)j * P];
      for (int i = j - 1; i >= 0; i--) acc += col[(size_t)i * P] * z[i];
      prop[j] = th[j] + acc;
    }
    const bool inb = M::checkbounds(prop, d, ctx);
    double ssn[NY];
    double prn = 0.0;
    bool reject;
    if (!inb) {  // MCMC_run_ram.F90:52-55: alpha12 keeps its previous value (Q11)
      bnd++;
      reject = true;
    } else {
      prn = M::priorfun(prop, d, ctx);
      M::ssfunction(prop, d, NY, ctx, ssn);
      double sum = 0.0;
#pragma unroll
      for (int k = 0; k < NY; k++) sum += (ssn[k] - ss1[k]) / s2[k];
      rama = alpha_from_tst(-0.5 * (sum + (prn - pri1)));
      reject = mh_reject(rama, g);
    }
    // ---------------- end of step, MCMC_run_ram.F90:66-78
    const int i = simuind + 1;
    simuind = i;
    if (reject) {
      stayed++;
      cnt++; pend++;
    } else {
      if (stored && chainind - 1 < p.store_rows) scnt[chainind - 1] = (double)cnt;
      for (int k = 0; k < d; k++) th[k] = prop[k];
#pragma unroll
      for (int k = 0; k < NY; k++) ss1[k] = ssn[k];
      pri1 = prn;
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
      if (!reject) {
        if (chainind - 1 < p.store_rows) {
          for (int k = 0; k < d; k++) srow[(size_t)(chainind - 1) * (d + NY) + k] = th[k];
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
    // ---------------- MCMC_adapt_ram, MCMC_run_ram.F90:104-179
    if (c.doadapt && !(i < c.burnintime && c.doburnin)) {
      const double a = 1.0 / pow((double)(float)i, c.nuparam) * (rama - c.alphatarget);
      if (a >= 0.0) {  // cholupdate(R, u/sum(u**2)*a): dchud.f:122-139, column by column
        for (int j = 0; j < d; j++) {
          double* col = R + k4_pk(0, j) * P;
          double xj = z[j] / su2 * a;
          for (int r = 0; r < j; r++) {
            const double rij = col[(size_t)r * P];
            const double t = cv[r] * rij + sv[r] * xj;
            xj = cv[r] * xj - sv[r] * rij;
            col[(size_t)r * P] = t;
          }
          double rjj = col[(size_t)j * P], cj, sj;
          drotg(rjj, xj, cj, sj);
          col[(size_t)j * P] = rjj;
          cv[j] = cj; sv[j] = sj;
        }
      } else {  // choldowndate(R, -u/sum(u**2)*a): dchdd.f:141-179
        sv[0] = (-z[0] / su2 * a) / R[0];
        for (int j = 1; j < d; j++) {
          const double* col = R + k4_pk(0, j) * P;
          double t = 0.0;
          for (int r = 0; r < j; r++) t += col[(size_t)r * P] * sv[r];
          sv[j] = ((-z[j] / su2 * a) - t) / col[(size_t)j * P];
        }
        double norm;
        if (d == 1) {
          norm = fabs(sv[0]);
        } else {  // classic dnrm2
          double scale = 0.0, ssq = 1.0;
          for (int k = 0; k < d; k++) {
            if (sv[k] != 0.0) {
              const double av = fabs(sv[k]);
              if (scale < av) { const double t = scale / av; ssq = 1.0 + ssq * t * t; scale = av; }
              else { const double t = av / scale; ssq = ssq + t * t; }
            }
          }
          norm = scale * sqrt(ssq);
        }
        if (!(norm < 1.0)) {
          status |= MCMCB_ST_DOWNDATE_FAIL;  // the reference stops here (matutils.F90:716-722): flag, skip the update
        } else {
          double alpha = sqrt(1.0 - norm * norm);
          for (int r = d - 1; r >= 0; r--) {
            const double scale = alpha + fabs(sv[r]);
            const double aa = alpha / scale, bb = sv[r] / scale;
            const double nr = sqrt(aa * aa + bb * bb);
            cv[r] = aa / nr;
            sv[r] = bb / nr;
            alpha = scale * nr;
          }
          for (int j = 0; j < d; j++) {
            double* col = R + k4_pk(0, j) * P;
            double xx = 0.0;
            for (int r = j; r >= 0; r--) {
              const double rij = col[(size_t)r * P];
              const double t = cv[r] * xx + sv[r] * rij;
              col[(size_t)r * P] = cv[r] * rij - sv[r] * xx;
              xx = t;
            }
          }
        }
      }
    }
    if (g.exhausted) status |= MCMCB_ST_RNG_EXHAUSTED;
    done++;
  }

  // ---- write state back
  for (int k = 0; k < d; k++) gth[k] = th[k];
#pragma unroll
  for (int k = 0; k < NY; k++) { st[(Lo.ss + k) * P] = ss1[k]; st[(Lo.s2 + k) * P] = s2[k]; }
  st[Lo.pri * P] = pri1; st[Lo.rama * P] = rama; st[Lo.spare * P] = g.spare;
  ist[Lo.i_stayed * P] = stayed; ist[Lo.i_bnd * P] = bnd; ist[Lo.i_chainind * P] = chainind;
  ist[Lo.i_simuind * P] = simuind; ist[Lo.i_status * P] = status;
  ist[Lo.i_hasspare * P] = g.has_spare ? 1 : 0;
  ist[Lo.i_cnt * P] = cnt; ist[Lo.i_pend * P] = pend;
  ist[Lo.i_ndlo * P] = (int)(unsigned)(g.nd & 0xffffffffull);
  ist[Lo.i_ndhi * P] = (int)(unsigned)(g.nd >> 32);
}

}  // namespace mcmcb
