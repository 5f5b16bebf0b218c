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
// Algorithmic HBM bytes per chain-step: 16 d(d+1)/2 + state.  This kernel reads the factor twice and
// writes it twice per step (32 T bytes, T = d(d+1)/2): the sweeps that rewrite it also form the next proposal.
//
// Operation order: proposal = dtrmv('u','t','n') (matutils.F90:108-109), dchud / dchdd / drotg / dnrm2 as the
// register kernel's (k1_small.cuh), i.e. the reference's; draws in the reference's order from the chain's own stream.
#pragma once
#include "k2_large.cuh"

namespace mcmcb {

constexpr int K4_THREADS = 128;
#ifndef MCMCB_K4_U
#define MCMCB_K4_U 8
#endif
constexpr int K4_U = MCMCB_K4_U;               // factor elements loaded ahead of the dependent recurrences (memory-level parallelism)
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

  double th[K4_DM], zbuf[2][K4_DM], prop[K4_DM], cv[K4_DM], sv[K4_DM];
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
  // The factor is streamed TWICE per step, not four times: the pass that rewrites it (update: one ascending sweep;
  // downdate: solve sweep + rotation sweep) also forms the NEXT step's proposal R'z from the columns it has just
  // finished, with that step's normals drawn right after this step's last draw (the adaptation itself draws nothing,
  // so the chain's draw order is the reference's).  Lanes of a warp that update and lanes that downdate share the
  // loads of the first sweep.  The last step of a launch does not look ahead: a launch leaves no pre-drawn state.
  int zi = 0;
  bool have_prop = false;
  double su2 = 0.0;
  while (done < p.nsteps) {
    double* z = zbuf[zi];
    if (!have_prop) {
      // ---------------- proposal theta + R'z (MCMC_propose_ram, MCMC_run_ram.F90:87-101; dtrmv order)
      su2 = 0.0;
      for (int k = 0; k < d; k++) { z[k] = g.normal(); }
      for (int k = 0; k < d; k++) su2 += z[k] * z[k];  // sum(u**2), MCMC_run_ram.F90:168-170
      // The column sums are formed in the order the fused sweeps below would have used had the previous step run in
      // this launch (descending rows after a downdate = dtrmv's own order, ascending otherwise), so that a run does
      // not depend, bit for bit, on how it is cut into launches.
      bool desc = false;
      if (simuind > 1 && c.doadapt && !(simuind < c.burnintime && c.doburnin))
        desc = 1.0 / pow((double)(float)simuind, c.nuparam) * (rama - c.alphatarget) < 0.0;
      for (int j = 0; j < d; j++) {
        const double* col = R + k4_pk(0, j) * P;
        double acc = 0.0;
        if (desc) {
          int r = j;
          for (; r >= K4_U - 1; r -= K4_U) {  // K4_U loads in flight before the dependent sum consumes them
            double rv[K4_U];
#pragma unroll
            for (int u = 0; u < K4_U; u++) rv[u] = col[(size_t)(r - u) * P];
#pragma unroll
            for (int u = 0; u < K4_U; u++) acc += rv[u] * z[r - u];
          }
          for (; r >= 0; r--) acc += col[(size_t)r * P] * z[r];
          prop[j] = th[j] + acc;
        } else {
          int r = 0;
          for (; r + K4_U <= j; r += K4_U) {
            double rv[K4_U];
#pragma unroll
            for (int u = 0; u < K4_U; u++) rv[u] = col[(size_t)(r + u) * P];
#pragma unroll
            for (int u = 0; u < K4_U; u++) acc += rv[u] * z[r + u];
          }
          for (; r < j; r++) acc += col[(size_t)r * P] * z[r];
          prop[j] = th[j] + (acc + col[(size_t)j * P] * z[j]);
        }
      }
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
    // ---------------- MCMC_adapt_ram (MCMC_run_ram.F90:104-179) fused with the next step's proposal
    int mode = 0;  // 0: factor unchanged, 1: cholupdate (dchud.f:122-139), 2: choldowndate (dchdd.f:141-179)
    double a = 0.0;
    if (c.doadapt && !(i < c.burnintime && c.doburnin)) {
      a = 1.0 / pow((double)(float)i, c.nuparam) * (rama - c.alphatarget);
      mode = (a >= 0.0) ? 1 : 2;
    }
    const bool last = done + 1 >= p.nsteps;
    double* zn = zbuf[zi ^ 1];
    double su2n = 0.0;
    if (!last) {
      for (int k = 0; k < d; k++) { zn[k] = g.normal(); }
      for (int k = 0; k < d; k++) su2n += zn[k] * zn[k];
    } else {
      for (int k = 0; k < d; k++) zn[k] = 0.0;  // the look-ahead sums below are formed and dropped
    }
    if (mode != 0 || !last) {
      // ---- sweep 1, ascending: update lanes rotate and finish their columns; downdate lanes solve R'a = x;
      //      lanes whose factor stays as it is only form the next proposal
      for (int j = 0; j < d; j++) {
        double* col = R + k4_pk(0, j) * P;
        double xj = (mode == 1) ? z[j] / su2 * a : ((mode == 2) ? -z[j] / su2 * a : 0.0);
        double acc = 0.0;  // modes 0, 1: next proposal's column sum; mode 2: the solve's ddot
        int r = 0;
        for (; r + K4_U <= j; r += K4_U) {
          double rv[K4_U];
#pragma unroll
          for (int u = 0; u < K4_U; u++) rv[u] = col[(size_t)(r + u) * P];
          if (mode == 1) {
#pragma unroll
            for (int u = 0; u < K4_U; u++) {
              const double t = cv[r + u] * rv[u] + sv[r + u] * xj;
              xj = cv[r + u] * xj - sv[r + u] * rv[u];
              col[(size_t)(r + u) * P] = t;
              acc += t * zn[r + u];
            }
          } else if (mode == 2) {
#pragma unroll
            for (int u = 0; u < K4_U; u++) acc += rv[u] * sv[r + u];
          } else {
#pragma unroll
            for (int u = 0; u < K4_U; u++) acc += rv[u] * zn[r + u];
          }
        }
        for (; r < j; r++) {
          const double rij = col[(size_t)r * P];
          if (mode == 1) {
            const double t = cv[r] * rij + sv[r] * xj;
            xj = cv[r] * xj - sv[r] * rij;
            col[(size_t)r * P] = t;
            acc += t * zn[r];
          } else if (mode == 2) {
            acc += rij * sv[r];
          } else {
            acc += rij * zn[r];
          }
        }
        double rjj = col[(size_t)j * P];
        if (mode == 1) {
          double cj, sj;
          drotg(rjj, xj, cj, sj);
          col[(size_t)j * P] = rjj;
          cv[j] = cj; sv[j] = sj;
          prop[j] = th[j] + (acc + rjj * zn[j]);
        } else if (mode == 2) {
          sv[j] = (xj - acc) / rjj;  // dchdd.f:142-147
        } else {
          prop[j] = th[j] + (acc + rjj * zn[j]);
        }
      }
      if (mode == 2) {
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
        const bool ok = norm < 1.0;
        if (!ok) {
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
        }
        // ---- sweep 2, columns descending like dchdd.f:171-179: apply the rotations (when the downdate exists) and
        //      form the next proposal from the finished columns, in dtrmv's own order
        for (int j = 0; j < d; j++) {
          double* col = R + k4_pk(0, j) * P;
          double xx = 0.0, acc = 0.0;
          int r = j;
          for (; r >= K4_U - 1; r -= K4_U) {
            double rv[K4_U];
#pragma unroll
            for (int u = 0; u < K4_U; u++) rv[u] = col[(size_t)(r - u) * P];
#pragma unroll
            for (int u = 0; u < K4_U; u++) {
              double nr = rv[u];
              if (ok) {
                const double t = cv[r - u] * xx + sv[r - u] * rv[u];
                nr = cv[r - u] * rv[u] - sv[r - u] * xx;
                col[(size_t)(r - u) * P] = nr;
                xx = t;
              }
              acc += nr * zn[r - u];
            }
          }
          for (; r >= 0; r--) {
            double nr = col[(size_t)r * P];
            if (ok) {
              const double t = cv[r] * xx + sv[r] * nr;
              const double v = cv[r] * nr - sv[r] * xx;
              col[(size_t)r * P] = v;
              xx = t;
              nr = v;
            }
            acc += nr * zn[r];
          }
          prop[j] = th[j] + acc;
        }
      }
    }
    have_prop = !last;
    if (!last) { zi ^= 1; su2 = su2n; }
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
