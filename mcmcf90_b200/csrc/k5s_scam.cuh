// K5S: the thread-per-chain SCAM sweep of k5_scam.cuh with every chain's theta RESIDENT IN SHARED MEMORY.
//
// k5_scam_step_kernel keeps theta in local memory: a component move reads the npar elements once for the model and, when
// the move is accepted, rewrites them -- 3.2 KB per move at npar = 200 against ~4400 FP64 instructions of model.  At the
// FP64 rate that is ~9 TB/s of theta traffic, more than HBM delivers, and the chains resident on the GPU (4 CTAs x 128
// threads x 148 SMs x 1.6 KB = 121 MB) do not stay in L2: the ncu capture of round 2 shows 3.05 TB/s of DRAM traffic, L2
// hit 63 %, FP64 pipe 37 % (profiles/r02_summary.md).  Here one CTA per SM holds its chains' theta in shared memory for
// the whole launch (blockDim x npar doubles, element k of thread t at th[k * blockDim + t]: conflict-free), so a move
// touches no global memory except the rotation's column:
//
//   * U(:,j) is the same for every chain of the population (pooled rotation, stride 0).  Each warp stages the column
//     of its current move in its own shared buffer; the column of the NEXT move is fetched into registers while the model
//     runs, so its L2 latency is never exposed (a warp is alone on its scheduler here: nothing else would hide it);
//   * the model evaluates the view theta + delta U(:,j) (ssfunction_view<V>, mcmcb200_model.cuh) with V::ILP = 8
//     accumulation chains in flight: with four warps per SM the instruction-level parallelism has to come from the
//     thread itself (255 registers are available to it);
//   * an accepted move rewrites theta in shared memory.
//
// Draw order, acceptance rule, row logging and the state layout are k5_scam_step_kernel's (MCMC_run_scam.F90:26-88): both
// kernels apply, element by element, the fma sequence of an eager update, so they agree to the last bit
// (tests/test_r02_coverage.py).  Models without ssfunction_view, private rotations and populations whose theta does not
// fit stay on the other kernels.
#pragma once
#include "k5_scam.cuh"

namespace mcmcb {

constexpr int K5S_THREADS = 128;

template <int W>
struct K5SView {
  static constexpr int ILP = W;
  const double* th;  // shared: element k of this thread's chain at th[k * T]
  const double* u;   // shared: the column of this warp's move
  double dl;
  int T;
  __device__ __forceinline__ double operator[](int k) const { return fma(u[k], dl, th[(size_t)k * T]); }
};

// bytes of dynamic shared memory: blob | per-warp column buffers | theta
__host__ __device__ __forceinline__ size_t k5s_smem_bytes(int d, int threads, size_t blob_bytes) {
  const size_t dpad = (size_t)(d + 1) & ~(size_t)1;
  return ((blob_bytes + 15) & ~(size_t)15) + sizeof(double) * ((size_t)(threads / 32) * dpad + (size_t)d * threads);
}

template <class M, int W>
__global__ void __launch_bounds__(K5S_THREADS, 1) k5s_scam_step_kernel(const __grid_constant__ K2Params p) {
  constexpr int NY = M::NY;
  constexpr K2Layout Lo = k2_layout(NY);
  constexpr int NPF = K4_DM / 32;  // column elements a lane prefetches
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  tma_stage_blob(smem_raw, p.blob, p.blob_bytes, &mbar);  // every thread of the CTA takes part (barrier inside)
  const double* data = reinterpret_cast<const double*>(smem_raw);
  const int d = p.d, T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dpad = (d + 1) & ~1;
  double* ucol = reinterpret_cast<double*>(smem_raw + (((size_t)p.blob_bytes + 15) & ~(size_t)15)) + (size_t)warp * dpad;
  double* th = reinterpret_cast<double*>(smem_raw + (((size_t)p.blob_bytes + 15) & ~(size_t)15)) + (size_t)(T >> 5) * dpad + tid;

  // a thread past the last chain shadows the last chain (its warp's column staging needs every lane) and writes nothing
  const long long c0 = (long long)blockIdx.x * T + tid;
  const bool live = c0 < p.nchains;
  const long long cc = live ? c0 : p.nchains - 1;
  const DevCfg& c = p.c;
  const size_t P = (size_t)p.pitch;
  double* st = p.st + cc;
  int* ist = p.ist + cc;
  const double* U = p.Rm;    // shared rotation, column-major: column j at U + j d
  const double* gq = p.qstd; // shared qcovstd
  double* gth = p.theta + cc * p.dp;
  double* rb = p.rowbuf + (size_t)cc * (p.rowcap + 1) * (d + 1);

  for (int k = 0; k < d; k++) th[(size_t)k * T] = gth[k];
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
  const bool stored = live && cc < p.store_chains;
  double* srow = p.store_rows_p + (size_t)cc * p.store_rows * (d + NY);
  double* scnt = p.store_cnt_p + (size_t)cc * p.store_rows;
  double* ss2st = p.store_s2_p + (size_t)cc * p.store_rows * NY;

  mcmcb_ctx ctx;
  ctx.data = data; ctx.ndata = p.blob_n; ctx.prior = p.prior; ctx.lane = 0; ctx.nlanes = 1;
  ctx.exp_tl = 0u; ctx.exp_c1 = MCMCB_EXP_C1L; ctx.exp_c2 = MCMCB_EXP_C2L; ctx.scratch = nullptr;

  // the current point as a view: a zero move along a zeroed column (only the first launch evaluates it)
  K5SView<W> tv;
  tv.th = th; tv.u = ucol; tv.dl = 0.0; tv.T = T;

  for (int k = lane; k < d; k += 32) ucol[k] = 0.0;  // theta + 0 * 0: the current point itself
  __syncwarp();
  if (simuind == 0) {  // MCMC_run_scam.F90:26-36: initial point, saved as row 1
    double ssn[NY];
    M::ssfunction_view(tv, d, NY, ctx, ssn);
#pragma unroll
    for (int k = 0; k < NY; k++) ss1[k] = ssn[k];
    pri1 = k5_default_prior_view(tv, d, ctx);
    chainind = 1; simuind = 1; cnt = 1; pend = 1;
    if (stored) {
      for (int k = 0; k < d; k++) srow[k] = th[(size_t)k * T];
#pragma unroll
      for (int k = 0; k < NY; k++) { srow[d + k] = ss1[k]; if (c.updatesigma) ss2st[k] = s2[k]; }
    }
  }

  // column 0 and its scale, fetched ahead
  double pf[NPF], qn = gq[0];
#pragma unroll
  for (int i = 0; i < NPF; i++) { const int k = lane + 32 * i; pf[i] = k < d ? U[k] : 0.0; }

  for (int done = 0; done < p.nsteps; done++) {
    bool rejall = true;
    bool logged = false;  // the row that is about to be replaced has been written to the row buffer
    for (int j = 0; j < d; j++) {
      // ---- stage this move's column (everybody is done with the previous one), fetch the next
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NPF; i++) { const int k = lane + 32 * i; if (k < d) ucol[k] = pf[i]; }
      const double qj = qn;
      __syncwarp();
      {
        const int jn = j + 1 < d ? j + 1 : 0;
        const double* ncol = U + (size_t)jn * d;
#pragma unroll
        for (int i = 0; i < NPF; i++) { const int k = lane + 32 * i; pf[i] = k < d ? ncol[k] : 0.0; }
        qn = gq[jn];
      }
      // MCMC_propose_sc (MCMC_run_scam.F90:94-117) in the O(d) form theta + delta U(:,j)
      const double delta = g.normal() * qj;
      tv.dl = delta;
      double ssn[NY];
      const double prn = k5_default_prior_view(tv, d, ctx);
      M::ssfunction_view(tv, d, NY, ctx, ssn);
      double sum = 0.0;
#pragma unroll
      for (int k = 0; k < NY; k++) sum += (ssn[k] - ss1[k]) / s2[k];
      const bool reject = mh_reject(alpha_from_tst(-0.5 * (sum + (prn - pri1))), g);
      if (!reject) {  // MCMC_run_scam.F90:63-68
        if (!logged) {
          // first acceptance of this sweep: the current row is complete -- log it for the adaptation
          // kernel before theta changes (the reference reads it back from the stored chain)
          const bool absorbing = c.doadapt && !(c.adaptend > 0 && simuind + 1 > c.adaptend);
          if (absorbing) {
            if (nbuf < p.rowcap) {
              if (live) {
                for (int k = 0; k < d; k++) rb[(size_t)nbuf * (d + 1) + k] = th[(size_t)k * T];
                rb[(size_t)nbuf * (d + 1) + d] = (double)((c.doadapt && c.adapthist > 1) ? cnt : pend);
              }
              nbuf++;
            } else {
              status |= MCMCB_ST_STORE_FULL;
            }
          }
          logged = true;
        }
        for (int k = 0; k < d; k++) th[(size_t)k * T] = fma(ucol[k], delta, th[(size_t)k * T]);
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
          for (int k = 0; k < d; k++) srow[(size_t)(chainind - 1) * (d + NY) + k] = th[(size_t)k * T];
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

  if (!live) return;
  // ---- write state back
  for (int k = 0; k < d; k++) gth[k] = th[(size_t)k * T];
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
