// K5S: the SCAM sweep of k5_scam.cuh for large populations that share one rotation, with every chain's theta RESIDENT IN
// SHARED MEMORY, L lanes per chain and C chains per thread.
//
// k5_scam_step_kernel (one thread per chain) keeps theta in local memory: a component move reads the npar elements once
// for the model and, when the move is accepted, rewrites them -- 3.2 KB per move at npar = 200 against ~5800 FP64
// instructions.  At the FP64 rate that is ~9 TB/s of theta traffic, and the chains resident on the GPU (4 CTAs x 128
// threads x 148 SMs x 1.6 KB = 121 MB) do not stay in L2: 3.05 TB/s of DRAM traffic, L2 hit 63 %, FP64 pipe 37 %, the
// board at its power cap (profiles/r02_summary.md).  Here one CTA per SM holds 128 chains' theta in shared memory for the
// whole launch (element k of slot s at th[k * 128 + s'], s' = s rotated inside its group of 16 slots by (k mod L) * 16 / L:
// the 16 addresses of a half-warp -- L consecutive rows x 16 / L neighbouring slots -- fall into 16 different bank pairs),
// and a move touches no global memory except the rotation's column:
//
//   * U(:,j) and qcovstd(j) are the same for every chain (pooled rotation, stride 0).  The CTA keeps two column buffers:
//     every thread fetches its share of the NEXT move's column into a register at the start of a move (the L2 latency
//     runs under the model), publishes it after the accept step, and one __syncthreads per move flips the buffers.
//     Dropping the barrier and reading the column through L1 was measured slower (5.10 s against 4.50 s per 100 sweeps);
//   * L lanes share a chain (default 4): they take the model's groups round-robin (mcmcb_ctx::lane / nlanes), their
//     partial sums are added with a shuffle butterfly (every lane gets the same bits), each lane rewrites its own elements
//     of theta on acceptance; the chain's generator runs on all of its lanes (same draws, same decisions, no extra issue
//     slots: the lanes sit in one warp).  128 threads per CTA (L = 1) leave one warp per scheduler: latency-bound;
//   * C chains per thread (default 2) go through the model in ONE sweep over its data (optional model member
//     ssfunction_view_batch<C, V>): with one chain per thread the kernel is bound by shared-memory wavefronts (a 64-bit
//     shared load costs two wavefronts whatever its addresses: ncu, L1 data stage 90 % busy, 2.3e9 wavefronts per
//     wave), the datum loads being 2/3 of them.
//
// Draw order, acceptance rule, row logging and the state layout are k5_scam_step_kernel's (MCMC_run_scam.F90:26-88): both
// kernels apply, element by element, the fma sequence of an eager update.  With L = 1 they agree to the last bit; with
// L > 1 the sum of squares is the sum of the lanes' partial sums (rounding-level differences, as in the warp-per-chain
// kernel); C does not change a bit (tests/test_r02_coverage.py).  Models without ssfunction_view, private rotations and
// populations whose theta does not fit stay on the other kernels.
#pragma once
#include "k5_scam.cuh"

namespace mcmcb {

constexpr int K5S_CHAINS = 128;  // chains per CTA (fewer when npar is larger than ~200: k5s_chains)

// Element k of the chain in slot s of the CTA's theta block.  Rows are NC slots long (NC a multiple of 16).  The L lanes
// of a chain read L consecutive rows at the same time and a half-warp holds 16 / L neighbouring slots, so row k is
// rotated by (k mod L) * (16 / L) slots inside every group of 16: the half-warp's 16 addresses fall into 16 different
// 8-byte bank pairs (measured before: 2.3 wavefronts per shared load, 4 per store).
template <int L>
__device__ __forceinline__ int k5s_at(int k, int s, int NC) {
  return k * NC + (s ^ ((k & (L - 1)) * (16 / L)));
}

template <int W, int L>
struct K5SView {
  static constexpr int ILP = W;
  const double* th;  // shared: the CTA's theta block
  const double* u;   // shared: the column of the CTA's current move
  double dl;
  int NC, s;
  __device__ __forceinline__ double operator[](int k) const { return fma(u[k], dl, th[k5s_at<L>(k, s, NC)]); }
};

// does the model evaluate C views in one sweep over its data (ssfunction_view_batch<C, V>)?
template <class M, class = void>
struct has_ssfunction_view_batch { static constexpr bool value = false; };
template <class M>
struct has_ssfunction_view_batch<M, decltype((void)&M::template ssfunction_view_batch<2, mcmcb_view_probe>)> {
  static constexpr bool value = true;
};

// bytes of dynamic shared memory: blob | two column buffers (npar + 1 doubles each: U(:,j) and qcovstd(j)) | theta
__host__ __device__ __forceinline__ size_t k5s_smem_bytes(int d, int chains, size_t blob_bytes) {
  const size_t dp2 = (size_t)(d + 2) & ~(size_t)1;
  return ((blob_bytes + 15) & ~(size_t)15) + sizeof(double) * (2 * dp2 + (size_t)d * chains);
}

template <int L>
__device__ __forceinline__ double k5s_lanes_sum(double v) {
#pragma unroll
  for (int off = 1; off < L; off <<= 1) v += __shfl_xor_sync(FULL, v, off);
  return v;
}

// default prior (priorfun.f90:97-100) of a view, split over the chain's lanes
template <class V>
__device__ __forceinline__ double k5s_default_prior_view(const V& theta, int len, const mcmcb_ctx& c) {
  if (c.prior == nullptr) return 0.0;
  double p = 0.0;
  for (int i = c.lane; i < len; i += c.nlanes) {
    const double sg = c.prior[len + i];
    if (sg > 0.0) {
      const double t = (theta[i] - c.prior[i]) / sg;
      p += t * t;
    }
  }
  return p;
}

// the scalars of one chain a thread carries through the launch
template <int NY>
struct K5SChain {
  double ss1[NY], s2[NY], pri1;
  int stayed, bnd, chainind, simuind, status, cnt, pend, nbuf;
  Rng g;
  long long cc;
  bool live, rejall, logged;
};

// W = accumulation chains per chain the model keeps in flight, L = lanes per chain, C = chains per thread;
// blockDim = L x (chains per CTA) / C
template <class M, int W, int L, int C>
__global__ void __launch_bounds__(K5S_CHAINS * L / C, 1) k5s_scam_step_kernel(const __grid_constant__ K2Params p) {
  constexpr int NY = M::NY;
  constexpr K2Layout Lo = k2_layout(NY);
  constexpr int NPF = (K4_DM * C + 64 * L - 1) / (64 * L);  // column elements a thread carries (blockDim >= 64 L / C)
  using View = K5SView<W, L>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  tma_stage_blob(smem_raw, p.blob, p.blob_bytes, &mbar);  // every thread of the CTA takes part (barrier inside)
  const double* data = reinterpret_cast<const double*>(smem_raw);
  const int d = p.d, T = blockDim.x, NT = T / L, NC = NT * C, tid = threadIdx.x, grp = tid / L, sub = tid % L;
  const int dp2 = (d + 2) & ~1;
  double* ucol = reinterpret_cast<double*>(smem_raw + (((size_t)p.blob_bytes + 15) & ~(size_t)15));
  double* th = ucol + 2 * (size_t)dp2;
  const DevCfg& c = p.c;
  const size_t P = (size_t)p.pitch;
  const double* U = p.Rm;    // shared rotation, column-major: column j at U + j d
  const double* gq = p.qstd; // shared qcovstd

  mcmcb_ctx ctx;
  ctx.data = data; ctx.ndata = p.blob_n; ctx.prior = p.prior; ctx.lane = sub; ctx.nlanes = L;
  ctx.exp_tl = 0u; ctx.exp_c1 = MCMCB_EXP_C1L; ctx.exp_c2 = MCMCB_EXP_C2L; ctx.scratch = nullptr;

  // chain i of this thread sits in slot i * NT + grp.  A slot past the last chain shadows the last chain (barriers and
  // shuffles need every thread) and writes nothing.
  K5SChain<NY> ch[C];
  View tv[C];
#pragma unroll
  for (int i = 0; i < C; i++) {
    K5SChain<NY>& q = ch[i];
    const int s = i * NT + grp;
    const long long c0 = (long long)blockIdx.x * NC + s;
    q.live = c0 < p.nchains;
    q.cc = q.live ? c0 : p.nchains - 1;
    const double* st = p.st + q.cc;
    const int* ist = p.ist + q.cc;
    const double* gth = p.theta + q.cc * p.dp;
    for (int k = sub; k < d; k += L) th[k5s_at<L>(k, s, NC)] = gth[k];
#pragma unroll
    for (int k = 0; k < NY; k++) { q.ss1[k] = st[(Lo.ss + k) * P]; q.s2[k] = st[(Lo.s2 + k) * P]; }
    q.pri1 = st[Lo.pri * P];
    q.stayed = ist[Lo.i_stayed * P]; q.bnd = ist[Lo.i_bnd * P]; q.chainind = ist[Lo.i_chainind * P];
    q.simuind = ist[Lo.i_simuind * P]; q.status = ist[Lo.i_status * P]; q.cnt = ist[Lo.i_cnt * P];
    q.pend = ist[Lo.i_pend * P]; q.nbuf = ist[Lo.i_nbuf * P];
    Rng& g = q.g;  // the chain's L lanes run the same generator: same draws, same decisions
    g.nd = ((unsigned long long)(unsigned)ist[Lo.i_ndhi * P] << 32) | (unsigned)ist[Lo.i_ndlo * P];
    g.seed = p.seed; g.chain = (unsigned long long)(p.chain_offset + q.cc);
    g.inj = p.inj ? p.inj + (unsigned long long)q.cc * p.inj_per_chain : nullptr;
    g.inj_n = p.inj_per_chain;
    g.cache_valid = false; g.cache_lo = g.cache_hi = 0; g.cache_blk = 0;
    g.has_spare = ist[Lo.i_hasspare * P] != 0;
    g.spare = st[Lo.spare * P];
    g.exhausted = 0;
    tv[i].th = th; tv[i].u = ucol + dp2; tv[i].dl = 0.0; tv[i].NC = NC; tv[i].s = s;
  }

  // C views -> C sums of squares (summed over the chain's lanes) and C priors
  auto evaluate = [&](double (&ssn)[C][NY], double (&prn)[C]) {
#pragma unroll
    for (int i = 0; i < C; i++) prn[i] = k5s_lanes_sum<L>(k5s_default_prior_view(tv[i], d, ctx));
    if constexpr (C > 1 && has_ssfunction_view_batch<M>::value) {
      M::template ssfunction_view_batch<C>(tv, d, NY, ctx, &ssn[0][0]);
    } else {
#pragma unroll
      for (int i = 0; i < C; i++) M::ssfunction_view(tv[i], d, NY, ctx, ssn[i]);
    }
#pragma unroll
    for (int i = 0; i < C; i++)
#pragma unroll
      for (int k = 0; k < NY; k++) ssn[i][k] = k5s_lanes_sum<L>(ssn[i][k]);
  };

  // buffer 0 <- column 0 and its scale; buffer 1 <- zeros (the current point as a view: a zero move along a zero column)
  for (int k = tid; k < d; k += T) { ucol[k] = U[k]; ucol[dp2 + k] = 0.0; }
  if (tid == 0) ucol[d] = gq[0];
  __syncthreads();
  if (ch[0].simuind == 0) {  // MCMC_run_scam.F90:26-36: initial point, saved as row 1 (every chain of a launch starts together)
    double ssn[C][NY], prn[C];
    evaluate(ssn, prn);
#pragma unroll
    for (int i = 0; i < C; i++) {
      K5SChain<NY>& q = ch[i];
#pragma unroll
      for (int k = 0; k < NY; k++) q.ss1[k] = ssn[i][k];
      q.pri1 = prn[i];
      q.chainind = 1; q.simuind = 1; q.cnt = 1; q.pend = 1;
      if (q.live && q.cc < p.store_chains) {
        double* srow = p.store_rows_p + (size_t)q.cc * p.store_rows * (d + NY);
        for (int k = sub; k < d; k += L) srow[k] = th[k5s_at<L>(k, tv[i].s, NC)];
        if (sub == 0) {
#pragma unroll
          for (int k = 0; k < NY; k++) {
            srow[d + k] = q.ss1[k];
            if (c.updatesigma) p.store_s2_p[(size_t)q.cc * p.store_rows * NY + k] = q.s2[k];
          }
        }
      }
    }
  }
  __syncthreads();  // nobody reads the zero column any more: it becomes the buffer of move 1

  int jb = 0;
  for (int done = 0; done < p.nsteps; done++) {
#pragma unroll
    for (int i = 0; i < C; i++) { ch[i].rejall = true; ch[i].logged = false; }
    for (int j = 0; j < d; j++) {
      const double* uc = ucol + (size_t)jb * dp2;
      // ---- the next move's column and scale start their trip from L2 now and are published after this move
      double pf[NPF], qn = 0.0;
      {
        const int jn = j + 1 < d ? j + 1 : 0;
        const double* ncol = U + (size_t)jn * d;
#pragma unroll
        for (int i = 0; i < NPF; i++) { const int k = tid + T * i; pf[i] = k < d ? ncol[k] : 0.0; }
        if (tid == 0) qn = gq[jn];
      }
      // MCMC_propose_sc (MCMC_run_scam.F90:94-117) in the O(d) form theta + delta U(:,j)
      const double qj = uc[d];
#pragma unroll
      for (int i = 0; i < C; i++) { tv[i].u = uc; tv[i].dl = ch[i].g.normal() * qj; }
      double ssn[C][NY], prn[C];
      evaluate(ssn, prn);
#pragma unroll
      for (int i = 0; i < C; i++) {
        K5SChain<NY>& q = ch[i];
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < NY; k++) sum += (ssn[i][k] - q.ss1[k]) / q.s2[k];
        const bool reject = mh_reject(alpha_from_tst(-0.5 * (sum + (prn[i] - q.pri1))), q.g);
        if (!reject) {  // MCMC_run_scam.F90:63-68
          const int s = tv[i].s;
          if (!q.logged) {
            // first acceptance of this sweep: the current row is complete -- log it for the adaptation
            // kernel before theta changes (the reference reads it back from the stored chain)
            const bool absorbing = c.doadapt && !(c.adaptend > 0 && q.simuind + 1 > c.adaptend);
            if (absorbing) {
              if (q.nbuf < p.rowcap) {
                if (q.live) {
                  double* rb = p.rowbuf + ((size_t)q.cc * (p.rowcap + 1) + q.nbuf) * (d + 1);
                  for (int k = sub; k < d; k += L) rb[k] = th[k5s_at<L>(k, s, NC)];
                  if (sub == 0) rb[d] = (double)((c.doadapt && c.adapthist > 1) ? q.cnt : q.pend);
                }
                q.nbuf++;
              } else {
                q.status |= MCMCB_ST_STORE_FULL;
              }
            }
            q.logged = true;
          }
          const double delta = tv[i].dl;
          for (int k = sub; k < d; k += L) { const int e = k5s_at<L>(k, s, NC); th[e] = fma(uc[k], delta, th[e]); }
#pragma unroll
          for (int k = 0; k < NY; k++) q.ss1[k] = ssn[i][k];
          q.pri1 = prn[i];
          q.rejall = false;
        }
      }
      // ---- publish the next column (its buffer was last read during the previous move) and meet: theta elements
      // written by one lane of a chain are read by the others in the next move
      {
        double* un = ucol + (size_t)(jb ^ 1) * dp2;
#pragma unroll
        for (int i = 0; i < NPF; i++) { const int k = tid + T * i; if (k < d) un[k] = pf[i]; }
        if (tid == 0) un[d] = qn;
      }
      __syncthreads();
      jb ^= 1;
    }
    // ---------------- end of sweep, MCMC_run_scam.F90:74-86
#pragma unroll
    for (int i = 0; i < C; i++) {
      K5SChain<NY>& q = ch[i];
      const bool stored = q.live && q.cc < p.store_chains;
      double* srow = p.store_rows_p + (size_t)q.cc * p.store_rows * (d + NY);
      double* scnt = p.store_cnt_p + (size_t)q.cc * p.store_rows;
      double* ss2st = p.store_s2_p + (size_t)q.cc * p.store_rows * NY;
      const int it = q.simuind + 1;
      q.simuind = it;
      if (q.rejall) {
        q.stayed++;
        q.cnt++; q.pend++;
      } else {
        if (stored && sub == 0 && q.chainind - 1 < p.store_rows) scnt[q.chainind - 1] = (double)q.cnt;
        q.chainind++;
        q.cnt = 1; q.pend = 1;
      }
      if (c.updatesigma) {
#pragma unroll
        for (int k = 0; k < NY; k++) {
          const double gg = q.g.gamma(c.N0 / 2.0 + (double)p.nobs[k] / 2.0, 2.0 / (c.N0 * c.S02 + q.ss1[k]));
          q.s2[k] = 1.0 / gg;
        }
      }
      if (stored) {
        if (!q.rejall) {
          if (q.chainind - 1 < p.store_rows) {
            for (int k = sub; k < d; k += L) srow[(size_t)(q.chainind - 1) * (d + NY) + k] = th[k5s_at<L>(k, tv[i].s, NC)];
            if (sub == 0) {
#pragma unroll
              for (int k = 0; k < NY; k++) srow[(size_t)(q.chainind - 1) * (d + NY) + d + k] = q.ss1[k];
            }
          } else {
            q.status |= MCMCB_ST_STORE_FULL;
          }
        }
        if (c.updatesigma && sub == 0 && it - 1 < p.store_rows) {
#pragma unroll
          for (int k = 0; k < NY; k++) ss2st[(size_t)(it - 1) * NY + k] = q.s2[k];
        }
      }
      if (q.g.exhausted) q.status |= MCMCB_ST_RNG_EXHAUSTED;
    }
  }

  // ---- write state back
#pragma unroll
  for (int i = 0; i < C; i++) {
    K5SChain<NY>& q = ch[i];
    if (!q.live) continue;
    double* gth = p.theta + q.cc * p.dp;
    for (int k = sub; k < d; k += L) gth[k] = th[k5s_at<L>(k, tv[i].s, NC)];
    if (sub != 0) continue;
    double* st = p.st + q.cc;
    int* ist = p.ist + q.cc;
#pragma unroll
    for (int k = 0; k < NY; k++) { st[(Lo.ss + k) * P] = q.ss1[k]; st[(Lo.s2 + k) * P] = q.s2[k]; }
    st[Lo.pri * P] = q.pri1; st[Lo.spare * P] = q.g.spare;
    ist[Lo.i_stayed * P] = q.stayed; ist[Lo.i_bnd * P] = q.bnd;
    ist[Lo.i_chainind * P] = q.chainind; ist[Lo.i_simuind * P] = q.simuind; ist[Lo.i_status * P] = q.status;
    ist[Lo.i_hasspare * P] = q.g.has_spare ? 1 : 0;
    ist[Lo.i_cnt * P] = q.cnt; ist[Lo.i_pend * P] = q.pend; ist[Lo.i_nbuf * P] = q.nbuf;
    ist[Lo.i_ndlo * P] = (int)(unsigned)(q.g.nd & 0xffffffffull);
    ist[Lo.i_ndhi * P] = (int)(unsigned)(q.g.nd >> 32);
  }
}

}  // namespace mcmcb
