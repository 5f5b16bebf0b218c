// K2G: large-npar DRAM/AM (and early-rejection) sampler with a GROUP of warps per chain and the chain's
// Cholesky factor resident in shared memory.
//
// The warp-per-chain kernel (k2_large.cuh) re-reads a chain's factor from L2/HBM for every proposal -- at
// npar = 100 that is 40 KB per proposal through one warp, and the kernel ends up bound by instruction issue and
// load latency (profiles/r01_summary.md H).  Here GT = 32..256 threads own one chain:
//   * the upper factor is packed column by column (LAPACK 'U' packed, conflict-free: see k2g_col) into shared memory ONCE per launch
//     segment (it only changes at adaptation ticks, which are separate kernels) -- HBM sees d(d+1)/2 doubles per
//     chain per segment instead of per proposal;
//   * thread t owns column t (+GT, +2GT ..) of theta + R'z and of the model's matrix-vector product, so a
//     proposal is one dependent chain of <= npar DFMAs per thread instead of npar/32 of them per lane; every column
//     still accumulates its rows in ascending order == dtrmv('u','t') (matutils.F90:108-109), bit for bit;
//   * the group's warps meet at a named barrier (bar.sync id, GT); sums over the group are a warp shuffle
//     reduction followed by a fixed-order sum of the warps' partials in shared memory (deterministic);
//   * the chain's scalar state and its RNG position are replicated in every thread of the group (every thread
//     takes the same accept/reject decisions); the normals are produced by the group's first warp with the same
//     stream-order compaction as the warp kernel and the new stream position is published through shared memory.
// Several groups share a CTA (and the TMA-staged model blob).
//
// STATUS (round 1, profiles/r01_summary.md K): parity-green, but on the C2 shape (npar = 100, 4096 chains) it is on par
// with the warp-per-chain kernel, not faster -- 40 KB of factor per chain leaves room for 3 chains = 12 warps per SM,
// every one of them in a serial dependency chain (issue active 37 %; four interleaved partial sums per column changed
// nothing), so latency is hidden even less than with 16 warp-chains per SM.  It therefore runs only on request
// (MCMCB_K2_GROUP=1); the warp-per-chain kernel stays the default.  Reference map as k2_large.cuh: loop
// MCMC_run.F90:41-107 / MCMC_run_er.F90:50-83, primitives MCMC_DRAM.F90:20-206.
#pragma once
#include "k2_large.cuh"

namespace mcmcb {

constexpr int K2G_MAX_THREADS = 512;
constexpr int K2G_MAXM = 4;      // columns per thread: npar <= GT * K2G_MAXM
constexpr int K2G_RED = 8;       // warps per group at most

struct K2Group {
  int id, nthreads, gt, nwarps, flip;
  double* red;  // 2 * K2G_RED doubles
  __device__ __forceinline__ void bar() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
  // deterministic sum over the group's threads: shuffle tree inside each warp, then the warps' partials in order
  __device__ __forceinline__ double sum(double v) {
    v = warp_sum(v);
    double* r = red + (flip ? K2G_RED : 0);
    flip ^= 1;
    if ((gt & 31) == 0) r[gt >> 5] = v;
    bar();
    double s = r[0];
    for (int w = 1; w < nwarps; w++) s += r[w];
    return s;
  }
};

// The factor in shared memory: upper triangle packed by COLUMNS, R(i,j) at j(j+1)/2 + i (LAPACK 'U' packed).
// Thread j walks its own column with unit stride, so a row costs one load with an immediate offset and no address
// arithmetic; the starts of 16 consecutive columns are the triangular numbers T_j, which are pairwise distinct
// modulo 16 for any 16 consecutive j starting at a multiple of 16 (T_{16a+t} = T_16a + 16at + T_t, and t -> T_t
// mod 16 is the triangular-probing permutation), so the 8-byte loads of a half-warp hit 16 different bank pairs:
// conflict-free.
__device__ __forceinline__ int k2g_col(int j) { return (j * (j + 1)) >> 1; }

// acc[m] = sum_{i <= j} R(i,j) z_i for the columns j = gt + m*GT this thread owns, rows in ascending order
// (== dtrmv('u','t'), matutils.F90:108-109, bit for bit)
__device__ __forceinline__ void k2g_tri_matvec_t(const double* Rs, const double* zs, int d, int gt, int GT,
                                                 double (&acc)[K2G_MAXM]) {
#pragma unroll
  for (int m = 0; m < K2G_MAXM; m++) {
    acc[m] = 0.0;
    const int j = gt + m * GT;
    if (m * GT < d) {  // group-uniform
      const int jc = min(j, d - 1);  // threads past the last column walk a valid one; nobody reads their sum
      const double* col = Rs + k2g_col(jc);
      const int jlo = min(j & ~31, d - 1);  // first column of this warp: rows 0..jlo are inside every lane's column
      const int jhi = min(j | 31, d - 1);
      double a = 0.0;
      int i = 0;
      for (; i + 8 <= jlo + 1; i += 8) {
        double r[8];
#pragma unroll
        for (int u = 0; u < 8; u++) r[u] = col[i + u];
#pragma unroll
        for (int u = 0; u < 8; u++) a = fma(r[u], zs[i + u], a);
      }
      for (; i <= jlo; i++) a = fma(col[i], zs[i], a);
      // the 31-row triangle: a lane stops when the row passes its column (reads beyond the column stay inside Rs)
      for (; i <= jhi; i++) {
        const double r = col[min(i, jc)];
        if (i <= jc) a = fma(r, zs[i], a);
      }
      acc[m] = a;
    }
  }
}

template <class M, bool SMEM>
__global__ void __launch_bounds__(K2G_MAX_THREADS, 1) k2g_step_kernel(const __grid_constant__ K2Params p, int GT) {
  constexpr int NY = M::NY;
  constexpr K2Layout Lo = k2_layout(NY);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  const int d = p.d, dp = p.dp;
  const int ngroups = blockDim.x / GT;
  const int grp = threadIdx.x / GT, gt = threadIdx.x % GT, lane = threadIdx.x & 31;
  const DevCfg& c = p.c;
  const int T = (d * (d + 1)) >> 1, Tp = (T + 1) & ~1;

  // dynamic shared memory: per group [6 vectors of dp][packed factor Tp][2*K2G_RED reduction][4 publish], then the blob
  const size_t per_group = (size_t)K2_NVEC * dp + Tp + 2 * K2G_RED + 4;
  double* gbase = reinterpret_cast<double*>(smem_raw) + (size_t)grp * per_group;
  double *th = gbase, *prop = gbase + dp, *z1 = gbase + 2 * dp, *z2 = gbase + 3 * dp, *w2 = gbase + 5 * dp;
  double* Rs = gbase + (size_t)K2_NVEC * dp;
  double* pub = Rs + Tp + 2 * K2G_RED;  // [0] next chain, [1] nd, [2] spare, [3] has_spare + 2*exhausted
  K2Group G;
  G.id = 1 + grp; G.nthreads = GT; G.gt = gt; G.nwarps = GT >> 5; G.flip = 0; G.red = Rs + Tp;
  const double* data = p.blob;
  if (SMEM) {
    unsigned char* blob_s = smem_raw + sizeof(double) * (size_t)ngroups * per_group;
    tma_stage_blob(blob_s, p.blob, p.blob_bytes, &mbar);
    data = reinterpret_cast<const double*>(blob_s);
  }
  mcmcb_ctx ctx;
  ctx.data = data; ctx.ndata = p.blob_n; ctx.prior = p.prior; ctx.lane = gt; ctx.nlanes = GT;
  ctx.exp_tl = 0u; ctx.exp_c1 = MCMCB_EXP_C1L; ctx.exp_c2 = MCMCB_EXP_C2L;
  ctx.scratch = w2;
  ctx.bar_id = G.id; ctx.bar_threads = GT;

  for (;;) {
    if (gt == 0) pub[0] = (double)atomicAdd(p.tile_counter, 1u);
    G.bar();
    const long long cc = (long long)pub[0];
    G.bar();
    if (cc >= p.nchains) break;
    double* st = p.st + cc;
    int* ist = p.ist + cc;
    const double* Rg = p.Rm + (size_t)cc * p.r_stride;
    double* gth = p.theta + cc * dp;
    double* rb = p.rowbuf + (size_t)cc * (p.rowcap + 1) * (d + 1);

    // the factor, once per launch segment: row-major d x d upper in HBM -> packed columns in shared memory
    for (int e = gt; e < d * d; e += GT) {
      const int i = e / d, j = e - i * d;
      if (i <= j) Rs[k2g_col(j) + i] = Rg[e];
    }
    for (int k = gt; k < dp; k += GT) { th[k] = gth[k]; prop[k] = gth[k]; }
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
    G.bar();

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
        if (gt < 32) {  // the group's first warp draws the normals; everyone then takes over its stream position
          warp_normals(g, zs, d, lane);
          if (gt == 0) {
            pub[1] = __longlong_as_double((long long)g.nd);
            pub[2] = g.spare;
            pub[3] = (double)((g.has_spare ? 1 : 0) + (g.exhausted ? 2 : 0));
          }
        }
        G.bar();
        g.nd = (unsigned long long)__double_as_longlong(pub[1]);
        g.spare = pub[2];
        const int fl = (int)pub[3];
        g.has_spare = (fl & 1) != 0;
        if (fl & 2) g.exhausted = 1;
        g.cache_valid = false;
        double acc[K2G_MAXM];
        k2g_tri_matvec_t(Rs, zs, d, gt, GT, acc);
        const double sc = (phase == 0) ? 1.0 : 1.0 / c.drscale;
#pragma unroll
        for (int m = 0; m < K2G_MAXM; m++) {
          const int j = gt + m * GT;
          if (j < d) prop[j] = th[j] + (phase == 0 ? acc[m] : acc[m] * sc);
        }
        G.bar();
        inb = M::checkbounds(prop, d, ctx);
        if (c.method == MCMCB_ER && inb) {  // MCMC_sscrit, MCMC_DRAM.F90:124-135
          const double u = g.uniform();
          double sum = 0.0;
#pragma unroll
          for (int k = 0; k < NY; k++) sum += ss1[k] / s2[k];
          sscrit = -2.0 * log(u) + sum + pri1;
        }
      }
      // ---------------- user model (cooperative over the GT threads)
      double ssn[NY];
      M::ssfunction(prop, d, NY, ctx, ssn);
#pragma unroll
      for (int k = 0; k < NY; k++) ssn[k] = G.sum(ssn[k]);
      const double prn = M::priorfun(prop, d, ctx);
      // ---------------- accept / reject (group uniform)
      bool reject = false;
      if (phase < 0) {
#pragma unroll
        for (int k = 0; k < NY; k++) ss1[k] = ssn[k];
        pri1 = prn;
        chainind = 1; simuind = 1; cnt = 1; pend = 1;
        if (stored) {
          for (int k = gt; k < d; k += GT) srow[k] = th[k];
          if (gt == 0) {
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
          if (!c.dodr) bnd++;
#pragma unroll
          for (int k = 0; k < NY; k++) ssn[k] = DBL_HUGE;
          a12 = 0.0;
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
          for (int k = gt; k < d; k += GT) t = fma(z1[k], z1[k], t);
          z1sq = G.sum(t);
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
          for (int k = gt; k < d; k += GT) { const double v = z2[k] * sc - z1[k]; t = fma(v, v, t); }
          const double q1 = -0.5 * (G.sum(t) - z1sq);  // matrix-free DR ratio, see k2_large.cuh
          double a13 = exp_subnormal_safe(l2 + q1) * (1.0 - a32) / (1.0 - a12);
          if (a13 == a13) a13 = fmin(1.0, a13);
          reject = mh_reject(a13, g);
          if (!reject) dracc++;
        }
        phase = 0;
      }
      // ---------------- end of step, MCMC_run.F90:93-105
      const int i = simuind + 1;
      simuind = i;
      const bool absorbing = (c.doadapt && !(c.adaptend > 0 && i > c.adaptend)) || (c.greedy && c.doburnin && i <= c.burnintime);
      if (reject) {
        stayed++;
        cnt++; pend++;
      } else {
        if (absorbing) {  // log the completed row and its not-yet-counted weight for the adaptation kernel
          if (nbuf < p.rowcap) {
            for (int k = gt; k < d; k += GT) rb[(size_t)nbuf * (d + 1) + k] = th[k];
            if (gt == 0) rb[(size_t)nbuf * (d + 1) + d] = (double)((c.doadapt && c.adapthist > 1) ? cnt : pend);
            nbuf++;
          } else {
            status |= MCMCB_ST_STORE_FULL;
          }
        }
        if (stored && chainind - 1 < p.store_rows && gt == 0) scnt[chainind - 1] = (double)cnt;
        G.bar();  // every thread has read th (row log above, DR ratio) before it is overwritten
        for (int k = gt; k < d; k += GT) th[k] = prop[k];
#pragma unroll
        for (int k = 0; k < NY; k++) ss1[k] = ssn[k];
        pri1 = prn;
        chainind++;
        cnt = 1; pend = 1;
        G.bar();
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
            for (int k = gt; k < d; k += GT) srow[(size_t)(chainind - 1) * (d + NY) + k] = th[k];
            if (gt == 0) {
#pragma unroll
              for (int k = 0; k < NY; k++) srow[(size_t)(chainind - 1) * (d + NY) + d + k] = ss1[k];
            }
          } else {
            status |= MCMCB_ST_STORE_FULL;
          }
        }
        if (c.updatesigma && i - 1 < p.store_rows && gt == 0) {
#pragma unroll
          for (int k = 0; k < NY; k++) ss2st[(size_t)(i - 1) * NY + k] = s2[k];
        }
      }
      if (g.exhausted) status |= MCMCB_ST_RNG_EXHAUSTED;
      done++;
    }

    // ---- write state back
    G.bar();
    for (int k = gt; k < dp; k += GT) gth[k] = th[k];
    if (gt == 0) {
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
    G.bar();
  }
}

}  // namespace mcmcb
