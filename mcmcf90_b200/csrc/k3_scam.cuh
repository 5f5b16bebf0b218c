// K3: SVD-factor paths on top of the warp-per-chain infrastructure of K2.
//
//  * SCAM, single-component adaptive Metropolis (MCMC_run_scam.F90:12-138): one MH test per
//    component per step in the basis of the eigenvectors U of the chain covariance, with
//    proposal standard deviations qcovstd = sqrt(eigenvalues) (scam_svd, matutils.F90:583-653;
//    MCMC_calculate_R, MCMC_adapt.F90:189-200: no 2.4/sqrt(d) scaling).
//  * usesvd DRAM/AM (condmax > 0): R = U diag(sqrt(s)) * 2.4/sqrt(d) with the condmax floor
//    (covtor_svd, matutils.F90:378-453; MCMC_adapt.F90:204-216), proposal theta + R z (dgemv 'N',
//    MCMC_DRAM.F90:27).
//
// Storage: the factor of a chain is d x d COLUMN-major in the same HBM buffer K2 uses for its
// row-major Cholesky factor (Rm), so column j -- the only part of U a SCAM component move needs --
// is one coalesced read of d doubles.
//
// What differs from the reference's arithmetic: the reference forms a component proposal as
// U (U' theta + delta e_j) with two dgemv per component (MCMC_run_scam.F90:122-138), 4 d^2 flops
// and 2 d^2 doubles of factor traffic; here it is theta + delta U(:,j), the same point when U is
// orthogonal (rounding-level difference, d^2 times less work; the reference's own unused
// MCMC_scam_update, MCMC_run_scam.F90:142-153, is this form).  dgesvd is replaced by the cyclic
// Jacobi sweep order of the oracle (oracle/mcmc_oracle.c orc_symeig), eigenvalues sorted
// descending, column signs fixed by the largest component; singular vectors of a (nearly)
// degenerate covariance are not unique, parity there is distributional only.
#pragma once
#include "k2_large.cuh"

namespace mcmcb {

constexpr int FACTOR_CHOL = 0, FACTOR_SVD = 1, FACTOR_SCAM = 2;

__device__ __forceinline__ double cta_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
  return s;
}

// Eigen-decomposition of the symmetric PSD matrix cm (d x d, full) by the whole CTA (blockDim >= d):
// cyclic two-sided Jacobi in the oracle's (p,q) order; A = scratch d*d, U = output d*d column-major,
// sv = shared d doubles (eigenvalues, descending), perm = shared d ints.  Returns 0 / 1 (not converged).
__device__ __forceinline__ int cta_symeig(const double* cm, double* A, double* U, double* sv, int* perm, int d,
                                          double* red) {
  const int tid = threadIdx.x;
  for (int k = tid; k < d * d; k += blockDim.x) {
    const int j = k / d, i = k - j * d;  // column-major (i,j)
    U[k] = (i == j) ? 1.0 : 0.0;
    A[k] = (i > j) ? cm[(size_t)j * d + i] : cm[(size_t)i * d + j];  // cm is symmetric; upper triangle authoritative
  }
  __syncthreads();
  int conv = 0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0.0, dia = 0.0;
    for (int k = tid; k < d * d; k += blockDim.x) {
      const int j = k / d, i = k - j * d;
      const double a = A[k];
      if (i == j) dia = fma(a, a, dia);
      else if (i < j) off = fma(a, a, off);
    }
    off = cta_sum(off, red);
    dia = cta_sum(dia, red);
    if (off <= 1e-32 * dia || off == 0.0) { conv = 1; break; }
    for (int p = 0; p < d - 1; p++)
      for (int q = p + 1; q < d; q++) {
        const double apq = A[(size_t)q * d + p];
        const double app = A[(size_t)p * d + p], aqq = A[(size_t)q * d + q];
        __syncthreads();  // everyone has read the pivot block before its owners overwrite it
        if (apq == 0.0) continue;
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        if (tid < d) {  // A <- A J and U <- U J, row k = tid
          const double akp = A[(size_t)p * d + tid], akq = A[(size_t)q * d + tid];
          A[(size_t)p * d + tid] = c * akp - sn * akq;
          A[(size_t)q * d + tid] = sn * akp + c * akq;
          const double ukp = U[(size_t)p * d + tid], ukq = U[(size_t)q * d + tid];
          U[(size_t)p * d + tid] = c * ukp - sn * ukq;
          U[(size_t)q * d + tid] = sn * ukp + c * ukq;
        }
        __syncthreads();
        if (tid < d) {  // A <- J' A, column k = tid
          const double apk = A[(size_t)tid * d + p], aqk = A[(size_t)tid * d + q];
          A[(size_t)tid * d + p] = c * apk - sn * aqk;
          A[(size_t)tid * d + q] = sn * apk + c * aqk;
        }
        __syncthreads();
      }
  }
  for (int j = tid; j < d; j += blockDim.x) sv[j] = fabs(A[(size_t)j * d + j]);
  __syncthreads();
  // selection sort, descending, first maximum wins (stable for ties): serial on the d values, the
  // column moves are applied afterwards through the permutation
  if (tid == 0) {
    for (int j = 0; j < d; j++) perm[j] = j;
    for (int a = 0; a < d - 1; a++) {
      int best = a;
      for (int b = a + 1; b < d; b++)
        if (sv[b] > sv[best]) best = b;
      if (best != a) {
        const double ts = sv[a]; sv[a] = sv[best]; sv[best] = ts;
        const int tp = perm[a]; perm[a] = perm[best]; perm[best] = tp;
      }
    }
  }
  __syncthreads();
  // permuted copy through the scratch matrix, then the sign convention per column
  for (int k = tid; k < d * d; k += blockDim.x) {
    const int j = k / d, i = k - j * d;
    A[k] = U[(size_t)perm[j] * d + i];
  }
  __syncthreads();
  for (int j = tid; j < d; j += blockDim.x) {
    int im = 0;
    for (int i = 1; i < d; i++)
      if (fabs(A[(size_t)j * d + i]) > fabs(A[(size_t)j * d + im])) im = i;
    const double sg = (A[(size_t)j * d + im] < 0.0) ? -1.0 : 1.0;
    for (int i = 0; i < d; i++) U[(size_t)j * d + i] = sg * A[(size_t)j * d + i];
  }
  __syncthreads();
  return conv ? 0 : 1;
}

// MCMC_calculate_R for the SVD factor modes (MCMC_adapt.F90:189-216 with matutils.F90:378-453 /
// 583-653).  cm may be rewritten (cmat := R0 R0' when singular values were floored).  Returns the
// status bits to OR into the chain's status word.  tmpA, tmpU = 2 scratch matrices.
__device__ __forceinline__ int cta_calculate_R_svd(double* cm, double* Rm, double* qstd, double* tmpA, double* tmpU,
                                                   int d, int mode, double condmax, double* sv, int* perm,
                                                   double* red) {
  int st = 0;
  int info = cta_symeig(cm, tmpA, tmpU, sv, perm, d, red);  // 0, or 1 = not converged
  if (info) st |= MCMCB_ST_SVDFAIL;
  const bool zero = (sv[0] == 0.0);  // matutils.F90:418-421 / 624-627: info = n, nothing else computed
  const double tol = sv[0] / condmax;
  const bool floored = !zero && (sv[d - 1] <= tol);
  __syncthreads();
  if (floored) {
    for (int i = threadIdx.x; i < d; i += blockDim.x)
      if (sv[i] < tol) sv[i] = tol;
    info = 0;  // info = -1 in the reference, reset to 0 by both callers (MCMC_adapt.F90:195,205-208)
  }
  __syncthreads();
  if (mode == FACTOR_SCAM) {
    if (!zero)
      for (int i = threadIdx.x; i < d; i += blockDim.x) qstd[i] = sqrt(sv[i]);
    for (int k = threadIdx.x; k < d * d; k += blockDim.x) Rm[k] = tmpU[k];  // R = R0 either way (MCMC_adapt.F90:197)
    __syncthreads();
    return st;
  }
  if (zero || info) return st | MCMCB_ST_CHOLFAIL;  // info /= 0: warn, keep the old R (MCMC_adapt.F90:169-171,213-214)
  for (int k = threadIdx.x; k < d * d; k += blockDim.x) {
    const int j = k / d;
    tmpU[k] = sqrt(sv[j]) * tmpU[k];  // dscal(n, sqrt(s(i)), R(1:n,i), 1), matutils.F90:442
  }
  __syncthreads();
  if (floored) {  // cmat = matmul(R0, transpose(R0)), MCMC_adapt.F90:205-208
    for (int k = threadIdx.x; k < d * d; k += blockDim.x) {
      const int j = k / d, i = k - j * d;
      double acc = 0.0;
      for (int m = 0; m < d; m++) acc = acc + tmpU[(size_t)m * d + i] * tmpU[(size_t)m * d + j];
      cm[k] = acc;
    }
  }
  const double sq = sqrt((double)d);
  for (int k = threadIdx.x; k < d * d; k += blockDim.x) Rm[k] = tmpU[k] * 2.4 / sq;  // MCMC_adapt.F90:216
  __syncthreads();
  return st;
}

// Initial factor (MCMC_init.F90:108-110) for the SVD modes; one CTA per chain.
static __global__ void k3_initR_kernel(K2Params p, double* scratch, int mode) {
  extern __shared__ double sh[];  // d doubles + d ints
  __shared__ double red[K2_ADAPT_THREADS / 32];
  constexpr K2Layout Lo = k2_layout(1);
  const long long c = blockIdx.x;
  const int d = p.d;
  double* sv = sh;
  int* perm = reinterpret_cast<int*>(sh + d);
  const int st = cta_calculate_R_svd(p.cmat + (size_t)c * d * d, p.Rm + (size_t)c * p.r_stride, p.qstd + c * p.q_stride,
                                     scratch + (size_t)c * 2 * d * d, scratch + (size_t)c * 2 * d * d + (size_t)d * d, d,
                                     mode, p.c.condmax, sv, perm, red);
  if (st && threadIdx.x == 0) p.ist[Lo.i_status * p.pitch + c] |= st;
}

// MCMC_adapt.F90:12-174 at step index p.tick_i for the SVD factor modes, one CTA per chain.
static __global__ void k3_adapt_kernel(K2Params p, double* scratch, int mode) {
  extern __shared__ double sh[];  // d doubles + d ints (as d doubles) + absorb_smem_doubles(d)
  __shared__ double red[K2_ADAPT_THREADS / 32];
  constexpr K2Layout Lo = k2_layout(1);
  const long long c = blockIdx.x;
  const DevCfg& cf = p.c;
  const int d = p.d, i = p.tick_i;
  double* sv = sh;
  int* perm = reinterpret_cast<int*>(sh + d);
  double* dvec = sh + 2 * d;
  double* st = p.st + c;
  int* ist = p.ist + c;
  double* cm = p.cmat + (size_t)c * d * d;
  double* Rm = p.Rm + (size_t)c * p.r_stride;
  double* mean = p.mean + c * p.dp;
  double* theta = p.theta + c * p.dp;
  double* rb = p.rowbuf + (size_t)c * (p.rowcap + 1) * (d + 1);
  double* tmpA = scratch + (size_t)c * 2 * d * d;
  double* tmpU = tmpA + (size_t)d * d;
  const int ma = cf.adaptint > 0 ? i % cf.adaptint : 1;
  const int mb = cf.badaptint > 0 ? i % cf.badaptint : 1;
  if (ma != 0 && mb != 0) return;
  double wsum = st[Lo.wsum * p.pitch];
  const int nbuf = ist[Lo.i_nbuf * p.pitch];
  int status = 0;
  if (i < cf.burnintime && cf.doburnin && mb == 0) {  // MCMC_adapt.F90:60-102 (never for SCAM: doburnin forced off)
    const double staypc = (double)ist[Lo.i_stayed * p.pitch] / (double)i;
    if (staypc > 1.0 - cf.scalelimit) {
      for (int k = threadIdx.x; k < d * d; k += blockDim.x) Rm[k] = Rm[k] / cf.scalefactor;
    } else if (staypc < cf.scalelimit) {
      for (int k = threadIdx.x; k < d * d; k += blockDim.x) Rm[k] = Rm[k] * cf.scalefactor;
    } else {
      for (int k = threadIdx.x; k < d * d; k += blockDim.x) cm[k] = p.cmat0[k];
      for (int k = threadIdx.x; k < d; k += blockDim.x) mean[k] = p.par0[c * d + k];
      __syncthreads();
      if (threadIdx.x == 0) {
        st[Lo.wsum * p.pitch] = (double)cf.initcmatn;
        ist[Lo.i_pend * p.pitch] = ist[Lo.i_cnt * p.pitch];
        ist[Lo.i_nbuf * p.pitch] = 0;
      }
      status = cta_calculate_R_svd(cm, Rm, p.qstd + c * p.q_stride, tmpA, tmpU, d, mode, cf.condmax, sv, perm, red);
    }
  } else if (i >= cf.burnintime + cf.adaptint + cf.adapthist && cf.doadapt && cf.adapthist > 1) {  // AP, MCMC_adapt.F90:116-136
    cta_ap_window(rb, nbuf, theta, cm, mean, st, ist, p.pitch, d, cf.adapthist);
    if (!cf.pool) status = cta_calculate_R_svd(cm, Rm, p.qstd + c * p.q_stride, tmpA, tmpU, d, mode, cf.condmax, sv, perm, red);
  } else if (i >= cf.burnintime + cf.adaptint + cf.adapthist && cf.doadapt) {  // MCMC_adapt.F90:105-159
    if (!p.absorbed) {
      for (int k = threadIdx.x; k < d; k += blockDim.x) rb[(size_t)nbuf * (d + 1) + k] = theta[k];
      if (threadIdx.x == 0) rb[(size_t)nbuf * (d + 1) + d] = (double)ist[Lo.i_pend * p.pitch];
      __syncthreads();
      cta_absorb_rows(rb, nbuf + 1, cm, mean, wsum, d, p.coef + (size_t)blockIdx.x * 2 * (p.rowcap + 1), dvec);
      if (threadIdx.x == 0) {
        st[Lo.wsum * p.pitch] = wsum;
        ist[Lo.i_pend * p.pitch] = 0;
        ist[Lo.i_nbuf * p.pitch] = 0;
      }
    }
    if (!cf.pool) status = cta_calculate_R_svd(cm, Rm, p.qstd + c * p.q_stride, tmpA, tmpU, d, mode, cf.condmax, sv, perm, red);
  }
  if (status && threadIdx.x == 0) ist[Lo.i_status * p.pitch] |= status;
}

// SCAM step kernel: one warp per chain; a step = one sweep over the d components
// (MCMC_run_scam.F90:38-88).
template <class M, bool SMEM>
__global__ void __launch_bounds__(K2_MAX_THREADS, 1) k3_scam_step_kernel(const __grid_constant__ K2Params p) {
  constexpr int NY = M::NY;
  constexpr K2Layout Lo = k2_layout(NY);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  const int d = p.d, dp = p.dp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const DevCfg& c = p.c;

  double* vecs = reinterpret_cast<double*>(smem_raw) + (size_t)warp * K2_NVEC * dp;
  double *th = vecs, *prop = vecs + dp, *zs = vecs + 2 * dp, *qs = vecs + 3 * dp;
  const double* data = p.blob;
  if (SMEM) {
    unsigned char* blob_s = smem_raw + sizeof(double) * (size_t)(blockDim.x >> 5) * K2_NVEC * dp;
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
    const double* U = p.Rm + (size_t)cc * p.r_stride;
    double* gth = p.theta + cc * dp;
    const double* gq = p.qstd + cc * p.q_stride;
    double* rb = p.rowbuf + (size_t)cc * (p.rowcap + 1) * (d + 1);

    for (int k = lane; k < dp; k += 32) { th[k] = gth[k]; prop[k] = gth[k]; qs[k] = gq[k]; }
    double ss1[NY], s2[NY];
#pragma unroll
    for (int k = 0; k < NY; k++) { ss1[k] = st[(Lo.ss + k) * p.pitch]; s2[k] = st[(Lo.s2 + k) * p.pitch]; }
    double pri1 = st[Lo.pri * p.pitch];
    int stayed = ist[Lo.i_stayed * p.pitch], bnd = ist[Lo.i_bnd * p.pitch];
    int chainind = ist[Lo.i_chainind * p.pitch], simuind = ist[Lo.i_simuind * p.pitch], status = ist[Lo.i_status * p.pitch];
    int cnt = ist[Lo.i_cnt * p.pitch], pend = ist[Lo.i_pend * p.pitch], nbuf = ist[Lo.i_nbuf * p.pitch];
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

    if (simuind == 0) {  // MCMC_run_scam.F90:26-36: initial point, saved as row 1
      double ssn[NY];
      M::ssfunction(th, d, NY, ctx, ssn);
#pragma unroll
      for (int k = 0; k < NY; k++) ss1[k] = warp_sum(ssn[k]);
      pri1 = M::priorfun(th, d, ctx);
      chainind = 1; simuind = 1; cnt = 1; pend = 1;
      if (stored) {
        for (int k = lane; k < d; k += 32) srow[k] = th[k];
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < NY; k++) { srow[d + k] = ss1[k]; if (c.updatesigma) ss2st[k] = s2[k]; }
        }
      }
    }

    for (int done = 0; done < p.nsteps; done++) {
      bool rejall = true;
      bool logged = false;  // the row that is about to be replaced has been written to the row buffer
      for (int j = 0; j < d; j++) {
        // MCMC_propose_sc (MCMC_run_scam.F90:94-117) in the O(d) form theta + delta U(:,j)
        warp_normals(g, zs, 1, lane);
        const double delta = zs[0] * qs[j];
        const double* col = U + (size_t)j * d;
        for (int k = lane; k < d; k += 32) prop[k] = fma(col[k], delta, th[k]);
        {
          // the next component's column (the first one again after the last): requested now, it arrives while the model
          // is evaluated -- the read above was the kernel's largest stall (profiles/r01_summary.md P)
          const double* nxt = U + (size_t)((j + 1 < d) ? j + 1 : 0) * d;
          for (int k = lane * 16; k < d; k += 32 * 16) asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt + k));
        }
        __syncwarp();
        bool reject;
        double ssn[NY], prn = 0.0;
        if (!M::checkbounds(prop, d, ctx)) {
          bnd++;  // dodr is forced off for SCAM (mcmcinit.F90:328-330)
          reject = true;
        } else {
          prn = M::priorfun(prop, d, ctx);
          M::ssfunction(prop, d, NY, ctx, ssn);
#pragma unroll
          for (int k = 0; k < NY; k++) ssn[k] = warp_sum(ssn[k]);
          double sum = 0.0;
#pragma unroll
          for (int k = 0; k < NY; k++) sum += (ssn[k] - ss1[k]) / s2[k];
          reject = mh_reject(alpha_from_tst(-0.5 * (sum + (prn - pri1))), g);
        }
        if (!reject) {  // MCMC_run_scam.F90:63-68
          if (!logged) {
            // first acceptance of this sweep: the current row is complete -- log it for the adaptation
            // kernel before theta changes (the reference reads it back from the stored chain)
            const bool absorbing = c.doadapt && !(c.adaptend > 0 && simuind + 1 > c.adaptend);
            if (absorbing) {
              if (nbuf < p.rowcap) {
                for (int k = lane; k < d; k += 32) rb[(size_t)nbuf * (d + 1) + k] = th[k];
                if (lane == 0) rb[(size_t)nbuf * (d + 1) + d] = (double)((c.doadapt && c.adapthist > 1) ? cnt : pend);
                nbuf++;
              } else {
                status |= MCMCB_ST_STORE_FULL;
              }
            }
            logged = true;
          }
          __syncwarp();
          for (int k = lane; k < d; k += 32) th[k] = prop[k];
#pragma unroll
          for (int k = 0; k < NY; k++) ss1[k] = ssn[k];
          pri1 = prn;
          rejall = false;
          __syncwarp();
        }
      }
      // ---------------- end of sweep, MCMC_run_scam.F90:74-86
      const int i = simuind + 1;
      simuind = i;
      if (rejall) {
        stayed++;
        cnt++; pend++;
      } else {
        if (stored && chainind - 1 < p.store_rows && lane == 0) scnt[chainind - 1] = (double)cnt;
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
      if (g.exhausted) status |= MCMCB_ST_RNG_EXHAUSTED;
    }

    // ---- write state back
    for (int k = lane; k < dp; k += 32) gth[k] = th[k];
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NY; k++) { st[(Lo.ss + k) * p.pitch] = ss1[k]; st[(Lo.s2 + k) * p.pitch] = s2[k]; }
      st[Lo.pri * p.pitch] = pri1; st[Lo.spare * p.pitch] = g.spare;
      ist[Lo.i_stayed * p.pitch] = stayed; ist[Lo.i_bnd * p.pitch] = bnd;
      ist[Lo.i_chainind * p.pitch] = chainind; ist[Lo.i_simuind * p.pitch] = simuind; ist[Lo.i_status * p.pitch] = status;
      ist[Lo.i_hasspare * p.pitch] = g.has_spare ? 1 : 0;
      ist[Lo.i_cnt * p.pitch] = cnt; ist[Lo.i_pend * p.pitch] = pend; ist[Lo.i_nbuf * p.pitch] = nbuf;
      ist[Lo.i_ndlo * p.pitch] = (int)(unsigned)(g.nd & 0xffffffffull);
      ist[Lo.i_ndhi * p.pitch] = (int)(unsigned)(g.nd >> 32);
    }
    __syncwarp();
  }
}

}  // namespace mcmcb
