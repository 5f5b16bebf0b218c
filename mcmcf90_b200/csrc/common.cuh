// Shared device-side pieces of the batched MH kernels: configuration constants,
// counter-based RNG with the reference's variate transforms, acceptance rules.
// Reference citations are file:line into mjlaine/mcmcf90.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mcmcb200.h"
#include "mcmcb200_model.cuh"

namespace mcmcb {

// kernel-relevant values of namelist &mcmc (mcmcinit.F90:74-82) after
// check_mcmcinit_parameters (mcmcinit.F90:235-368)
struct DevCfg {
  int method, nsimu;
  int doadapt, adaptint, adapthist, adaptend, initcmatn;
  int doburnin, burnintime, badaptint, greedy;
  int updatesigma, dodr, doscam, usesvd;
  int pool;  // pooled cross-chain adaptation: factors are written by the pool kernels, not at the chain's own ticks
  double scalelimit, scalefactor, drscale, condmax;
  double N0, S02;
  double alphatarget, nuparam;
};

constexpr unsigned FULL = 0xffffffffu;
// log(tiny(0d0)), mcmcprec.F90:34-41
constexpr double LOG_REALMIN = -708.3964185322641;
constexpr double DBL_HUGE = 1.7976931348623157e308;  // huge(0d0), MCMC_run.F90:50

// ----------------------------------------------------------------------------- Philox
// Philox4x32-10 (Salmon et al. 2011); replaces the Fortran runtime's random_number
// (mcmcrand.F90:55,104,138,156,177; MCMC_DRAM.F90:151).  Stream = (seed, global chain
// id); uniform number k lives in block k>>1, half k&1, as a 53-bit value in [0,1).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
  unsigned long long nd;  // uniforms consumed so far by this chain
  unsigned long long seed, chain;
  const double* inj;      // injected stream of this chain (nullptr => Philox)
  unsigned long long inj_n;
  uint32_t cache_lo, cache_hi;
  unsigned long long cache_blk;  // Philox block the cached upper half belongs to
  bool cache_valid;
  bool has_spare;  // polar spare, mcmcrand.F90:172-173
  double spare;
  int exhausted;

  // uniform number k of this chain's stream, without touching the stream position
  __device__ __forceinline__ double uniform_at(unsigned long long k) {
    if (inj != nullptr) {
      if (k < inj_n) return inj[k];
      exhausted = 1;
      return 0.5;
    }
    const unsigned long long blk = k >> 1;
    uint32_t o[4];
    philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)chain, (uint32_t)(chain >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), o);
    const unsigned long long bits = (k & 1ull) ? (((unsigned long long)o[3] << 32) | o[2])
                                               : (((unsigned long long)o[1] << 32) | o[0]);
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
  }

  // uniforms k and k+1 of the stream: one Philox block when k is even (both halves of block k/2)
  __device__ __forceinline__ void uniform_pair_at(unsigned long long k, double& u1, double& u2) {
    if (inj != nullptr || (k & 1ull)) {
      u1 = uniform_at(k);
      u2 = uniform_at(k + 1ull);
      return;
    }
    const unsigned long long blk = k >> 1;
    uint32_t o[4];
    philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)chain, (uint32_t)(chain >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), o);
    u1 = (double)((((unsigned long long)o[1] << 32) | o[0]) >> 11) * (1.0 / 9007199254740992.0);
    u2 = (double)((((unsigned long long)o[3] << 32) | o[2]) >> 11) * (1.0 / 9007199254740992.0);
  }

  __device__ __forceinline__ double uniform() {
    if (inj != nullptr) {
      double u = 0.5;
      if (nd < inj_n) u = inj[nd]; else exhausted = 1;
      nd++;
      return u;
    }
    unsigned long long k = nd++;
    uint32_t lo, hi;
    if ((k & 1ull) && cache_valid && cache_blk == (k >> 1)) {
      lo = cache_lo; hi = cache_hi;
      cache_valid = false;
    } else {
      unsigned long long blk = k >> 1;
      uint32_t o[4];
      philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)chain, (uint32_t)(chain >> 32), (uint32_t)seed,
                    (uint32_t)(seed >> 32), o);
      if (k & 1ull) {
        lo = o[2]; hi = o[3];
      } else {
        lo = o[0]; hi = o[1];
        cache_lo = o[2]; cache_hi = o[3];
        cache_blk = blk;
        cache_valid = true;
      }
    }
    unsigned long long bits = ((unsigned long long)hi << 32) | lo;
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
  }

  // mcmcrand.F90:166-190 normal_bm: Marsaglia polar; z*x(2) is returned first, z*x(1) saved
  __device__ __forceinline__ double normal() {
    if (has_spare) {
      has_spare = false;
      return spare;
    }
    double x1, x2, xx;
    for (;;) {
      x1 = 2.0 * uniform() - 1.0;
      x2 = 2.0 * uniform() - 1.0;
      xx = x1 * x1 + x2 * x2;
      if ((xx < 1.0 && xx != 0.0) || exhausted) break;
    }
    double z = sqrt(-2.0 * log(xx) / xx);
    spare = z * x1;
    has_spare = true;
    return z * x2;
  }

  // mcmcrand.F90:120-162 gammar_mt (Marsaglia & Tsang 2000), a >= 1
  __device__ __forceinline__ double gamma_mt(double a, double b) {
    double d = a - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    double x, v, u;
    for (;;) {
      for (;;) {
        x = normal();
        v = 1.0 + c * x;
        if (v > 0.0 || exhausted) break;
      }
      if (exhausted) { v = 1.0; break; }
      v = v * v * v;
      u = uniform();
      double x2 = x * x;
      if (u < 1.0 - 0.0331 * (x2 * x2)) break;
      if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) break;
    }
    return b * d * v;
  }

  // mcmcrand.F90:86-111 random_gamma(1,a,b)
  __device__ __forceinline__ double gamma(double a, double b) {
    if (a < 1.0) {
      double u = uniform();
      return gamma_mt(1.0 + a, b) * pow(u, 1.0 / a);
    }
    return gamma_mt(a, b);
  }
};

// exp() whose subnormal results are (up to a 1e-13-wide window) the correctly rounded ones, as
// host libm's are.  CUDA's exp() is 1 ulp, which at the bottom of the subnormal range means it can
// return 0 where libm returns the smallest subnormal -- and "draw a uniform iff alpha > 0"
// (MCMC_DRAM.F90:147-153) then consumes a different number of draws.  alpha13 has no underflow
// clamp in the reference (MCMC_DRAM.F90:184), so this range is reachable.
__device__ __forceinline__ double exp_subnormal_safe(double x) {
  if (x < -708.0) return exp(x + 69.314718055994530942 /* 100 ln 2 */) * 7.8886090522101180541e-31 /* 2^-100 */;
  return exp(x);
}

// MCMC_DRAM.F90:100-118 (tst is the caller's -0.5*(sum((ss2-ss1)/sigma2) + (sspri2-sspri1)))
__device__ __forceinline__ double alpha_from_tst(double tst) {
  if (tst >= 0.0) return 1.0;
  if (tst < LOG_REALMIN) return 0.0;
  return exp(tst);
}

// MCMC_DRAM.F90:140-155: a uniform is consumed only when 0 < alpha < 1; NaN rejects
__device__ __forceinline__ bool mh_reject(double alpha, Rng& g) {
  bool reject = true;
  if (alpha >= 1.0) {
    reject = false;
  } else if (alpha > 0.0) {
    double u = g.uniform();
    if (u <= alpha) reject = false;
  }
  return reject;
}

// classic reference-BLAS drotg, as called by dchud.f:138
__device__ __forceinline__ void drotg(double& da, double& db, double& c, double& s) {
  double roe = db, r, z;
  if (fabs(da) > fabs(db)) roe = da;
  double scale = fabs(da) + fabs(db);
  if (scale == 0.0) {
    c = 1.0; s = 0.0; r = 0.0; z = 0.0;
  } else {
    double ta = da / scale, tb = db / scale;
    r = scale * sqrt(ta * ta + tb * tb);
    r = (roe < 0.0) ? -r : r;
    c = da / r;
    s = db / r;
    z = 1.0;
    if (fabs(da) > fabs(db)) z = s;
    if (fabs(db) >= fabs(da) && c != 0.0) z = 1.0 / c;
  }
  da = r;
  db = z;
}

// ----------------------------------------------------------------- TMA bulk staging
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 8 bytes global -> shared without a register in between (LDGSTS); pair with cp.async.commit_group / wait_group
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// One thread stages `bytes` (multiple of 16) from global to shared with cp.async.bulk
// (TMA, SASS UBLKCP) completing on an mbarrier; every thread then waits on the barrier.
__device__ __forceinline__ void tma_stage_blob(void* smem_dst, const void* gsrc, uint32_t bytes,
                                               unsigned long long* mbar) {
  const uint32_t bar = smem_u32(mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    uint32_t done = 0;
    const uint32_t CH = 65536;  // chunk the copy; each chunk completes tx bytes on the same barrier
    while (done < bytes) {
      uint32_t n = bytes - done < CH ? bytes - done : CH;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32((const char*)smem_dst + done)),
          "l"((const char*)gsrc + done), "r"(n), "r"(bar)
          : "memory");
      done += n;
    }
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(0u)
        : "memory");
  }
}

}  // namespace mcmcb
