// Built-in user models (the plugin side of external_inc.h:14-28), written against
// include/mcmcb200_model.cuh.  Blob layouts are shared with the host helpers in
// mcmcf90_b200/models.py.
#pragma once
#include "mcmcb200_model.cuh"

namespace mcmcb {

// Exponential-decay regression y = theta1*exp(-theta2*x) (testcases/mcmcrun.F90:89,104),
// bounds theta > 0 (testcases/mcmcrun.F90:112-122).  BASELINE configs C1 and C3.
// blob: [n, max|x|, x[npad], y[npad]], npad = n rounded up to even (16-byte aligned arrays).
struct ExpReg {
  static constexpr int NPAR = 2;
  static constexpr int NY = 1;
  static const char* name() { return "expreg"; }

  __device__ __forceinline__ static bool checkbounds(const double* theta, int npar, const mcmcb_ctx&) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < NPAR; i++) ok = ok && !(theta[i] <= 0.0);
    (void)npar;
    return ok;
  }
  __device__ __forceinline__ static double priorfun(const double* theta, int len, const mcmcb_ctx& c) {
    return mcmcb_default_priorfun(theta, len, c);
  }
  __device__ __forceinline__ static void ssfunction(const double* theta, int, int, const mcmcb_ctx& c, double* ss) {
    ss_impl<false>(theta, c, 0.0, ss);
  }
  // early-rejection form (external_inc.h:20-24, ssfunction_er(theta,npar,ny,sscrit)): the sum may stop once it has
  // reached sscrit.  Thread-per-chain only: every 128 data the warp votes, and leaves the loop when every lane's
  // partial sum (all terms are >= 0) has reached its own critical value.
  __device__ __forceinline__ static void ssfunction_er(const double* theta, int, int, const mcmcb_ctx& c, double sscrit,
                                                       double* ss) {
    ss_impl<true>(theta, c, sscrit, ss);
  }
  // one exponential of the data loop: through the direct table (mcmcb_expmul_direct) when the evaluation's whole
  // argument range fits it, else through the masked 2048-entry table -- bit-identical results either way
  template <bool DIRECT>
  __device__ __forceinline__ static double ex(double x, double ks, const mcmcb_ctx& c, double c1, double c2) {
    if constexpr (DIRECT) return mcmcb_expmul_direct(x, ks, c.exp_td, c1, c.exp_dn);
    else return mcmcb_expmul_fast(x, ks, c.exp_tl, c1, c2);
  }
  template <int B, bool DIRECT>
  __device__ __forceinline__ static void batch_loop(const double (&t1)[B], const double (&ks)[B], const mcmcb_ctx& c,
                                                    double (&acc)[B]) {
    const int n = (int)c.data[0];
    const int npad = (n + 1) & ~1;
    const double* __restrict__ x = c.data + 2;
    const double* __restrict__ y = c.data + 2 + npad;
    const double c1 = c.exp_c1, c2 = c.exp_c2;
    constexpr int U = (8 / B) < 2 ? 2 : (8 / B);  // data per trip: B*U exponentials in flight
    int i = 0;
    for (; i + U - 1 < n; i += U) {
      double xv[U], yv[U], e[B][U];
#pragma unroll
      for (int u = 0; u < U; u += 2) {
        const double2 xx = *reinterpret_cast<const double2*>(x + i + u);
        const double2 yy = *reinterpret_cast<const double2*>(y + i + u);
        xv[u] = xx.x; xv[u + 1] = xx.y; yv[u] = yy.x; yv[u + 1] = yy.y;
      }
#pragma unroll
      for (int b = 0; b < B; b++)
#pragma unroll
        for (int u = 0; u < U; u++) e[b][u] = ex<DIRECT>(xv[u], ks[b], c, c1, c2);
#pragma unroll
      for (int b = 0; b < B; b++)
#pragma unroll
        for (int u = 0; u < U; u++) {
          const double r = fma(-t1[b], e[b][u], yv[u]);
          acc[b] = fma(r, r, acc[b]);
        }
    }
    for (; i < n; i++) {
      const double xi = x[i], yi = y[i];
#pragma unroll
      for (int b = 0; b < B; b++) {
        const double r = fma(-t1[b], ex<DIRECT>(xi, ks[b], c, c1, c2), yi);
        acc[b] = fma(r, r, acc[b]);
      }
    }
  }
  // B parameter vectors (theta[b*npar + k]) in one sweep over the data, thread per chain: every datum read from
  // shared memory serves B chains.  Per chain the operations and their order are those of ssfunction.
  template <int B>
  __device__ __forceinline__ static void ssfunction_batch(const double* theta, int npar, int ny, const mcmcb_ctx& c,
                                                          double* ss) {
    const unsigned tl = c.exp_tl;
    const double xmax = fabs(c.data[1]);
    double t1[B], ks[B], acc[B];
    bool fast = tl != 0u && c.nlanes == 1;
    bool direct = c.exp_td != 0u && !(c.data[1] < 0.0);  // blob[1] < 0: some x is negative (blob_expreg)
#pragma unroll
    for (int b = 0; b < B; b++) {
      t1[b] = theta[b * npar];
      const double nt2 = -theta[b * npar + 1];
      fast = fast && fabs(nt2) * xmax < 700.0;
      direct = direct && nt2 <= 0.0 && mcmcb_exp_direct_ok(nt2, xmax, c.exp_dn);
      ks[b] = mcmcb_expmul_scale(nt2);
      acc[b] = 0.0;
    }
    if (!__all_sync(0xffffffffu, fast)) {  // rare: some exponent leaves the fast range -- one chain at a time
#pragma unroll
      for (int b = 0; b < B; b++) ssfunction(theta + b * npar, npar, ny, c, ss + b * NY);
      return;
    }
    if (__all_sync(0xffffffffu, direct)) batch_loop<B, true>(t1, ks, c, acc);
    else batch_loop<B, false>(t1, ks, c, acc);
#pragma unroll
    for (int b = 0; b < B; b++) ss[b * NY] = acc[b];
  }
  template <bool ER>
  __device__ __forceinline__ static void ss_impl(const double* theta, const mcmcb_ctx& c, double sscrit, double* ss) {
    const int n = (int)c.data[0];
    const int npad = (n + 1) & ~1;
    const double* __restrict__ x = c.data + 2;
    const double* __restrict__ y = c.data + 2 + npad;
    const double t1 = theta[0], nt2 = -theta[1];
    const unsigned tl = c.exp_tl;  // shared-window address of the staged 2^(j/2048) table
    const double c1 = c.exp_c1, c2 = c.exp_c2;
    double acc = 0.0;
    int i = c.lane;
    const int step = c.nlanes;
    // blob[1] = +-max|x| (set by blob_expreg): one range test per evaluation instead of one per
    // datum decides whether every exponent is inside mcmcb_exp_fast's range
    const double xmax = fabs(c.data[1]);
    const bool fast = tl != 0u && fabs(nt2) * xmax < 700.0;
    const double ks = mcmcb_expmul_scale(nt2);
    // the vote is a warp-wide operation: only when every lane of the warp runs the fast loop
    const bool vote = ER && __all_sync(0xffffffffu, fast);
    if (fast) {
      if (step == 1) {
        const bool direct = c.exp_td != 0u && !(c.data[1] < 0.0) && nt2 <= 0.0 && mcmcb_exp_direct_ok(nt2, xmax, c.exp_dn);
        if (!ER && __all_sync(0xffffffffu, direct)) {
          double t1a[1] = {t1}, ksa[1] = {ks}, acca[1] = {0.0};
          batch_loop<1, true>(t1a, ksa, c, acca);  // the same operations in the same order as the loop below
          ss[0] = acca[0];
          return;
        }
        // one lane owns the whole chain: consecutive data, 16-byte shared loads, 8 exps in flight
        for (; i + 7 < n; i += 8) {
          if (ER && vote && (i & 127) == 0 && __all_sync(0xffffffffu, acc >= sscrit)) { i = n; break; }
          double xv[8], yv[8];
#pragma unroll
          for (int u = 0; u < 8; u += 2) {
            const double2 xx = *reinterpret_cast<const double2*>(x + i + u);
            const double2 yy = *reinterpret_cast<const double2*>(y + i + u);
            xv[u] = xx.x; xv[u + 1] = xx.y; yv[u] = yy.x; yv[u + 1] = yy.y;
          }
#pragma unroll
          for (int u = 0; u < 8; u++) xv[u] = mcmcb_expmul_fast(xv[u], ks, tl, c1, c2);
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const double r = fma(-t1, xv[u], yv[u]);
            acc = fma(r, r, acc);
          }
        }
      } else {
        for (; i + 3 * step < n; i += 4 * step) {
          double e[4], yv[4];
#pragma unroll
          for (int u = 0; u < 4; u++) { e[u] = x[i + u * step]; yv[u] = y[i + u * step]; }
#pragma unroll
          for (int u = 0; u < 4; u++) e[u] = mcmcb_expmul_fast(e[u], ks, tl, c1, c2);
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const double r = fma(-t1, e[u], yv[u]);
            acc = fma(r, r, acc);
          }
        }
      }
      for (; i < n; i += step) {
        const double r = fma(-t1, mcmcb_expmul_fast(x[i], ks, tl, c1, c2), y[i]);
        acc = fma(r, r, acc);
      }
    } else {
      for (; i < n; i += step) {
        const double r = fma(-t1, exp(nt2 * x[i]), y[i]);
        acc = fma(r, r, acc);
      }
    }
    ss[0] = acc;
  }
};

// ---------------------------------------------------------------------------------------
// Run-time-npar models for the warp-per-chain kernel (NPAR = 0): theta lives in shared memory,
// the 32 lanes of the warp split the work and the kernel adds their partial sums.

// The exponential-decay regression again, for the warp-per-chain kernels (run-time npar = 2): the samplers that
// need an SVD factor (method='scam', condmax > 0) live there, and the reference's shipped testcase must be runnable
// with them.  Same blob as ExpReg; the lanes stride over the data.
struct ExpRegN {
  static constexpr int NPAR = 0;
  static constexpr int NY = 1;
  static const char* name() { return "expreg"; }
  __device__ __forceinline__ static bool checkbounds(const double* theta, int npar, const mcmcb_ctx&) {
    bool ok = true;
    for (int i = 0; i < npar; i++) ok = ok && !(theta[i] <= 0.0);
    return ok;
  }
  __device__ __forceinline__ static double priorfun(const double* theta, int len, const mcmcb_ctx& c) {
    return mcmcb_default_priorfun(theta, len, c);
  }
  __device__ __forceinline__ static void ssfunction(const double* theta, int, int, const mcmcb_ctx& c, double* ss) {
    const int n = (int)c.data[0];
    const int npad = (n + 1) & ~1;
    const double* __restrict__ x = c.data + 2;
    const double* __restrict__ y = c.data + 2 + npad;
    const double t1 = theta[0], nt2 = -theta[1];
    double acc = 0.0;
    for (int i = c.lane; i < n; i += c.nlanes) {
      const double r = fma(-t1, exp(nt2 * x[i]), y[i]);
      acc = fma(r, r, acc);
    }
    ss[0] = acc;
  }
};

// Gaussian target ss = (theta-mu)' Lam (theta-mu) (testcases/mcmcrun4.F90:47); BASELINE config C2.
// blob: [d, 0, mu[dpad], Lam[d*d]]; Lam must be symmetric (row i is read as column i).
struct GaussN {
  static constexpr int NPAR = 0;
  static constexpr int NY = 1;
  static const char* name() { return "gauss"; }
  __device__ __forceinline__ static bool checkbounds(const double*, int, const mcmcb_ctx&) { return true; }
  __device__ __forceinline__ static double priorfun(const double* theta, int len, const mcmcb_ctx& c) {
    return mcmcb_default_priorfun(theta, len, c);
  }
  __device__ __forceinline__ static void ssfunction(const double* theta, int npar, int, const mcmcb_ctx& c,
                                                    double* ss) {
    const int d = npar, dpad = (d + 1) & ~1;
    const double* __restrict__ mu = c.data + 2;
    const double* __restrict__ lam = c.data + 2 + dpad;
    double acc = 0.0;
    if (c.scratch != nullptr) {
      // theta - mu once per evaluation instead of once per matrix entry (same values, same order of the sum)
      double* df = c.scratch;
      for (int j = c.lane; j < d; j += c.nlanes) df[j] = theta[j] - mu[j];
      mcmcb_sync_lanes(c);
      for (int i = c.lane; i < d; i += c.nlanes) {
        const double* col = lam + i;
        double w = 0.0;
        int j = 0;
        for (; j + 4 <= d; j += 4) {
          const double l0 = col[0], l1 = col[d], l2 = col[2 * d], l3 = col[3 * d];
          w = fma(l0, df[j], w); w = fma(l1, df[j + 1], w); w = fma(l2, df[j + 2], w); w = fma(l3, df[j + 3], w);
          col += 4 * d;
        }
        for (; j < d; j++, col += d) w = fma(col[0], df[j], w);
        acc = fma(w, df[i], acc);
      }
      mcmcb_sync_lanes(c);
    } else {
      for (int i = c.lane; i < d; i += c.nlanes) {
        double w = 0.0;
        for (int j = 0; j < d; j++) w = fma(lam[(size_t)j * d + i], theta[j] - mu[j], w);
        acc = fma(w, theta[i] - mu[i], acc);
      }
    }
    ss[0] = acc;
  }
};

// The same Gaussian target with a COMPILE-TIME number of parameters, for the register kernel (thread or lane group per
// chain): testcases/mcmcrun4.F90 has npar = 5 (mcmcrun4.F90:11), and a warp per chain would idle 27 of its 32 lanes on
// it.  Registered under the same name "gauss" for D = 3..8 (builtin_gauss_k1_*.cu); mcmcb_set_initial picks the
// register kernel when the run-time npar has a registration (and the sampler needs no SVD factor), else the
// warp-per-chain kernel.  Same blob as GaussN.
template <int D>
struct GaussK {
  static constexpr int NPAR = D;
  static constexpr int NY = 1;
  static const char* name() { return "gauss"; }
  __device__ __forceinline__ static bool checkbounds(const double*, int, const mcmcb_ctx&) { return true; }
  __device__ __forceinline__ static double priorfun(const double* theta, int len, const mcmcb_ctx& c) {
    return mcmcb_default_priorfun(theta, len, c);
  }
  __device__ __forceinline__ static void ssfunction(const double* theta, int, int, const mcmcb_ctx& c, double* ss) {
    constexpr int dpad = (D + 1) & ~1;
    const double* __restrict__ mu = c.data + 2;
    const double* __restrict__ lam = c.data + 2 + dpad;
    double df[D];
#pragma unroll
    for (int j = 0; j < D; j++) df[j] = theta[j] - mu[j];
    double acc = 0.0;
    if (c.nlanes == 1) {  // one lane owns the chain: rows in order, like the reference's matmul + dot_product
#pragma unroll
      for (int i = 0; i < D; i++) {
        double w = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) w = fma(lam[j * D + i], df[j], w);
        acc = fma(w, df[i], acc);
      }
    } else {
      for (int i = c.lane; i < D; i += c.nlanes) {
        double w = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) w = fma(lam[j * D + i], df[j], w);
        double dfi = df[0];
#pragma unroll
        for (int j = 1; j < D; j++) dfi = (j == i) ? df[j] : dfi;  // df[i] without a dynamically indexed local array
        acc = fma(w, dfi, acc);
      }
    }
    ss[0] = acc;
  }
};

// Twisted Gaussian ("banana"): phi = (t1, t2 + b t1^2 - 100 b, t3..td),
// ss = phi1^2/100 + sum_{i>=2} phi_i^2.   blob: [d, b]
struct BananaN {
  static constexpr int NPAR = 0;
  static constexpr int NY = 1;
  static const char* name() { return "banana"; }
  __device__ __forceinline__ static bool checkbounds(const double*, int, const mcmcb_ctx&) { return true; }
  __device__ __forceinline__ static double priorfun(const double* theta, int len, const mcmcb_ctx& c) {
    return mcmcb_default_priorfun(theta, len, c);
  }
  __device__ __forceinline__ static void ssfunction(const double* theta, int npar, int, const mcmcb_ctx& c,
                                                    double* ss) {
    const double b = c.data[1];
    double acc = 0.0;
    for (int i = c.lane; i < npar; i += c.nlanes) {
      double v;
      if (i == 0) { v = theta[0]; acc = fma(v, v / 100.0, acc); continue; }
      if (i == 1) v = theta[1] + b * theta[0] * theta[0] - 100.0 * b;
      else v = theta[i];
      acc = fma(v, v, acc);
    }
    ss[0] = acc;
  }
};

// Hierarchical normal means: params (theta_1..G, mu, log tau), y_gj ~ N(theta_g, 1),
// theta_g ~ N(mu, tau^2), weak hyperpriors mu ~ N(0,10^2), log tau ~ N(0,2^2).   blob: [G, J, y[J][G]] -- observation
// j of every group contiguous, so that lanes (one group each) read consecutive shared-memory addresses
struct HierN {
  static constexpr int NPAR = 0;
  static constexpr int NY = 1;
  static const char* name() { return "hier"; }
  __device__ __forceinline__ static bool checkbounds(const double*, int, const mcmcb_ctx&) { return true; }
  __device__ __forceinline__ static double priorfun(const double* theta, int len, const mcmcb_ctx& c) {
    return mcmcb_default_priorfun(theta, len, c);
  }
  __device__ __forceinline__ static void ssfunction(const double* theta, int, int, const mcmcb_ctx& c, double* ss) {
    const int G = (int)c.data[0], J = (int)c.data[1];
    const double* __restrict__ y = c.data + 2;
    const double mu = theta[G], ltau = theta[G + 1];
    const double itau2 = exp(-2.0 * ltau);
    double acc = 0.0;
    for (int g = c.lane; g < G; g += c.nlanes) {
      const double tg = theta[g];
      double a = 0.0;
      for (int j = 0; j < J; j++) { const double r = y[(size_t)j * G + g] - tg; a = fma(r, r, a); }
      const double dm = tg - mu;
      acc += a + dm * dm * itau2;
    }
    if (c.lane == 0) acc += 2.0 * G * ltau + mu * mu / 100.0 + ltau * ltau / 4.0;
    ss[0] = acc;
  }
  // The same sum over a VIEW of the parameter vector (anything with operator[]): the thread-per-chain SCAM kernels
  // (k5_scam.cuh, k5s_scam.cuh) hand in theta + delta U(:,j) composed on the fly, so a single-component move
  // (MCMC_propose_sc, MCMC_run_scam.F90:94-117) never materialises its proposal.  The chain's lanes take the groups
  // round-robin (lane, lane + nlanes, ...) and return partial sums; one lane = every group, in ssfunction's order.  W
  // groups' sums of squares run as independent accumulation chains (W = mcmcb_view_ilp<V>) and join the total in group
  // order; the batched form does that for C chains of one thread at once, every datum read once for all of them.
  static constexpr bool MCMCB_VIEW_DEFAULTS = true;  // checkbounds is always true, priorfun is the default prior

  // WB groups g0, g0 + nl, ... of C chains
  template <int C, int WB, class V>
  __device__ __forceinline__ static void view_block(const V* theta, const double* __restrict__ y, int G, int J, int g0, int nl,
                                                    const double* mu, const double* itau2, double* acc) {
    double tg[C][WB], a[C][WB];
#pragma unroll
    for (int c = 0; c < C; c++)
#pragma unroll
      for (int q = 0; q < WB; q++) { tg[c][q] = theta[c][g0 + nl * q]; a[c][q] = 0.0; }
    for (int j = 0; j < J; j++) {
      const double* yj = y + (size_t)j * G + g0;
#pragma unroll
      for (int q = 0; q < WB; q++) {
        const double yv = yj[nl * q];
#pragma unroll
        for (int c = 0; c < C; c++) { const double r = yv - tg[c][q]; a[c][q] = fma(r, r, a[c][q]); }
      }
    }
#pragma unroll
    for (int c = 0; c < C; c++)
#pragma unroll
      for (int q = 0; q < WB; q++) { const double dm = tg[c][q] - mu[c]; acc[c] += a[c][q] + dm * dm * itau2[c]; }
  }

  template <int C, class V>
  __device__ __forceinline__ static void ssfunction_view_batch(const V* theta, int, int, const mcmcb_ctx& c, double* ss) {
    const int G = (int)c.data[0], J = (int)c.data[1];
    const double* __restrict__ y = c.data + 2;
    constexpr int W = mcmcb_view_ilp<V>::value;  // groups in flight per chain (4 unless the view asks otherwise)
    static_assert(W == 1 || W == 2 || W == 4 || W == 8, "groups in flight: a power of two up to 8");
    double mu[C], ltau[C], itau2[C], acc[C];
#pragma unroll
    for (int k = 0; k < C; k++) { mu[k] = theta[k][G]; ltau[k] = theta[k][G + 1]; itau2[k] = exp(-2.0 * ltau[k]); acc[k] = 0.0; }
    const int lane = c.lane, nl = c.nlanes;
    const int mine = (G - lane + nl - 1) / nl;  // groups of this lane
    int i = 0;
    for (; i + W <= mine; i += W) view_block<C, W>(theta, y, G, J, lane + nl * i, nl, mu, itau2, acc);
    // the ragged rest in blocks of W/2, W/4, ...: no group is evaluated twice, the join order stays the group order
    if constexpr (W >= 8) { if (mine - i >= 4) { view_block<C, 4>(theta, y, G, J, lane + nl * i, nl, mu, itau2, acc); i += 4; } }
    if constexpr (W >= 4) { if (mine - i >= 2) { view_block<C, 2>(theta, y, G, J, lane + nl * i, nl, mu, itau2, acc); i += 2; } }
    if constexpr (W >= 2) { if (mine - i >= 1) { view_block<C, 1>(theta, y, G, J, lane + nl * i, nl, mu, itau2, acc); i += 1; } }
#pragma unroll
    for (int k = 0; k < C; k++) {
      if (lane == 0) acc[k] += 2.0 * G * ltau[k] + mu[k] * mu[k] / 100.0 + ltau[k] * ltau[k] / 4.0;
      ss[k] = acc[k];  // NY = 1
    }
  }

  template <class V>
  __device__ __forceinline__ static void ssfunction_view(const V& theta, int len, int ny, const mcmcb_ctx& c, double* ss) {
    ssfunction_view_batch<1>(&theta, len, ny, c, ss);
  }
};

}  // namespace mcmcb
