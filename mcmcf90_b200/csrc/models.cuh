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
    const int n = (int)c.data[0];
    const int npad = (n + 1) & ~1;
    const double* __restrict__ x = c.data + 2;
    const double* __restrict__ y = c.data + 2 + npad;
    const double t1 = theta[0], nt2 = -theta[1];
    const double* __restrict__ tab = c.exp2_tab;
    const int l16 = c.tab_slot;
    double acc = 0.0;
    int i = c.lane;
    const int step = c.nlanes;
    // blob[1] = max|x| (set by blob_expreg): one range test per evaluation instead of one per
    // datum decides whether every exponent is inside mcmcb_exp_fast's range
    const bool fast = tab != nullptr && fabs(nt2) * c.data[1] < 700.0;
    if (fast) {
      if (step == 1) {
        // one lane owns the whole chain: consecutive data, 16-byte shared loads, 8 exps in flight
        for (; i + 7 < n; i += 8) {
          double xv[8], yv[8];
#pragma unroll
          for (int u = 0; u < 8; u += 2) {
            const double2 xx = *reinterpret_cast<const double2*>(x + i + u);
            const double2 yy = *reinterpret_cast<const double2*>(y + i + u);
            xv[u] = xx.x; xv[u + 1] = xx.y; yv[u] = yy.x; yv[u + 1] = yy.y;
          }
#pragma unroll
          for (int u = 0; u < 8; u++) xv[u] = mcmcb_exp_fast(nt2 * xv[u], tab, l16);
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
          for (int u = 0; u < 4; u++) { e[u] = nt2 * x[i + u * step]; yv[u] = y[i + u * step]; }
#pragma unroll
          for (int u = 0; u < 4; u++) e[u] = mcmcb_exp_fast(e[u], tab, l16);
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const double r = fma(-t1, e[u], yv[u]);
            acc = fma(r, r, acc);
          }
        }
      }
      for (; i < n; i += step) {
        const double r = fma(-t1, mcmcb_exp_fast(nt2 * x[i], tab, l16), y[i]);
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

}  // namespace mcmcb
