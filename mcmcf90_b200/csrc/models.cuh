// Built-in user models (the plugin side of external_inc.h:14-28), written against
// include/mcmcb200_model.cuh.  Blob layouts are shared with the host helpers in
// mcmcf90_b200/models.py.
#pragma once
#include "mcmcb200_model.cuh"

namespace mcmcb {

// Exponential-decay regression y = theta1*exp(-theta2*x) (testcases/mcmcrun.F90:89,104),
// bounds theta > 0 (testcases/mcmcrun.F90:112-122).  BASELINE configs C1 and C3.
// blob: [n, 0, x[npad], y[npad]], npad = n rounded up to even (16-byte aligned arrays).
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
    double acc = 0.0;
    int i = c.lane;
    const int step = c.nlanes;
    // 4 independent exp chains in flight per lane; accumulation stays in index order
    for (; i + 3 * step < n; i += 4 * step) {
      double e0 = exp(nt2 * x[i]);
      double e1 = exp(nt2 * x[i + step]);
      double e2 = exp(nt2 * x[i + 2 * step]);
      double e3 = exp(nt2 * x[i + 3 * step]);
      double r0 = fma(-t1, e0, y[i]);
      double r1 = fma(-t1, e1, y[i + step]);
      double r2 = fma(-t1, e2, y[i + 2 * step]);
      double r3 = fma(-t1, e3, y[i + 3 * step]);
      acc = fma(r0, r0, acc);
      acc = fma(r1, r1, acc);
      acc = fma(r2, r2, acc);
      acc = fma(r3, r3, acc);
    }
    for (; i < n; i += step) {
      double r = fma(-t1, exp(nt2 * x[i]), y[i]);
      acc = fma(r, r, acc);
    }
    ss[0] = acc;
  }
};

}  // namespace mcmcb
