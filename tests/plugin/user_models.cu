// Example user plugin (tests/test_plugin.py): two models supplied from OUTSIDE the library, the way a user of
// the reference links their own ssfunction.  "user_expreg" is the testcase model written naively with libm's
// exp() (testcases/mcmcrun.F90:89-122); "user_isogauss" is ss = sum(theta^2) for any npar.
#include "mcmcb200_plugin.cuh"

struct UserExpReg {
  static constexpr int NPAR = 2;
  static constexpr int NY = 1;
  static const char* name() { return "user_expreg"; }
  __device__ static bool checkbounds(const double* theta, int, const mcmcb_ctx&) { return theta[0] > 0.0 && theta[1] > 0.0; }
  __device__ static double priorfun(const double* theta, int len, const mcmcb_ctx& c) { return mcmcb_default_priorfun(theta, len, c); }
  __device__ static void ssfunction(const double* theta, int, int, const mcmcb_ctx& c, double* ss) {
    // blob: [n, x[n], y[n]] -- the plugin owns its data layout
    const int n = (int)c.data[0];
    const double* x = c.data + 1;
    const double* y = c.data + 1 + n;
    double acc = 0.0;
    for (int i = c.lane; i < n; i += c.nlanes) {
      const double r = y[i] - theta[0] * exp(-theta[1] * x[i]);
      acc += r * r;
    }
    ss[0] = acc;
  }
};
MCMCB_REGISTER_MODEL_K1(UserExpReg)

struct UserIsoGauss {
  static constexpr int NPAR = 0;  // run-time npar: warp-per-chain kernels
  static constexpr int NY = 1;
  static const char* name() { return "user_isogauss"; }
  __device__ static bool checkbounds(const double*, int, const mcmcb_ctx&) { return true; }
  __device__ static double priorfun(const double*, int, const mcmcb_ctx&) { return 0.0; }
  __device__ static void ssfunction(const double* theta, int npar, int, const mcmcb_ctx& c, double* ss) {
    double acc = 0.0;
    for (int i = c.lane; i < npar; i += c.nlanes) acc = fma(theta[i], theta[i], acc);
    ss[0] = acc;
  }
};
MCMCB_REGISTER_MODEL_K2(UserIsoGauss)
