/*
 * mcmcb200_model.cuh -- device-side user-model contract.
 *
 * In the reference the user model is three external procedures resolved at link time
 * (external_inc.h:4-28): ssfunction(theta,npar,ny) -> ss(ny) = -2 log p(y|theta),
 * priorfun(theta,len) -> -2 log p(theta), checkbounds(theta) -> logical.  Here the same
 * three functions are static __device__ members of a model struct that the sampling
 * kernels are instantiated with (compile-time resolution == full inlining of the hot
 * ssfunction loop).  Names, argument order and meaning follow the Fortran interface; the
 * only addition is the context argument, which carries what the Fortran plugin keeps in
 * `save`d variables (its data, testcases/mcmcrun.F90:69-86) plus the cooperative-lane
 * coordinates.
 *
 * Cooperative evaluation: `nlanes` threads share one chain.  ssfunction must return the
 * PARTIAL sum over the data items this lane owns (i = lane, lane+nlanes, ...); the kernel
 * adds the partials with warp shuffles.  Terms that do not depend on the data index must
 * be added by lane 0 only.  priorfun and checkbounds are evaluated redundantly by every
 * lane and must be lane-independent.
 *
 * A model struct provides:
 *   static constexpr int NPAR;   // >0: compile-time npar (small-npar register kernel)
 *                                //  0: runtime npar (large-npar warp kernel)
 *   static constexpr int NY;     // nycol, number of ss columns (usually 1)
 *   static const char* name();
 *   __device__ static bool   checkbounds(const double* theta, int npar, const mcmcb_ctx& c);
 *   __device__ static double priorfun   (const double* theta, int len,  const mcmcb_ctx& c);
 *   __device__ static void   ssfunction (const double* theta, int npar, int ny,
 *                                        const mcmcb_ctx& c, double* ss);
 */
#ifndef MCMCB200_MODEL_CUH
#define MCMCB200_MODEL_CUH

struct mcmcb_ctx {
  const double* data;       /* model blob: shared memory when it fits (TMA-staged once per CTA), else global */
  unsigned long long ndata; /* blob length in doubles */
  const double* prior;      /* default Gaussian prior: mu[npar] then sig[npar]; nullptr = flat */
  int lane, nlanes;         /* this thread's rank among the lanes that share the chain */
};

/* default prior, priorfun.f90:97-100: sum(((theta-mu)/sig)**2, mask = sig>0) */
__device__ __forceinline__ double mcmcb_default_priorfun(const double* theta, int len, const mcmcb_ctx& c) {
  if (c.prior == nullptr) return 0.0;
  double p = 0.0;
  for (int i = 0; i < len; i++) {
    double sg = c.prior[len + i];
    if (sg > 0.0) {
      double t = (theta[i] - c.prior[i]) / sg;
      p += t * t;
    }
  }
  return p;
}

#endif
