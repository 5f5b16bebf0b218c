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
  const double* exp2_tab;   /* shared memory: 2^(j/64), j=0..63, each entry replicated 16x (see mcmcb_exp) */
  int tab_slot;             /* which of the 16 table copies this thread reads: physical lane id mod 16 */
};

/* ---------------------------------------------------------------------------------------
 * mcmcb_exp: FP64 exp() for model code, built for the FP64 pipe of sm_100a.
 * exp(a) = 2^m * 2^(j/64) * e^r with k = round(a*64/ln2) = 64 m + j and |r| <= ln2/128, so a
 * degree-5 polynomial reaches double precision: 10 FP64-pipe instructions per call against 16
 * for libdevice's table-free exp().  2^(j/64) comes from a 64-entry table that the sampling
 * kernels stage in shared memory with every entry replicated 16 times -- hardware lane l reads copy
 * (l mod 16), so the 16 lanes of a half-warp hit 16 distinct 8-byte bank pairs and the lookup
 * is conflict-free whatever j each lane needs.  The fast path is branch-free (independent
 * calls interleave in the instruction stream); mcmcb_exp_ok() tells whether the argument is
 * in the fast range (|a| < 708), otherwise the caller falls back to exp().
 * ------------------------------------------------------------------------------------- */
#define MCMCB_EXP_TAB_N 64
#define MCMCB_EXP_TAB_REP 16
#define MCMCB_EXP_TAB_DOUBLES (MCMCB_EXP_TAB_N * MCMCB_EXP_TAB_REP)

/* coefficients live in the constant bank so that DFMA/DMUL read them as c[][] operands instead
 * of re-materialising 64-bit immediates inside the loop (FP64 instructions hold the issue port
 * for two cycles on sm_100a, so every other instruction in the loop costs a full cycle) */
__constant__ double MCMCB_EXPC[6] = {
    92.33248261689366,          /* 64/ln2 */
    -0x1.62e42fef00000p-7,      /* -ln2/64, high 33 bits */
    -0x1.473de6af278edp-40,     /* -ln2/64, low part */
    8.3333333333333332e-3,      /* 1/120 */
    4.1666666666666664e-2,      /* 1/24 */
    1.6666666666666666e-1};     /* 1/6 */

__device__ __forceinline__ double mcmcb_exp_fast(double a, const double* __restrict__ tab, int lane16) {
  const double MAGIC = 6755399441055744.0; /* 1.5 * 2^52: rounds to integer, k in the low word */
  double t = fma(a, MCMCB_EXPC[0], MAGIC);
  const int k = __double2loint(t);
  t -= MAGIC;
  double r = fma(t, MCMCB_EXPC[1], a);
  r = fma(t, MCMCB_EXPC[2], r);
  double q = fma(r, MCMCB_EXPC[3], MCMCB_EXPC[4]);
  q = fma(r, q, MCMCB_EXPC[5]);
  q = fma(r, q, 0.5);
  q = fma(r, q, 1.0);
  const double s = r * q; /* e^r - 1 */
  const double tj = tab[((k & (MCMCB_EXP_TAB_N - 1)) * MCMCB_EXP_TAB_REP) | lane16];
  const double res = fma(tj, s, tj);
  return __hiloint2double(__double2hiint(res) + (k >> 6) * 1048576, __double2loint(res));
}
/* true when mcmcb_exp_fast(a) is valid: |a| < 708 (result normal, no overflow) and a is not NaN */
__device__ __forceinline__ bool mcmcb_exp_ok(double a) {
  return (unsigned)(__double2hiint(a) & 0x7fffffff) < 0x40862000u;
}
__device__ __forceinline__ double mcmcb_exp(double a, const mcmcb_ctx& c) {
  if (c.exp2_tab != nullptr && mcmcb_exp_ok(a)) return mcmcb_exp_fast(a, c.exp2_tab, c.tab_slot);
  return exp(a);
}

/* correctly rounded 2^(j/64) (generated with mpmath at 60 digits) */
__device__ static const double MCMCB_EXP2_TABLE[MCMCB_EXP_TAB_N] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};

/* stage the replicated table; call from every thread of the CTA, then __syncthreads() */
__device__ __forceinline__ void mcmcb_stage_exp_table(double* smem_tab) {
  for (int i = threadIdx.x; i < MCMCB_EXP_TAB_DOUBLES; i += blockDim.x)
    smem_tab[i] = MCMCB_EXP2_TABLE[i / MCMCB_EXP_TAB_REP];
}

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
