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
  double exp_c1, exp_c2;    /* MCMCB_EXP_C1L, MCMCB_EXP_C2L handed through the kernel-parameter bank (see mcmcb_expmul_fast) */
  double* scratch;          /* warp-per-chain kernels: npar doubles of shared memory private to the chain's warp
                               (nullptr in the register kernel) */
  unsigned exp_tl;          /* shared-window byte address of this thread's column of the replicated 2^(j/256)
                               table (see mcmcb_exp; mcmcb_exp_column()); 0 = no table staged, use exp() */
};

/* ---------------------------------------------------------------------------------------
 * mcmcb_exp: FP64 exp() for model code, built for the FP64 pipe of sm_100a.
 * exp(a) = 2^m * 2^(j/256) * e^u with k = round(a*256/ln2) = 256 m + j and |u| <= ln2/512, so a
 * degree-4 polynomial reaches double precision (economised: max error 2.4e-18).  2^(j/256) comes
 * from a 256-entry table that the sampling kernels stage in shared memory with every entry
 * replicated 16 times -- hardware lane l reads copy (l mod 16), so the 16 lanes of a half-warp
 * hit 16 distinct 8-byte bank pairs and the lookup is conflict-free whatever j each lane needs.
 * The fast paths are branch-free (independent calls interleave in the instruction stream).
 *
 *   mcmcb_exp_fast(a)         9 FP64 instructions  (libdevice exp(): 16)
 *   mcmcb_expmul_fast(x, ks)  8 FP64 instructions  = exp(x * s) with ks = mcmcb_expmul_scale(s)
 *                             computed once per evaluation: the product, the range reduction
 *                             and the change of units to ln2/256 are one DFMA pair.
 *
 * An FP64 warp instruction occupies the scheduler's issue port for two cycles on sm_100a, so
 * the instruction COUNT of the datum loop, integer and load instructions included, is what
 * sets the speed of a model evaluation.  mcmcb_exp_ok() tells whether an
 * argument is in the fast range (|a| < 708: result normal, no overflow, not NaN).
 * ------------------------------------------------------------------------------------- */
#define MCMCB_EXP_TAB_N 256
#define MCMCB_EXP_TAB_REP 16
#define MCMCB_EXP_TAB_DOUBLES (MCMCB_EXP_TAB_N * MCMCB_EXP_TAB_REP)

/* coefficients live in the constant bank so that DFMA/DMUL read them as c[][] operands instead
 * of re-materialising 64-bit immediates inside the loop */
__constant__ double MCMCB_EXPC[10] = {
    0x1.71547652b82fep+8,   /* [0] 256/ln2 */
    -0x1.62e42fee00000p-9,  /* [1] -ln2/256, high part (21 trailing zero bits: k*hi is exact) */
    -0x1.a39ef35793c76p-41, /* [2] -ln2/256, low part */
    /* e^u - 1 = u (c1 + u (1/2 + u (c3 + u/24))) on |u| <= h = ln2/512: degree-5 Taylor with the u^5
     * term Chebyshev-economised into c1, c3 (max error 2.4e-18 against 3.8e-17 for plain degree 4) */
    0x1.5555555555555p-5,   /* [3] 1/24 */
    0x1.555557e54fd55p-3,   /* [4] c3 = 1/6 + h^2/96 */
    /* the same polynomial in r = u/L, L = ln2/256, r in [-1/2, 1/2] */
    0x1.62e42fefa39b9p-9,   /* [5] c1 L */
    0x1.ebfbdff82c58fp-19,  /* [6] L^2/2 */
    0x1.c6b090da1e082p-29,  /* [7] c3 L^3 */
    0x1.3b2ab6fba4e77p-39,  /* [8] L^4/24 */
    0x1.fffffffffffb1p-1};  /* [9] c1 = 1 - h^4/384 */

/* [5] and [6] again as macros: the host writes them into the kernel parameters (mcmcb_ctx::exp_c1/c2) */
#define MCMCB_EXP_C1L 0x1.62e42fefa39b9p-9
#define MCMCB_EXP_C2L 0x1.ebfbdff82c58fp-19

#define MCMCB_EXP_MAGIC 6755399441055744.0 /* 1.5 * 2^52: rounds to integer, k in the low word */

/* 2^(k/256) * (1 + s) from the reduced pieces.  Three integer instructions: mask, address, and
 * ONE multiply-add for the exponent: table entry j is stored with j*2^12 subtracted from its high
 * word, so that adding k*2^12 = m*2^20 + j*2^12 to it yields the high word of 2^m * 2^(j/256)
 * without isolating m.  `tl` is the shared-window address of the thread's own column of the
 * replicated table; the load is spelled as ld.shared so that the address stays one mask and one
 * shift-add of an opaque per-thread base. */
__device__ __forceinline__ double mcmcb_exp_assemble(int k, double s, unsigned tl) {
  double tj;
  unsigned addr;
  asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(addr) : "r"(k & (MCMCB_EXP_TAB_N - 1)), "r"(tl));  /* 8 B * 16 columns */
  asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(addr));
  const double sc = __hiloint2double(k * 4096 + __double2hiint(tj), __double2loint(tj));
  return fma(sc, s, sc);
}
/* address of this thread's column of a table staged at smem_tab (16 columns, lane mod 16) */
__device__ __forceinline__ unsigned mcmcb_exp_column(const double* smem_tab) {
  unsigned a = (unsigned)__cvta_generic_to_shared(smem_tab) + 8u * (threadIdx.x & (MCMCB_EXP_TAB_REP - 1));
  asm volatile("" : "+r"(a));  /* opaque: keeps the compiler from folding the lane term back into every lookup */
  return a;
}

__device__ __forceinline__ double mcmcb_exp_fast(double a, unsigned tl) {
  double t = fma(a, MCMCB_EXPC[0], MCMCB_EXP_MAGIC);
  const int k = __double2loint(t);
  t -= MCMCB_EXP_MAGIC;
  double r = fma(t, MCMCB_EXPC[1], a);
  r = fma(t, MCMCB_EXPC[2], r);
  double q = fma(r, MCMCB_EXPC[3], MCMCB_EXPC[4]);
  q = fma(r, q, 0.5);
  q = fma(r, q, MCMCB_EXPC[9]);
  return mcmcb_exp_assemble(k, r * q, tl);
}

/* ks for mcmcb_expmul_fast: s * 256/ln2 */
__device__ __forceinline__ double mcmcb_expmul_scale(double s) { return s * MCMCB_EXPC[0]; }

/* exp(x * s), ks = mcmcb_expmul_scale(s).  The reduced argument r = x*ks - round(x*ks) is formed
 * by one DFMA (exact product, one rounding of a value <= 1/2), so the only argument error is the
 * rounding of ks itself: |x s| * 2^-52 relative in the result, the same size as the rounding of
 * the product x*s inside exp(x*s).
 *
 * Operand placement matters: measured on B200 (scripts/ubench_fp64_operands.cu) a DFMA that reads
 * three DISTINCT 64-bit registers holds the FP64 pipe for 3 cycles, one that reads at most two
 * (plus a uniform-register / constant / immediate operand, or a repeated register) for 2.  The
 * Horner steps below are (r, q, constant): the constant must not be hoisted into a vector
 * register.  ptxas 12.9 keeps at most two hoisted constants per bank in uniform registers, so
 * c1, c2 arrive through the kernel-parameter bank (c[0x0], mcmcb_ctx) and c3, c4 through the
 * __constant__ bank (c[0x3]); the first step (r, c4, c3) then reads r + one register. */
__device__ __forceinline__ double mcmcb_expmul_fast(double x, double ks, unsigned tl, double c1, double c2) {
  const double t = fma(x, ks, MCMCB_EXP_MAGIC);
  const int k = __double2loint(t);
  const double r = fma(x, ks, MCMCB_EXP_MAGIC - t);
  double q = fma(r, MCMCB_EXPC[8], MCMCB_EXPC[7]);
  q = fma(r, q, c2);
  q = fma(r, q, c1);
  return mcmcb_exp_assemble(k, r * q, tl);
}
__device__ __forceinline__ double mcmcb_expmul_fast(double x, double ks, unsigned tl) {
  return mcmcb_expmul_fast(x, ks, tl, MCMCB_EXPC[5], MCMCB_EXPC[6]);
}

/* true when the fast paths are valid for argument a: |a| < 708 and a is not NaN */
__device__ __forceinline__ bool mcmcb_exp_ok(double a) {
  return (unsigned)(__double2hiint(a) & 0x7fffffff) < 0x40862000u;
}
__device__ __forceinline__ double mcmcb_exp(double a, const mcmcb_ctx& c) {
  if (c.exp_tl != 0u && mcmcb_exp_ok(a)) return mcmcb_exp_fast(a, c.exp_tl);
  return exp(a);
}

/* correctly rounded 2^(j/256) (scripts/gen_exp_table.py, mpmath at 80 digits) */
__device__ static const double MCMCB_EXP2_TABLE[MCMCB_EXP_TAB_N] = {
#include "mcmcb200_exp_table.inc"
};

/* stage the replicated table; call from every thread of the CTA, then __syncthreads() */
__device__ __forceinline__ void mcmcb_stage_exp_table(double* smem_tab) {
  for (int i = threadIdx.x; i < MCMCB_EXP_TAB_DOUBLES; i += blockDim.x)
  {
    const int j = i / MCMCB_EXP_TAB_REP;
    const double v = MCMCB_EXP2_TABLE[j];
    smem_tab[i] = __hiloint2double(__double2hiint(v) - j * 4096, __double2loint(v));  /* see mcmcb_exp_assemble */
  }
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
