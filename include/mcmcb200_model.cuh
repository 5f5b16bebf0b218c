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
 * Cooperative evaluation: `nlanes` threads share one chain (a fraction of a warp, a warp, or a group of warps).  ssfunction must return the
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
 *
 * Optional members (detected at compile time; a model without them runs unchanged):
 *   ssfunction_er(theta,npar,ny,ctx,sscrit,ss)   early-rejection form (external_inc.h:20-24)
 *   template<int B> ssfunction_batch(...)        B parameter vectors in one sweep over the data (register kernel)
 *   template<class V> ssfunction_view(const V& theta,npar,ny,ctx,ss)  ss of a VIEW of the parameter vector (anything with
 *       operator[]): the thread-per-chain SCAM kernel hands in theta + delta*U(:,j) composed on the fly, so a
 *       single-component move (MCMC_run_scam.F90:94-117) never materialises its proposal; needs
 *       `static constexpr bool MCMCB_VIEW_DEFAULTS = true` (checkbounds always true, priorfun = the default prior).
 *       The lanes ctx.lane of ctx.nlanes that share the chain each return a partial sum; the kernel adds them.
 *   template<int C, class V> ssfunction_view_batch(const V theta[C],npar,ny,ctx,double ss[C*ny])  the same for C
 *       chains of one thread in ONE sweep over the data: a datum read from shared memory is used C times (the SCAM
 *       kernel with theta in shared memory is bound by shared-memory wavefronts, not by FP64, without it).
 */
#ifndef MCMCB200_MODEL_CUH
#define MCMCB200_MODEL_CUH

/* what ssfunction_view is probed with (detection only) */
struct mcmcb_view_probe { __device__ double operator[](int) const { return 0.0; } };

/* how many independent accumulation chains a view asks the model to keep in flight (a view type may carry a
   `static constexpr int ILP`; kernels that run few warps per SM ask for more).  Models are free to ignore it: it must not
   change the order in which the partial sums join the total. */
template <class V, class = void>
struct mcmcb_view_ilp { static constexpr int value = 4; };
template <class V>
struct mcmcb_view_ilp<V, decltype((void)V::ILP)> { static constexpr int value = V::ILP; };

struct mcmcb_ctx {
  const double* data;       /* model blob: shared memory when it fits (TMA-staged once per CTA), else global */
  unsigned long long ndata; /* blob length in doubles */
  const double* prior;      /* default Gaussian prior: mu[npar] then sig[npar]; nullptr = flat */
  int lane, nlanes;         /* this thread's rank among the lanes that share the chain */
  double exp_c1, exp_c2;    /* MCMCB_EXP_C1L, MCMCB_EXP_C2L handed through the kernel-parameter bank (see mcmcb_expmul_fast) */
  double* scratch;          /* warp-per-chain kernels: npar doubles of shared memory private to the chain's warp
                               (nullptr in the register kernel) */
  unsigned exp_tl;          /* shared-window byte address of the staged 2^(j/2048) table (see mcmcb_exp;
                               mcmcb_exp_column()); 0 = no table staged, use exp() */
  int bar_id = 0, bar_threads = 0; /* nlanes > 32 (a group of warps shares the chain): the group's named barrier,
                               used by mcmcb_sync_lanes() */
  unsigned exp_td = 0;      /* shared-window byte address of entry k = 0 of the DIRECT table 2^(k/2048), k = -(exp_dn-1)..0
                               (see mcmcb_expmul_direct); 0 = not staged */
  int exp_dn = 0;           /* entries of the direct table */
};

/* Make the `scratch` writes of every lane that shares the chain visible to all of them: __syncwarp() when the
 * lanes are one warp (or less), the group's named barrier when several warps share the chain.  Every lane of the
 * chain must call it (model code is warp/group-converged). */
__device__ __forceinline__ void mcmcb_sync_lanes(const mcmcb_ctx& c) {
  if (c.bar_threads > 32) asm volatile("bar.sync %0, %1;" ::"r"(c.bar_id), "r"(c.bar_threads) : "memory");
  else __syncwarp();
}

/* ---------------------------------------------------------------------------------------
 * mcmcb_exp: FP64 exp() for model code, built for the FP64 pipe of sm_100a.
 * exp(a) = 2^m * 2^(j/2048) * e^u with k = round(a*2048/ln2) = 2048 m + j and |u| <= ln2/4096, so a
 * degree-2 polynomial for (e^u - 1)/u reaches double precision (near-minimax, relative error of e^u
 * 8.5e-18 = 0.08 ulp).  2^(j/2048) comes from a 2048-entry table (16 KB) that the sampling kernels
 * stage in shared memory.  The fast paths are branch-free (independent calls interleave in the
 * instruction stream).
 *
 *   mcmcb_exp_fast(a)         8 FP64 instructions  (libdevice exp(): 16)
 *   mcmcb_expmul_fast(x, ks)  7 FP64 instructions  = exp(x * s) with ks = mcmcb_expmul_scale(s)
 *                             computed once per evaluation: the product, the range reduction
 *                             and the change of units to ln2/2048 are one DFMA pair.
 *
 * An FP64 warp instruction occupies the scheduler's issue port for two cycles on sm_100a, so
 * the instruction COUNT of the datum loop, integer and load instructions included, is what
 * sets the speed of a model evaluation: a table 8 times larger than the earlier
 * 256-entry one (which was replicated 16 times to be free of bank conflicts) buys one Horner step;
 * the lookups now conflict (about 3 wavefronts per half-warp for random j), which costs shared-memory
 * pipe cycles but no issue slots.  mcmcb_exp_ok() tells whether an argument is in the fast range
 * (|a| < 708: result normal, no overflow, not NaN).
 * ------------------------------------------------------------------------------------- */
#define MCMCB_EXP_TAB_N 2048
#ifndef MCMCB_EXP_TAB_REP
#define MCMCB_EXP_TAB_REP 1 /* copies of every entry; lane l reads copy l mod REP (fewer bank conflicts, more shared memory) */
#endif
#define MCMCB_EXP_TAB_DOUBLES (MCMCB_EXP_TAB_N * MCMCB_EXP_TAB_REP)
#define MCMCB_EXP_TAB_SHIFT 9 /* 2^20 / MCMCB_EXP_TAB_N: k << 9 puts m = k >> 11 at the exponent field */

/* coefficients live in the constant bank so that DFMA/DMUL read them as c[][] operands instead
 * of re-materialising 64-bit immediates inside the loop (scripts/gen_exp_table.py prints them) */
__constant__ double MCMCB_EXPC[10] = {
    0x1.71547652b82fep+11,  /* [0] 2048/ln2 */
    -0x1.62e42fec00000p-12, /* [1] -ln2/2048, high part (22 trailing zero bits: k*hi is exact) */
    -0x1.d1cf79abc9e3bp-43, /* [2] -ln2/2048, low part */
    /* (e^u - 1)/u = a0 + a1 u + a2 u^2 on |u| <= ln2/4096, interpolated at the Chebyshev nodes */
    0x1.5555555b7bae9p-3,   /* [3] a2 */
    0x1.00000007afef8p-1,   /* [4] a1 */
    /* the same polynomial in r = u/L, L = ln2/2048, r in [-1/2, 1/2] */
    0x1.62e42fefa39efp-12,  /* [5] a0 L */
    0x1.ebfbe006f2598p-25,  /* [6] a1 L^2 */
    0x1.c6b08d787b3bfp-38,  /* [7] a2 L^3 */
    0.0,                    /* [8] unused */
    1.0};                   /* [9] a0 */

/* [5] and [6] again as macros: the host writes them into the kernel parameters (mcmcb_ctx::exp_c1/c2) */
#define MCMCB_EXP_C1L 0x1.ebfbe006f2598p-25 /* a1 L^2 */
#define MCMCB_EXP_C2L 0x1.c6b08d787b3bfp-38 /* a2 L^3 */

#define MCMCB_EXP_MAGIC 6755399441055744.0 /* 1.5 * 2^52: rounds to integer, k in the low word */

/* 2^(k/2048) * (1 + s) from the reduced pieces.  Three integer instructions: mask, address, and
 * ONE multiply-add for the exponent: table entry j is stored with j*2^9 subtracted from its high
 * word, so that adding k*2^9 = m*2^20 + j*2^9 to it yields the high word of 2^m * 2^(j/2048)
 * without isolating m.  `tl` is the shared-window address of the table; the load is spelled as
 * ld.shared so that the address stays one mask and one multiply-add. */
__device__ __forceinline__ double mcmcb_exp_assemble(int k, double s, unsigned tl) {
  double tj;
  unsigned addr;
  asm("mad.lo.u32 %0, %1, %3, %2;" : "=r"(addr) : "r"(k & (MCMCB_EXP_TAB_N - 1)), "r"(tl), "n"(8 * MCMCB_EXP_TAB_REP));
  asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(addr));
  const double sc = __hiloint2double((k << MCMCB_EXP_TAB_SHIFT) + __double2hiint(tj), __double2loint(tj));
  return fma(sc, s, sc);
}
/* shared-window address of a table staged at smem_tab */
__device__ __forceinline__ unsigned mcmcb_exp_column(const double* smem_tab) {
  unsigned a = (unsigned)__cvta_generic_to_shared(smem_tab) + 8u * (threadIdx.x & (MCMCB_EXP_TAB_REP - 1));
  asm volatile("" : "+r"(a));  /* opaque: one register base for every lookup */
  return a;
}

__device__ __forceinline__ double mcmcb_exp_fast(double a, unsigned tl) {
  double t = fma(a, MCMCB_EXPC[0], MCMCB_EXP_MAGIC);
  const int k = __double2loint(t);
  t -= MCMCB_EXP_MAGIC;
  double r = fma(t, MCMCB_EXPC[1], a);
  r = fma(t, MCMCB_EXPC[2], r);
  double q = fma(r, MCMCB_EXPC[3], MCMCB_EXPC[4]);
  q = fma(r, q, MCMCB_EXPC[9]);
  return mcmcb_exp_assemble(k, r * q, tl);
}

/* ks for mcmcb_expmul_fast: s * 2048/ln2 */
__device__ __forceinline__ double mcmcb_expmul_scale(double s) { return s * MCMCB_EXPC[0]; }

/* exp(x * s), ks = mcmcb_expmul_scale(s).  The reduced argument r = x*ks - round(x*ks) is formed
 * by one DFMA (exact product, one rounding of a value <= 1/2), so the only argument error is the
 * rounding of ks itself: |x s| * 2^-52 relative in the result, the same size as the rounding of
 * the product x*s inside exp(x*s).
 *
 * Operand placement matters: measured on B200 (scripts/ubench_fp64_operands.cu) a DFMA that reads
 * three DISTINCT 64-bit registers holds the FP64 pipe for 3 cycles, one that reads at most two
 * (plus a uniform-register / constant / immediate operand, or a repeated register) for 2.  The
 * Horner steps below are (r, C3, C2) and (r, q, C1): an instruction takes ONE uniform operand, so C2
 * has to sit in a vector register and C3, C1 in uniform registers.  ptxas 12.9 does that when C2
 * (`c1` below = MCMCB_EXP_C1L = a1 L^2) arrives through the kernel-parameter bank (c[0x0], mcmcb_ctx)
 * and C3, C1 come from the __constant__ bank (c[0x3]); other splits put two of them in vector
 * registers and cost a cycle per step (checked in the SASS: `DFMA R, R, UR, R` then `DFMA R, R, R, UR`). */
__device__ __forceinline__ double mcmcb_expmul_fast(double x, double ks, unsigned tl, double c1, double c2) {
  const double t = fma(x, ks, MCMCB_EXP_MAGIC);
  const int k = __double2loint(t);
  /* (-(double)k is the same value; I2F.F64 measured exactly as expensive as this DADD on B200) */
  const double r = fma(x, ks, MCMCB_EXP_MAGIC - t);
  double q = fma(r, MCMCB_EXPC[7], c1);
  q = fma(r, q, MCMCB_EXPC[5]);
  (void)c2;
  return mcmcb_exp_assemble(k, r * q, tl);
}
__device__ __forceinline__ double mcmcb_expmul_fast(double x, double ks, unsigned tl) {
  return mcmcb_expmul_fast(x, ks, tl, MCMCB_EXPC[6], MCMCB_EXPC[7]);
}

/* exp(x * s) for arguments known to lie in [-(dn-2) ln2/2048, 0]: the DIRECT table holds 2^(k/2048) itself for
 * every k the argument range can produce (k = -(dn-1)..0), so the lookup needs neither the mask that isolates j nor the
 * integer multiply-add that inserts the exponent m -- two non-FP64 instructions (address, load) instead of four in a
 * loop whose cost is 2 x (FP64 instructions) + (other instructions) issue cycles.  2^m * 2^(j/2048) is an
 * exact scaling of the correctly rounded table entry, so the result is bit-identical to mcmcb_expmul_fast's.  `td` is the
 * shared-window address of entry k = 0; entries of negative k sit below it.
 * MCMCB_EXP_DIRECT_SPLIT: the entry's high and low words live in two 4-byte tables (low words dn*4 bytes above the
 * high words): a warp's lookups then alias only when two k differ by a multiple of 32 (8-byte entries: 16). */
__device__ __forceinline__ double mcmcb_expmul_direct(double x, double ks, unsigned td, double c1, int dn) {
  const double t = fma(x, ks, MCMCB_EXP_MAGIC);
  const int k = __double2loint(t);
  const double r = fma(x, ks, MCMCB_EXP_MAGIC - t);
  double q = fma(r, MCMCB_EXPC[7], c1);
  q = fma(r, q, MCMCB_EXPC[5]);
  unsigned addr;
#ifdef MCMCB_EXP_DIRECT_SPLIT
  int hi, lo;
  asm("mad.lo.s32 %0, %1, 4, %2;" : "=r"(addr) : "r"(k), "r"(td));
  asm("ld.shared.b32 %0, [%1];" : "=r"(hi) : "r"(addr));
  asm("ld.shared.b32 %0, [%1];" : "=r"(lo) : "r"(addr + 4u * (unsigned)dn));
  const double sc = __hiloint2double(hi, lo);
#else
  double sc;
  (void)dn;
  asm("mad.lo.s32 %0, %1, 8, %2;" : "=r"(addr) : "r"(k), "r"(td));
  asm("ld.shared.f64 %0, [%1];" : "=d"(sc) : "r"(addr));
#endif
  return fma(sc, r * q, sc);
}
/* does exp(x*s) for every |x| <= xmax of one sign stay inside a direct table of dn entries?  (s*x <= 0 is the
 * caller's to guarantee; two entries of slack for the rounding of k) */
__device__ __forceinline__ bool mcmcb_exp_direct_ok(double s, double xmax, int dn) {
  return fabs(s) * xmax * MCMCB_EXPC[0] < (double)(dn - 2);
}

/* true when the fast paths are valid for argument a: |a| < 708 and a is not NaN */
__device__ __forceinline__ bool mcmcb_exp_ok(double a) {
  return (unsigned)(__double2hiint(a) & 0x7fffffff) < 0x40862000u;
}
__device__ __forceinline__ double mcmcb_exp(double a, const mcmcb_ctx& c) {
  if (c.exp_tl != 0u && mcmcb_exp_ok(a)) return mcmcb_exp_fast(a, c.exp_tl);
  return exp(a);
}

/* correctly rounded 2^(j/2048) (scripts/gen_exp_table.py, mpmath at 80 digits) */
__device__ static const double MCMCB_EXP2_TABLE[MCMCB_EXP_TAB_N] = {
#include "mcmcb200_exp_table.inc"
};

/* stage the table; call from every thread of the CTA, then __syncthreads() */
__device__ __forceinline__ void mcmcb_stage_exp_table(double* smem_tab) {
  for (int i = threadIdx.x; i < MCMCB_EXP_TAB_DOUBLES; i += blockDim.x) {
    const int j = i / MCMCB_EXP_TAB_REP;
    const double v = MCMCB_EXP2_TABLE[j];
    smem_tab[i] = __hiloint2double(__double2hiint(v) - (j << MCMCB_EXP_TAB_SHIFT), __double2loint(v));  /* see mcmcb_exp_assemble */
  }
}

/* stage the direct table 2^(k/2048), k = -(dn-1)..0, at smem_direct (entry e holds k = e-(dn-1)); call from every
 * thread of the CTA, then __syncthreads().  Returns nothing; mcmcb_exp_direct_base() gives the address of entry k = 0. */
__device__ __forceinline__ void mcmcb_stage_exp_direct(double* smem_direct, int dn) {
  for (int e = threadIdx.x; e < dn; e += blockDim.x) {
    const int k = e - (dn - 1);
    const int j = k & (MCMCB_EXP_TAB_N - 1), m = k >> 11;  /* k = 2048 m + j, m <= 0 */
    const double v = MCMCB_EXP2_TABLE[j];
    const int hi = __double2hiint(v) + m * (1 << 20), lo = __double2loint(v);
#ifdef MCMCB_EXP_DIRECT_SPLIT
    reinterpret_cast<int*>(smem_direct)[e] = hi;
    reinterpret_cast<int*>(smem_direct)[dn + e] = lo;
#else
    smem_direct[e] = __hiloint2double(hi, lo);
#endif
  }
}
__device__ __forceinline__ unsigned mcmcb_exp_direct_base(const double* smem_direct, int dn) {
#ifdef MCMCB_EXP_DIRECT_SPLIT
  unsigned a = (unsigned)__cvta_generic_to_shared(smem_direct) + 4u * (unsigned)(dn - 1);
#else
  unsigned a = (unsigned)__cvta_generic_to_shared(smem_direct) + 8u * (unsigned)(dn - 1);
#endif
  asm volatile("" : "+r"(a));
  return a;
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
