/*
 * mcmcb200_plugin.cuh -- build a user model into a plugin library.
 *
 * In the reference a user supplies `ssfunction`, `priorfun` and `checkbounds` as object files that
 * win over the library's defaults at link time (external_inc.h:4-28, Makefile:92-95).  Here the user
 * writes a model struct (include/mcmcb200_model.cuh), compiles ONE .cu against this header into a
 * shared library, and the sampler finds the model by name:
 *
 *     #include "mcmcb200_plugin.cuh"
 *     struct MyModel { static constexpr int NPAR = 3, NY = 1; static const char* name() { return "mymodel"; } ... };
 *     MCMCB_REGISTER_MODEL_K1(MyModel)      // NPAR > 0: register (thread/lane-group per chain) kernel
 *     // or MCMCB_REGISTER_MODEL_K2(MyModel)   NPAR == 0: run-time npar, warp-per-chain kernels (DRAM/AM, RAM, SCAM)
 *
 *     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC \
 *          -I<repo>/include -I<repo>/mcmcf90_b200/csrc mymodel.cu -L<repo>/mcmcf90_b200 -lmcmcb200 -o libmymodel.so
 *
 * The host loads it with mcmcb_load_plugin("libmymodel.so") (or dlopen / ctypes.CDLL) before
 * mcmcb_create with cfg.model = "mymodel".  The sampling kernels are instantiated with the model inside
 * the plugin, so its ssfunction is inlined into the hot loop exactly like a built-in model's.
 */
#ifndef MCMCB200_PLUGIN_CUH
#define MCMCB200_PLUGIN_CUH

#include "launchers.cuh" /* mcmcf90_b200/csrc */

#define MCMCB_REGISTER_MODEL_K1(Model)                                                                   \
  namespace {                                                                                            \
  struct Model##_mcmcb_registrar_k1 {                                                                    \
    Model##_mcmcb_registrar_k1() { mcmcb::register_model(mcmcb::launch::K1<Model>::entry()); }           \
  } Model##_mcmcb_registrar_k1_instance;                                                                 \
  }

#define MCMCB_REGISTER_MODEL_K2(Model)                                                                   \
  namespace {                                                                                            \
  struct Model##_mcmcb_registrar_k2 {                                                                    \
    Model##_mcmcb_registrar_k2() { mcmcb::register_model(mcmcb::launch::K2<Model>::entry()); }           \
  } Model##_mcmcb_registrar_k2_instance;                                                                 \
  }

#endif
