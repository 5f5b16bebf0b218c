// Built-in user model GaussK<6>: the Gaussian target of testcases/mcmcrun4.F90 with a compile-time npar = 6 for the
// register kernel, registered like a user plugin registers a model (include/mcmcb200_plugin.cuh).
#include "models.cuh"
#include "mcmcb200_plugin.cuh"

using GaussK6 = mcmcb::GaussK<6>;
MCMCB_REGISTER_MODEL_K1(GaussK6)
