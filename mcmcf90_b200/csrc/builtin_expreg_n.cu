// Built-in user model ExpRegN ("expreg" on the warp-per-chain kernels: SCAM and SVD-factor samplers of the shipped
// testcase), registered the way a user plugin registers a model (include/mcmcb200_plugin.cuh).
#include "models.cuh"
#include "mcmcb200_plugin.cuh"

using mcmcb::ExpRegN;
MCMCB_REGISTER_MODEL_K2(ExpRegN)
