// Built-in user model BananaN: its own translation unit, registered the way a user plugin registers a model
// (include/mcmcb200_plugin.cuh) -- the run-time form of the reference's link-time ssfunction override
// (external_inc.h:4-28).
#include "models.cuh"
#include "mcmcb200_plugin.cuh"

using mcmcb::BananaN;
MCMCB_REGISTER_MODEL_K2(BananaN)
