"""mcmcf90_b200 -- B200-native batched adaptive Metropolis-Hastings (mcmcf90's sampling hot path).

The product is the CUDA library `libmcmcb200.so` behind the C ABI in include/mcmcb200.h;
this package is the thin Python binding used by tests and bench.py.  There is no CPU
fallback: importing the binding without the built library raises.
"""
from .binding import (Config, Sampler, MCMCBError, load_library, library_path, default_config, dfma_peak, exp_selftest, load_plugin,
                      DRAM, RAM, SCAM, ER, RNG_PHILOX, RNG_INJECTED)
from . import models

__all__ = ["Config", "Sampler", "MCMCBError", "load_library", "library_path", "default_config", "dfma_peak", "exp_selftest", "load_plugin",
           "models", "DRAM", "RAM", "SCAM", "RNG_PHILOX", "RNG_INJECTED"]
