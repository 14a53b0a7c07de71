"""jaqmc_b200 -- B200-native (sm_100a) implementation of JaQMC's local-energy + sampling hot path.

Host-side mirror of the reference's wavefunction / sampler / estimator protocols over hand-written CUDA
kernels reached through the C ABI in ``include/jaqmc_b200.h``.  Importing the package does not load the
CUDA library; the first compute call does, and fails loudly if it is missing (there is no CPU fallback).
"""

__version__ = "0.1.0"
