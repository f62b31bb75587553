"""volatilespace_b200 -- B200 (sm_100a) implementation of VolatileSpace's physics hot path.

Modules mirror the reference's volatilespace/physics package:
  phys_editor   Physics (parent/COI model) and AllPairsPhysics (all-pairs model)
  phys_body     game-mode body propagation
  phys_vessel   game-mode vessel propagation (move / curve / curve_move)
  phys_shared, enhanced_kepler_solver   batched free functions
Everything computes in libvsb200.so (include/vsb200.h); importing this package
does not load CUDA, constructing a Device / Physics does and fails loudly
without the library or a GPU.
"""
__version__ = "0.1.0"
