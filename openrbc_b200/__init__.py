"""openrbc_b200 — B200-native (sm_100a) replacement for OpenRBC's per-timestep force / integrate loop.

The product is the C-ABI shared library `liborbc_b200.so` (include/orbc_b200.h, sources in csrc/).
This package is the thin host-side mirror used by the tests and by bench.py: it loads the library
with ctypes and exposes the reference's hot-path call sequence (src/openrbc.cpp:189-256) as methods
of `Simulation`.  There is no CPU fallback: importing works anywhere, but creating a `Simulation`
without the compiled library or without a CUDA device raises.
"""
from .engine import (Simulation, OrbcError, load_library, build_library, library_path, forcefield_canonical,  # noqa: F401
                     CLEAR_FORCE, POST_TORQUE, BOUNCE_BACK, VERLET_LANGEVIN, NH_INITIAL_FUSED, NH_FINAL_FUSED, NH_FINAL,
                     NH_UPDATE, OPT_MOVE)

__all__ = ["Simulation", "OrbcError", "load_library", "build_library", "library_path", "forcefield_canonical"]
