"""microclimate.jl_b200 — B200-native batched-sites engine for Microclimate.jl's soil path.

Host-side mirror of the reference interface for `solve_soil!` (src/simulation.jl:355-819) over a
C-ABI CUDA library (include/microclimate_b200.h).  Import as `microclimate_jl_b200` (the directory
name carries the reference's dot; `microclimate_jl_b200.py` at the repo root is the import shim).
"""
from . import _abi, types, marshal, weather
from .types import *  # noqa: F401,F403  (reference-shaped public types)

__all__ = ["_abi", "types", "marshal", "weather"]


def __getattr__(name):
    # engine/simulation need the CUDA library; import lazily so type-only users work without it
    if name in ("SoilEngine", "MultiSoilEngine", "solve_soil_batched", "solve_soil", "solve_minmax_batched", "solve"):
        from . import simulation
        return getattr(simulation, name)
    raise AttributeError(name)
