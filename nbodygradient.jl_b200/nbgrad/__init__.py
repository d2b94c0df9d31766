"""nbgrad — host-side mirror of NbodyGradient.jl's driver API for the batched AHL21 + Jacobian + transit-timing
hot path, executing in libnbgrad_b200.so (hand-written sm_100a kernels behind the C ABI of include/nbgrad.h)."""
from .ics import (Elements, ElementsIC, CartesianIC, get_default_ICs, available_systems, init_nbody_elements, nested_hierarchy,
                  trappist1_elements, GNEWT, YEAR)
from .integrator import State, dState, Integrator, TransitTiming, TransitParameters, CartesianOutput, ElementsOutput, ahl21, check_step, device_count, release_plans
from ._lib import NbgError, lib, SYMBOLS
from .sharding import shard_range, shard_counts, max_over_ranks, sum_over_ranks, gather_slices

__all__ = ["Elements", "ElementsIC", "CartesianIC", "get_default_ICs", "available_systems", "State", "dState", "Integrator", "TransitTiming",
           "TransitParameters", "CartesianOutput", "ElementsOutput", "ahl21", "NbgError"]
