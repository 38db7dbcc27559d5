"""polaroid_b200 — B200-native filter -> group_by -> agg / group_by_dynamic for Polarway.

The compute path is hand-written sm_100a CUDA behind a C ABI (include/polarway_b200.h,
polaroid_b200/csrc).  This package is the host-side mirror of the reference's LazyFrame API
for that path.  There is no CPU fallback: collecting without the CUDA library raises.
"""
from .plan import LazyFrame, col, len_ as len, sum_ as sum, count_ as count  # noqa: A001,F401

__all__ = ["LazyFrame", "col", "len", "sum", "count"]
