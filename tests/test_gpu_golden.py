"""CUDA path vs the reference's own known-answer tests (tests/golden/vectors.json), through the C ABI.
Every case runs under each table strategy (hot table + spill tier / HBM table only), the analogue of
the reference re-running its suite with POLARS_FORCE_PARTITION / engine swaps."""
import pytest

from polaroid_b200 import engine
from tests import golden_util as G

pytestmark = pytest.mark.gpu

CASES = G.load_cases({"group_by", "dynamic", "dynamic_total"})
STRATEGIES = {
    "auto": {},
    "hot": {"flags": engine.FLAG_FORCE_HOT},
    "hot_tiny": {"flags": engine.FLAG_FORCE_HOT, "hot_table_slots": 16},   # forces evictions + spill tier
    "global": {"flags": engine.FLAG_FORCE_GLOBAL},
    "global_tiny": {"flags": engine.FLAG_FORCE_GLOBAL, "initial_table_slots": 4},  # forces table growth retries
    "hash_dynamic": {"flags": engine.FLAG_NO_SEGMENTED},  # keys-less tumbling windows through the hash path
    "window_list": {"flags": engine.FLAG_FORCE_SEGMENTED},  # keys-less tumbling windows through the general window-list path
}


@pytest.mark.parametrize("strategy", list(STRATEGIES))
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_matches_reference_vectors(case, strategy):
    opts = STRATEGIES[strategy]
    G.run_case(case, lambda q: engine.run_group_by(q.table, q.plan, **opts))
