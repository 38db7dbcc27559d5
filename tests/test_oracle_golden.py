"""Pins the CPU oracle against the reference's own known-answer tests (tests/golden/vectors.json,
each case cites the reference file:line it was transcribed from)."""
import numpy as np
import pytest

from oracle import oracle
from tests import golden_util as G

CASES = G.load_cases({"group_by", "dynamic", "dynamic_total"})
WINDOWS = G.load_cases({"windows"})


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_vectors(case):
    G.run_case(case, lambda q: oracle.collect(q))


@pytest.mark.parametrize("case", [c for c in CASES if c["kind"] == "group_by"], ids=lambda c: c["name"])
def test_oracle_multithreaded_matches(case):
    # thread-local tables + merge (the streaming engine's structure) must give the same answer
    G.run_case(case, lambda q: oracle.collect(q, n_threads=3))


@pytest.mark.parametrize("case", WINDOWS, ids=[c["name"] for c in WINDOWS])
def test_oracle_group_by_windows(case):
    s, l, lo, up = oracle.group_by_windows(np.array(case["time"], dtype=np.int64), case["every"], case["period"],
                                           case["offset"], case["closed"])
    groups = [[int(a), int(b)] for a, b in zip(s, l)]
    if "groups" in case:
        assert groups == case["groups"]
    if "groups_prefix" in case:
        assert groups[:len(case["groups_prefix"])] == case["groups_prefix"]
    if "lower_prefix" in case:
        assert list(lo[:len(case["lower_prefix"])]) == case["lower_prefix"]
        assert list(up[:len(case["upper_prefix"])]) == case["upper_prefix"]
