"""Parity at BASELINE sizes ("no parity evidence at BASELINE sizes"): the full 1e8-row C2 and C3
inputs against the oracle on all host threads.  Tolerances: integer / count / min / max / first / last / group
membership bit-exact, f64 sum and mean relative 1e-12 (north star)."""
import numpy as np
import pyarrow as pa
import pytest

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import engine
from tests import golden_util as G
from tests import synth

pytestmark = pytest.mark.gpu


def test_c2_full_size_1e8_rows_vs_oracle():
    n, groups = 100_000_000, 1_000
    rng = np.random.default_rng(2)
    t = pa.table({"key": rng.integers(0, groups, n, dtype=np.int64), "value": rng.random(n) * 100.0})
    c = pw.col("value")
    q = pw.LazyFrame(t).group_by("key").agg(c.sum().alias("sum"), c.mean().alias("mean"), c.min().alias("min"), c.max().alias("max"))
    want = oracle.collect(q, n_threads=oracle.max_threads())
    frame = engine.DeviceFrame(t)
    try:
        for opts in ({}, {"flags": engine.FLAG_NO_DENSE_IDS}):
            got = frame.group_by(q.plan, **opts)
            G.assert_tables_equal(got, want, sort_by=["key"], rtol=1e-12)
        tm = engine.last_timings()
        assert tm["n_groups"] == groups and tm["retries"] == 0
    finally:
        frame.free()


def test_c3_full_size_1e8_rows_1e7_keys_vs_oracle():
    n, groups = 100_000_000, 10_000_000
    t = synth.c3_table(n, groups, seed=3)
    c = pw.col("value")
    q = pw.LazyFrame(t).group_by("key").agg(c.sum().alias("sum"), c.mean().alias("mean"), c.min().alias("min"), c.max().alias("max"),
                                            c.count().alias("count"), c.first().alias("first"), c.last().alias("last"))
    want = oracle.collect(q, n_threads=min(8, oracle.max_threads()))
    got = engine.run_group_by(q.table, q.plan)
    assert engine.last_timings()["strategy"] == 10   # the radix-partitioned tier (pw_radix.cuh)
    G.assert_tables_equal(got, want, sort_by=["key"], rtol=1e-12)
