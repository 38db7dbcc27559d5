"""The bucket tier with ids from a shared-memory key index (pw_bucket.cuh, strategy 9): low-cardinality keys that do not
span a small integer range — sparse integers, short strings, several key columns, nullable keys.  The pilot chooses it
on inputs above 2^18 rows; CUDA vs the oracle on the same seeded inputs, and the per-cell hot table (FLAG_NO_BUCKETS)
must give the same answer.  Bit-exact integers / counts / min / max / first / last, f64 sums 1e-12."""
import numpy as np
import pyarrow as pa
import pytest

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import engine
from tests import golden_util as G

pytestmark = pytest.mark.gpu

N = 700_000


def run_both(q, sort_by, expect=9):
    want = oracle.collect(q)
    got = engine.run_group_by(q.table, q.plan)
    if expect is not None:
        assert engine.last_timings()["strategy"] == expect, engine.last_timings()["strategy"]
    G.assert_tables_equal(got, want, sort_by=sort_by, rtol=1e-12)
    got2 = engine.run_group_by(q.table, q.plan, flags=engine.FLAG_NO_BUCKETS)
    assert engine.last_timings()["strategy"] not in (7, 9)
    G.assert_tables_equal(got2, want, sort_by=sort_by, rtol=1e-12)
    return got


AGGS = lambda: [pw.col("v").sum().alias("sum"), pw.col("v").mean().alias("mean"), pw.col("v").min().alias("min"),
                pw.col("v").max().alias("max"), pw.col("v").count().alias("count"), pw.col("v").first().alias("first"),
                pw.col("v").last().alias("last"), pw.len().alias("len")]


@pytest.mark.parametrize("groups", [60, 500, 1500])
def test_sparse_integer_keys(groups):
    rng = np.random.default_rng(71)
    t = pa.table({"k": pa.array(rng.integers(0, groups, N) * 1_000_003 - 7_000_000_000),
                  "v": pa.array(rng.normal(20.0, 3.0, N), mask=rng.random(N) < 0.05)})
    # (whether two value planes x the bucket depth fit next to the index depends on the group count: the planner may
    # keep the per-cell table; the single-plane shape below is the one that must take the tier)
    run_both(pw.LazyFrame(t).group_by("k").agg(*AGGS()), ["k"], expect=None)
    q1 = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("sum"), pw.col("v").min().alias("min"), pw.col("v").max().alias("max"),
                                           pw.col("v").count().alias("count"), pw.len().alias("len"))
    run_both(q1, ["k"], expect=None)


def test_short_string_keys_with_nulls():
    rng = np.random.default_rng(72)
    names = np.array([f"sym{i:04d}" for i in range(800)])
    t = pa.table({"k": pa.array(names[rng.integers(0, 800, N)], mask=rng.random(N) < 0.01), "v": pa.array(rng.integers(-500, 500, N))})
    got = run_both(pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("sum"), pw.col("v").min().alias("min"),
                                                     pw.col("v").max().alias("max"), pw.len().alias("len")), ["k"])   # one value plane: must take the tier
    assert got.num_rows == 801
    run_both(pw.LazyFrame(t).group_by("k").agg(*AGGS()), ["k"], expect=None)


def test_two_key_columns_maintain_order():
    rng = np.random.default_rng(73)
    t = pa.table({"a": pa.array(rng.integers(0, 30, N).astype(np.int32), mask=rng.random(N) < 0.02), "b": pa.array(rng.integers(0, 25, N) * 97),
                  "v": pa.array(rng.random(N) * 10.0), "w": pa.array(rng.integers(0, 100, N))})
    q = pw.LazyFrame(t).filter(pw.col("w") < 80).group_by("a", "b", maintain_order=True).agg(
        pw.col("v").sum().alias("s"), pw.col("v").max().alias("hi"), pw.col("w").min().alias("lo"), pw.len().alias("n"))
    run_both(q, None, expect=None)


def test_keys_equal_to_the_table_sentinels_and_null_keys():
    rng = np.random.default_rng(74)
    pool = np.concatenate([np.array([-1, -2, 0, 2**62, -2**63]), rng.integers(-10**15, 10**15, 300)])
    t = pa.table({"k": pa.array(pool[rng.integers(0, len(pool), N)], mask=rng.random(N) < 0.03), "v": pa.array(rng.integers(0, 1000, N))})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("s"), pw.col("v").min().alias("lo"), pw.col("v").max().alias("hi"), pw.len().alias("n"))
    run_both(q, ["k"])


def test_more_groups_than_the_index_was_sized_for():
    # the contiguous pilot block sees ~600 keys; the second half of the input brings 3000 more: the index fills up and the
    # surplus keys aggregate through the HBM table
    rng = np.random.default_rng(75)
    k = np.concatenate([rng.integers(0, 600, N // 2), rng.integers(0, 3600, N - N // 2)]) * 13_000_001
    t = pa.table({"k": pa.array(k), "v": pa.array(rng.normal(5.0, 1.0, N))})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("s"), pw.col("v").min().alias("lo"), pw.len().alias("n"))
    want = oracle.collect(q)
    got = engine.run_group_by(q.table, q.plan)
    G.assert_tables_equal(got, want, sort_by=["k"], rtol=1e-12)
