"""Dense-id hot table (a single integer key whose values span a small range: id = key - min, no key index).
The path is chosen from the pilot's key range on large inputs and, for unit-sized inputs, under FLAG_FORCE_HOT.
Rows whose key lies outside the sampled range, null keys and keys that alias the table sentinels (-1, -2 as
int64) must still aggregate — through the HBM table.  Bit-exact / 1e-12 as in test_gpu_parity.py."""
import numpy as np
import pyarrow as pa
import pytest

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import engine
from tests import golden_util as G

pytestmark = pytest.mark.gpu

AGGS = lambda: [pw.col("v").sum().alias("sum"), pw.col("v").mean().alias("mean"), pw.col("v").min().alias("min"),
                pw.col("v").max().alias("max"), pw.col("v").count().alias("count"), pw.col("v").first().alias("first"),
                pw.col("v").last().alias("last"), pw.len().alias("len")]


def run(q, sort_by, expect_dense=True, **opts):
    """Dense ids run through the bucket tier (strategy 7: rows bucketed per tile, accumulators in registers) when the
    shape allows it and through the per-cell hot table (strategy 4) otherwise; both are checked."""
    want = oracle.collect(q)
    got = engine.run_group_by(q.table, q.plan, **opts)
    first = engine.last_timings()["strategy"]
    assert (first in (4, 7)) == expect_dense
    G.assert_tables_equal(got, want, sort_by=sort_by, rtol=1e-12)
    if first == 7:
        o2 = dict(opts)
        o2["flags"] = o2.get("flags", 0) | engine.FLAG_NO_BUCKETS
        got2 = engine.run_group_by(q.table, q.plan, **o2)
        assert engine.last_timings()["strategy"] in (4, 1)   # 1: too many accumulators for the dense per-cell table
        G.assert_tables_equal(got2, want, sort_by=sort_by, rtol=1e-12)
    return got


@pytest.mark.parametrize("maintain_order", [False, True])
def test_dense_small_input_with_sentinel_aliases_and_null_keys(maintain_order):
    rng = np.random.default_rng(21)
    n = 60_000
    keys = rng.integers(-5, 300, n)            # includes -1 and -2 (KEY_EMPTY / KEY_NULL bit patterns)
    t = pa.table({"k": pa.array(keys, mask=rng.random(n) < 0.01),
                  "v": pa.array(rng.normal(size=n), mask=rng.random(n) < 0.05)})
    q = pw.LazyFrame(t).group_by("k", maintain_order=maintain_order).agg(*AGGS())
    run(q, None if maintain_order else ["k"], flags=engine.FLAG_FORCE_HOT)


@pytest.mark.parametrize("dtype", ["int8", "uint8", "int16", "uint16", "int32", "uint32", "int64"])
def test_dense_every_integer_key_dtype(dtype):
    rng = np.random.default_rng(22)
    n = 40_000
    lo = -60 if dtype.startswith("int") else 3
    t = pa.table({"k": pa.array(rng.integers(lo, lo + 120, n).astype(dtype), mask=rng.random(n) < 0.02),
                  "v": pa.array(rng.integers(-1000, 1000, n))})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("s"), pw.col("v").min().alias("lo"),
                                          pw.col("v").max().alias("hi"), pw.len().alias("n"))
    run(q, ["k"], flags=engine.FLAG_FORCE_HOT)


def test_dense_large_input_with_outliers_the_sample_misses():
    # 1.5e6 rows (> the pilot threshold): keys 0..999 dense; a handful of far-away keys and null keys sit at rows
    # the strided/contiguous pilots never read -> they must arrive through the spill tier
    rng = np.random.default_rng(23)
    n = 1_500_000
    keys = rng.integers(0, 1000, n)
    for r in (7, 100_001, 1_499_999):
        keys[r] = 10**12 + r
    mask = np.zeros(n, dtype=bool)
    mask[[3, 999_999]] = True
    t = pa.table({"k": pa.array(keys, mask=mask), "v": pa.array(rng.random(n) * 100.0)})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("sum"), pw.col("v").mean().alias("mean"),
                                          pw.col("v").min().alias("min"), pw.col("v").max().alias("max"))
    got = run(q, ["k"])
    assert got.num_rows == 1000 + 3 + 1
    assert engine.last_timings()["spilled_rows"] == 5


def test_sparse_range_keeps_the_hash_index():
    rng = np.random.default_rng(24)
    n = 1_000_000
    t = pa.table({"k": pa.array(rng.integers(0, 500, n) * 7919), "v": pa.array(rng.random(n))})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("sum"), pw.len().alias("n"))
    run(q, ["k"], expect_dense=False)


def test_dense_with_filter_and_no_len_falls_back():
    rng = np.random.default_rng(25)
    n = 50_000
    t = pa.table({"k": pa.array(rng.integers(0, 64, n)), "p": pa.array(rng.integers(0, 10, n)),
                  "v": pa.array(rng.integers(0, 9, n), mask=rng.random(n) < 0.1)})
    # nullable value + sum only -> no per-group row counter -> hash index
    q = pw.LazyFrame(t).filter(pw.col("p") < 7).group_by("k").agg(pw.col("v").sum().alias("s"))
    run(q, ["k"], expect_dense=False, flags=engine.FLAG_FORCE_HOT)
    # with len the dense path applies; the predicate leaves some ids of the range empty
    t2 = t.set_column(1, "p", pa.array(np.where(t["k"].to_numpy() % 5 == 0, 9, t["p"].to_numpy())))
    q2 = pw.LazyFrame(t2).filter(pw.col("p") < 7).group_by("k").agg(pw.col("v").sum().alias("s"), pw.len().alias("n"))
    got = run(q2, ["k"], flags=engine.FLAG_FORCE_HOT)
    assert got.num_rows == 64 - 13


@pytest.mark.parametrize("hot_share", [0.02, 0.3])
def test_dense_skewed_keys_fill_the_overflow_list(hot_share):
    # one id takes a large share of the rows: its bucket overflows in every tile, the CTA-wide overflow list fills up
    # and the rest goes to the HBM table — all three routes must add up exactly
    rng = np.random.default_rng(26)
    n = 1_200_000
    keys = rng.integers(0, 1000, n)
    keys[rng.random(n) < hot_share] = 417
    v = rng.integers(-10**6, 10**6, n)
    t = pa.table({"k": pa.array(keys), "v": pa.array(v), "f": pa.array(rng.normal(10.0, 1.0, n))})   # positive: no sum cancels to ~0
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("sum"), pw.col("v").min().alias("min"), pw.col("v").max().alias("max"),
                                          pw.col("f").sum().alias("fsum"), pw.col("f").min().alias("fmin"), pw.col("f").max().alias("fmax"),
                                          pw.len().alias("n"))
    run(q, ["k"])


@pytest.mark.parametrize("maintain_order", [False, True])
def test_dense_first_last_of_non_nullable_values_use_row_positions(maintain_order):
    # no value can be null: first / last / first-occurrence words come from two row positions per id and tile instead
    # of a meta word per row (ROWPOS, pw_bucket.cuh); rows that overflow a bucket or take the HBM path still count
    rng = np.random.default_rng(27)
    n = 900_000
    keys = rng.integers(0, 700, n)
    keys[rng.random(n) < 0.05] = 13            # a hot id: overflow list and HBM path in every tile
    t = pa.table({"k": pa.array(keys), "v": pa.array(rng.integers(-10**6, 10**6, n)), "f": pa.array(rng.normal(3.0, 1.0, n))})
    q = pw.LazyFrame(t).group_by("k", maintain_order=maintain_order).agg(
        pw.col("v").first().alias("v_first"), pw.col("v").last().alias("v_last"), pw.col("f").first(ignore_nulls=True).alias("f_first"),
        pw.col("f").last().alias("f_last"), pw.col("v").sum().alias("v_sum"), pw.col("f").max().alias("f_max"), pw.len().alias("n"))
    run(q, None if maintain_order else ["k"])
