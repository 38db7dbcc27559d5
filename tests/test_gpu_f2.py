"""the other reductions the streaming engine pre-aggregates (crates/polars-expr/src/reduce/convert.rs:46-150):
var / std (reduce/var_std.rs), first / last(ignore_nulls=True) (reduce/first_last_nonnull.rs), null_count (reduce/count.rs),
bitwise and / or / xor (reduce/bitwise.rs), any / all with ignore_nulls (reduce/any_all.rs).
CUDA path vs the CPU oracle on the same seeded inputs.  Integer, Boolean and index results bit-exact; var / std within
1e-9 relative (the device accumulates shifted sums sum(x - c), sum((x - c)^2) in f64 and merges them exactly; the oracle
runs the reference's Welford / Chan recurrence — both are a few ulps from the exact value, not from each other's bits)."""
import numpy as np
import pyarrow as pa
import pytest

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import engine
from tests import golden_util as G

pytestmark = pytest.mark.gpu

STRATEGIES = [{}, {"flags": engine.FLAG_FORCE_HOT}, {"flags": engine.FLAG_FORCE_GLOBAL}]


def check(q, sort_by, rtol=0.0, **opts):
    got = engine.run_group_by(q.table, q.plan, **opts)
    G.assert_tables_equal(got, oracle.collect(q), sort_by=sort_by, rtol=rtol)
    return got


@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("dtype", ["float64", "float32", "int32", "uint16", "int64"])
def test_var_std_every_numeric_class(dtype, strategy):
    rng = np.random.default_rng(41)
    n, groups = 300_000, 700
    x = (rng.normal(1000.0, 3.0, n) if dtype.startswith("float") else rng.integers(0, 5000, n)).astype(dtype)
    t = pa.table({"k": pa.array(rng.integers(0, groups, n)), "x": pa.array(x, mask=rng.random(n) < 0.1)})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("x").var().alias("var1"), pw.col("x").std().alias("std1"),
                                          pw.col("x").var(0).alias("var0"), pw.col("x").std(2).alias("std2"),
                                          pw.col("x").mean().alias("mean"))
    check(q, ["k"], rtol=1e-6 if dtype == "float32" else 1e-9, **strategy)


def test_var_of_small_groups_is_null_when_weight_le_ddof():
    t = pa.table({"k": pa.array([1, 2, 2, 3, 3, 3, 4], type=pa.int64()),
                  "x": pa.array([1.5, 2.0, 4.0, None, None, 7.0, None], type=pa.float64())})
    q = pw.LazyFrame(t).group_by("k", maintain_order=True).agg(pw.col("x").var().alias("v"), pw.col("x").std(0).alias("s0"))
    got = check(q, None, rtol=1e-12)
    assert got.column("v").to_pylist() == [None, 2.0, None, None]
    assert got.column("s0").to_pylist() == [0.0, 1.0, 0.0, None]


@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("maintain_order", [False, True])
def test_first_last_non_null_and_null_count(maintain_order, strategy):
    rng = np.random.default_rng(42)
    n, groups = 200_000, 3000
    t = pa.table({"k": pa.array(rng.integers(0, groups, n).astype(np.int32)),
                  "v": pa.array(rng.integers(-10**9, 10**9, n), mask=rng.random(n) < 0.6),
                  "f": pa.array(rng.normal(size=n), mask=rng.random(n) < 0.3),
                  "nn": pa.array(rng.integers(0, 9, n).astype(np.int16))})
    q = pw.LazyFrame(t).group_by("k", maintain_order=maintain_order).agg(
        pw.col("v").first(ignore_nulls=True).alias("v_first_nn"), pw.col("v").last(ignore_nulls=True).alias("v_last_nn"),
        pw.col("v").first().alias("v_first"), pw.col("v").last().alias("v_last"),
        pw.col("f").first(ignore_nulls=True).alias("f_first_nn"), pw.col("f").last(ignore_nulls=True).alias("f_last_nn"),
        pw.col("v").null_count().alias("v_nulls"), pw.col("f").null_count().alias("f_nulls"), pw.col("nn").null_count().alias("nn_nulls"),
        pw.col("v").count().alias("v_count"), pw.len().alias("len"))
    check(q, None if maintain_order else ["k"], **strategy)


def test_first_non_null_of_all_null_group_is_null():
    t = pa.table({"k": pa.array([1, 1, 2, 2], type=pa.int64()), "v": pa.array([None, None, None, 5], type=pa.int64())})
    q = pw.LazyFrame(t).group_by("k", maintain_order=True).agg(pw.col("v").first(ignore_nulls=True).alias("f"),
                                                               pw.col("v").last(ignore_nulls=True).alias("l"))
    got = check(q, None)
    assert got.column("f").to_pylist() == [None, 5] and got.column("l").to_pylist() == [None, 5]


@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("dtype", ["int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64"])
def test_bitwise_every_integer_dtype(dtype, strategy):
    rng = np.random.default_rng(43)
    n, groups = 100_000, 5000     # ~20 rows per group: AND / OR keep information
    info = np.iinfo(dtype)
    x = rng.integers(info.min, info.max, n, dtype=dtype, endpoint=True)
    t = pa.table({"k": pa.array(rng.integers(0, groups, n)), "x": pa.array(x, mask=rng.random(n) < 0.5)})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("x").bitwise_and().alias("and"), pw.col("x").bitwise_or().alias("or"),
                                          pw.col("x").bitwise_xor().alias("xor"))
    check(q, ["k"], **strategy)


@pytest.mark.parametrize("strategy", STRATEGIES)
def test_boolean_column_any_all_bitwise_sum(strategy):
    rng = np.random.default_rng(44)
    n, groups = 150_000, 20_000
    b = rng.random(n) < 0.7
    t = pa.table({"k": pa.array(rng.integers(0, groups, n)), "b": pa.array(b, mask=rng.random(n) < 0.4),
                  "c": pa.array(rng.random(n) < 0.05)})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("b").any().alias("any_b"), pw.col("b").all().alias("all_b"),
                                          pw.col("c").any().alias("any_c"), pw.col("c").all().alias("all_c"),
                                          pw.col("b").bitwise_and().alias("and_b"), pw.col("b").bitwise_or().alias("or_b"),
                                          pw.col("b").bitwise_xor().alias("xor_b"), pw.col("b").null_count().alias("nulls_b"),
                                          pw.col("b").count().alias("count_b"), pw.col("b").sum().alias("sum_b"))
    check(q, ["k"], **strategy)


def test_f2_reductions_under_group_by_dynamic():
    rng = np.random.default_rng(45)
    n = 50_000
    ts = np.sort(rng.integers(0, 10_000_000, n))
    t = pa.table({"t": pa.array(ts, type=pa.int64()), "x": pa.array(rng.normal(size=n), mask=rng.random(n) < 0.2),
                  "i": pa.array(rng.integers(0, 255, n).astype(np.uint8))})
    q = pw.LazyFrame(t).group_by_dynamic("t", every="1000i").agg(
        pw.col("x").std().alias("std"), pw.col("x").var(0).alias("var0"), pw.col("x").first(ignore_nulls=True).alias("first_nn"),
        pw.col("x").last(ignore_nulls=True).alias("last_nn"), pw.col("x").null_count().alias("nulls"),
        pw.col("i").bitwise_xor().alias("xor"), pw.col("i").bitwise_and().alias("and"))
    check(q, None, rtol=1e-9)
