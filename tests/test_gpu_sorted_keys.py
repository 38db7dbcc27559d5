"""sorted key columns.  The reference takes the GroupsSlice path — partition_to_groups
(polars-arrow/src/legacy/kernels/sort_partition.rs:168) builds [first, len] per run of equal keys and every aggregate
reduces its slices (polars-core/src/frame/group_by/into_groups.rs:65-129).  Here: the run-combining scan
(pw_runs.cuh, strategy 8) behind LazyFrame.set_sorted, and pw_b200_frame_group_slices for the slices themselves.
CUDA vs the oracle; integer / index results bit-exact, f64 sums 1e-12."""
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


def check(q, expect=8, **opts):
    got = engine.run_group_by(q.table, q.plan, **opts)
    assert engine.last_timings()["strategy"] == expect
    G.assert_tables_equal(got, oracle.collect(q), rtol=1e-12)   # maintain_order: first-occurrence order included
    return got


@pytest.mark.parametrize("run_len", [1, 7, 300, 40_000])
@pytest.mark.parametrize("dtype", ["int64", "int32", "float64"])
def test_sorted_single_key_run_lengths(dtype, run_len):
    rng = np.random.default_rng(51)
    n = 400_000
    keys = np.sort(rng.integers(0, max(1, n // run_len), n)).astype(dtype)
    t = pa.table({"k": pa.array(keys), "v": pa.array(rng.normal(50.0, 5.0, n), mask=rng.random(n) < 0.05)})
    q = pw.LazyFrame(t).set_sorted("k").group_by("k", maintain_order=True).agg(*AGGS())
    check(q)


def test_sorted_with_null_keys_first_and_filter():
    rng = np.random.default_rng(52)
    n = 300_000
    keys = np.sort(rng.integers(0, 2000, n))
    mask = np.zeros(n, dtype=bool)
    mask[:1234] = True                       # nulls first, as a sort places them
    t = pa.table({"k": pa.array(keys, mask=mask), "p": pa.array(rng.integers(0, 10, n)), "v": pa.array(rng.integers(-1000, 1000, n))})
    q = pw.LazyFrame(t).set_sorted("k").filter(pw.col("p") < 6).group_by("k", maintain_order=True).agg(
        pw.col("v").sum().alias("s"), pw.col("v").min().alias("lo"), pw.col("v").max().alias("hi"), pw.len().alias("n"))
    check(q)


def test_sorted_string_and_multi_column_keys():
    rng = np.random.default_rng(53)
    n = 200_000
    a = np.sort(rng.integers(0, 40, n))
    b = np.empty(n, dtype=np.int64)
    for lo in range(0, n, 5000):             # b sorted inside every 5000-row block of a-sorted rows: (a, b) runs of varying length
        b[lo:lo + 5000] = np.sort(rng.integers(0, 30, min(5000, n - lo)))
    names = np.array([f"sym{i:03d}" for i in range(40)])
    t = pa.table({"a": pa.array(names[a]), "b": pa.array(b), "v": pa.array(rng.random(n))})
    q = pw.LazyFrame(t).set_sorted("a").set_sorted("b").group_by("a", "b", maintain_order=True).agg(
        pw.col("v").sum().alias("s"), pw.col("v").max().alias("hi"), pw.len().alias("n"))
    check(q)


def test_a_wrong_promise_costs_time_not_correctness():
    rng = np.random.default_rng(54)
    n = 250_000
    t = pa.table({"k": pa.array(rng.integers(0, 500, n)), "v": pa.array(rng.integers(0, 100, n))})   # NOT sorted
    q = pw.LazyFrame(t).set_sorted("k").group_by("k", maintain_order=True).agg(pw.col("v").sum().alias("s"), pw.len().alias("n"))
    check(q)


def test_high_cardinality_sorted_keys():
    rng = np.random.default_rng(55)
    n = 6_000_000
    keys = np.sort(rng.integers(0, 1_500_000, n))
    t = pa.table({"k": pa.array(keys), "v": pa.array(rng.integers(0, 1000, n))})
    q = pw.LazyFrame(t).set_sorted("k").group_by("k", maintain_order=True).agg(pw.col("v").sum().alias("s"), pw.col("v").first().alias("f"),
                                                                              pw.col("v").last().alias("l"), pw.len().alias("n"))
    check(q)


def test_unflagged_sorted_input_takes_the_other_tiers_with_the_same_answer():
    rng = np.random.default_rng(56)
    n = 300_000
    t = pa.table({"k": pa.array(np.sort(rng.integers(0, 900, n))), "v": pa.array(rng.normal(size=n) + 10.0)})
    lf = pw.LazyFrame(t)
    a = engine.run_group_by(t, lf.group_by("k", maintain_order=True).agg(*AGGS()).plan)
    assert engine.last_timings()["strategy"] != 8
    q = lf.set_sorted("k").group_by("k", maintain_order=True).agg(*AGGS())
    b = check(q)
    G.assert_tables_equal(a, b, rtol=1e-12)


@pytest.mark.parametrize("n", [0, 1, 5, 100_003])
def test_group_slices_match_partition_to_groups(n):
    rng = np.random.default_rng(57)
    keys = np.sort(rng.integers(-50, 50, n))
    mask = np.zeros(n, dtype=bool)
    mask[: n // 10] = True
    arr = pa.array(keys, mask=mask) if n else pa.array([], type=pa.int64())
    t = pa.table({"k": arr, "x": pa.array(np.arange(n))})
    first, lens = engine.group_slices(t, ["k"])
    want_first, want_len = oracle.partition_to_groups(arr)
    assert first.type == pa.uint32() and lens.type == pa.uint32()
    assert first.to_numpy(zero_copy_only=False).tolist() == want_first.tolist()
    assert lens.to_numpy(zero_copy_only=False).tolist() == want_len.tolist()


def test_group_slices_two_string_columns():
    t = pa.table({"a": pa.array(["x", "x", "x", "y", "y", None, None, "z"]), "b": pa.array([1, 1, 2, 2, 2, 2, 3, 3])})
    first, lens = engine.group_slices(t, ["a", "b"])
    assert first.to_pylist() == [0, 2, 3, 5, 6, 7] and lens.to_pylist() == [2, 1, 2, 1, 1, 1]
