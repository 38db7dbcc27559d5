"""CUDA path vs the CPU oracle on seeded synthetic inputs (sizes the oracle finishes in seconds), plus
size-independent properties at larger sizes.  Bit-exact for integer/count/min/max/first/last and group
membership; f64 sum/mean within 1e-12 relative (BASELINE.json north_star); f32 within 1e-5."""
import numpy as np
import pyarrow as pa
import pytest

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import engine
from tests import golden_util as G
from tests import synth

pytestmark = pytest.mark.gpu

RTOL_F64 = 1e-12


def check(lazy, sort_by, rtol=RTOL_F64, **opts):
    got = engine.run_group_by(lazy.table, lazy.plan, **opts)
    want = oracle.collect(lazy)
    G.assert_tables_equal(got, want, sort_by=sort_by, rtol=rtol)
    return got


@pytest.mark.parametrize("strategy", [{}, {"flags": engine.FLAG_FORCE_HOT}, {"flags": engine.FLAG_FORCE_GLOBAL}])
@pytest.mark.parametrize("n,groups", [(1, 1), (127, 5), (128, 3), (1000, 1000), (100_003, 1000), (1_000_000, 1000), (300_000, 100_000)])
def test_low_and_high_cardinality_i64_key_f64_value(n, groups, strategy):
    t = synth.c2_table(n, groups, seed=2)
    q = pw.LazyFrame(t).group_by("key").agg(pw.col("value").sum().alias("sum"), pw.col("value").mean().alias("mean"),
                                            pw.col("value").min().alias("min"), pw.col("value").max().alias("max"))
    check(q, ["key"], **strategy)


@pytest.mark.parametrize("strategy", [{}, {"flags": engine.FLAG_FORCE_HOT, "hot_table_slots": 64}, {"flags": engine.FLAG_FORCE_GLOBAL}])
def test_c3_null_aware_aggs(strategy):
    t = synth.c3_table(400_000, 40_000, seed=3)
    q = pw.LazyFrame(t).group_by("key").agg(
        pw.col("value").sum().alias("sum"), pw.col("value").mean().alias("mean"), pw.col("value").min().alias("min"),
        pw.col("value").max().alias("max"), pw.col("value").count().alias("count"), pw.col("value").first().alias("first"),
        pw.col("value").last().alias("last"), pw.len().alias("len"))
    check(q, ["key"], **strategy)


@pytest.mark.parametrize("maintain_order", [False, True])
def test_q1_shape(maintain_order):
    t = synth.lineitem(200_000, seed=1)
    q = synth.q1_query(t, maintain_order=maintain_order)
    got = check(q, None if maintain_order else ["l_returnflag", "l_linestatus"])
    assert got.num_rows == 4


def test_q1_reference_fixture_schema():
    # the reference's own lineitem head (examples/datasets/pds_heads/lineitem.feather) fixes the dtypes;
    # tests/golden/lineitem_head.json is its 10 rows (see tests/golden/make_lineitem_head.py)
    t = synth.lineitem_head()
    q = synth.q1_query(t, maintain_order=True)
    check(q, None)


@pytest.mark.parametrize("dtype", ["int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64", "float32", "float64"])
def test_every_numeric_dtype(dtype):
    rng = np.random.default_rng(7)
    n = 50_000
    vals = rng.integers(0, 100, n).astype(dtype)
    mask = rng.random(n) < 0.1
    t = pa.table({"k": pa.array(rng.integers(0, 37, n), type=pa.int32()),
                  "v": pa.array(vals, mask=mask)})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("sum"), pw.col("v").mean().alias("mean"),
                                          pw.col("v").min().alias("min"), pw.col("v").max().alias("max"),
                                          pw.col("v").first().alias("first"), pw.col("v").last().alias("last"),
                                          pw.col("v").count().alias("count"))
    check(q, ["k"], rtol=1e-5 if dtype == "float32" else RTOL_F64)


def test_filter_fused_with_nulls_and_sliced_input():
    # sliced arrays -> non-zero ArrowArray.offset and a validity bit offset
    rng = np.random.default_rng(11)
    n = 70_001
    t = pa.table({"k": pa.array(rng.integers(0, 50, n), mask=rng.random(n) < 0.02),
                  "p": pa.array(rng.integers(-1000, 1000, n), mask=rng.random(n) < 0.05),
                  "v": pa.array(rng.normal(size=n), mask=rng.random(n) < 0.05)}).slice(13, n - 29)
    q = (pw.LazyFrame(t).filter((pw.col("p") >= -500) & (pw.col("p") < 700)).group_by("k", maintain_order=True)
         .agg(pw.col("v").sum().alias("s"), pw.col("v").count().alias("c"), pw.len().alias("n"), pw.col("p").max().alias("pm")))
    check(q, None)


def test_string_keys_and_multi_keys():
    rng = np.random.default_rng(5)
    n = 120_000
    words = np.array(["", "a", "bb", "ccc", "dddd", "eeeee", "twelve_bytes", "x y", "Ünï"])
    t = pa.table({"s": pa.array(words[rng.integers(0, len(words), n)], mask=rng.random(n) < 0.03),
                  "i": pa.array(rng.integers(-3, 3, n), type=pa.int16()),
                  "f": pa.array(rng.choice([0.0, -0.0, 1.5, np.nan], n)),
                  "v": pa.array(rng.integers(0, 1000, n))})
    q = pw.LazyFrame(t).group_by("s", "i", "f").agg(pw.col("v").sum().alias("v"), pw.len().alias("n"))
    got = engine.run_group_by(q.table, q.plan)
    want = oracle.collect(q)
    # -0.0 and 0.0 are one group (TotalEq); compare after normalising the representative's sign
    norm = lambda tb: tb.set_column(2, "f", pa.array(np.where(tb["f"].to_numpy(zero_copy_only=False) == 0, 0.0, tb["f"].to_numpy(zero_copy_only=False))))
    G.assert_tables_equal(norm(got), norm(want), sort_by=["s", "i", "f"])


def test_sentinel_valued_keys_and_global_agg():
    t = pa.table({"k": pa.array([-1, -2, -1, None, 5, -2, None], type=pa.int64()), "v": pa.array([1, 2, 3, 4, 5, 6, 7])})
    q = pw.LazyFrame(t).group_by("k", maintain_order=True).agg(pw.col("v").sum().alias("v"))
    for opts in ({}, {"flags": engine.FLAG_FORCE_GLOBAL}):
        check(q, None, **opts)
    # u64 max as a key value
    t = pa.table({"k": pa.array([2**64 - 1, 2**64 - 2, 2**64 - 1, 0], type=pa.uint64()), "v": pa.array([1, 2, 3, 4])})
    check(pw.LazyFrame(t).group_by("k", maintain_order=True).agg(pw.col("v").sum().alias("v")), None)
    # no keys at all: one global group
    check(pw.LazyFrame(t).group_by().agg(pw.col("v").sum().alias("v"), pw.len().alias("n")), None)


def test_empty_input():
    t = pa.table({"k": pa.array([], type=pa.int64()), "v": pa.array([], type=pa.float64())})
    q = pw.LazyFrame(t).group_by("k").agg(pw.col("v").sum().alias("v"))
    check(q, None)


def test_property_sum_of_partial_sums_large():
    # 2e7 rows: integer sums/counts are exact, so totals must match closed forms (no oracle needed)
    n, g = 20_000_000, 1000
    t = synth.c2_table(n, g, seed=9, int_value=True)
    q = pw.LazyFrame(t).group_by("key").agg(pw.col("value").sum().alias("s"), pw.len().alias("n"),
                                            pw.col("value").min().alias("lo"), pw.col("value").max().alias("hi"))
    got = engine.run_group_by(q.table, q.plan)
    v = t["value"].to_numpy()
    assert got.num_rows == g
    assert sum(got["n"].to_pylist()) == n
    assert sum(got["s"].to_pylist()) == int(v.sum())
    assert min(got["lo"].to_pylist()) == int(v.min()) and max(got["hi"].to_pylist()) == int(v.max())
    # idempotence: same answer under the other strategy
    got2 = engine.run_group_by(q.table, q.plan, flags=engine.FLAG_FORCE_GLOBAL)
    G.assert_tables_equal(got.sort_by("key"), got2.sort_by("key"))


def test_long_string_key_groups_by_its_bytes():
    t = pa.table({"s": pa.array(["short", "this string is longer than twelve bytes", "short", "this string is longer than twelve bytes"]), "v": pa.array([1, 2, 3, 4])})
    q = pw.LazyFrame(t).group_by("s").agg(pw.col("v").sum())
    got = engine.run_group_by(q.table, q.plan)
    G.assert_tables_equal(got, oracle.collect(q), sort_by=["s"])


def test_plugin_shim_end_to_end_multi_chunk():
    # the very call sequence of the reference's plugin loader (plugin.rs:75-142), chunked inputs included
    from polaroid_b200 import plugin_loader
    t = synth.lineitem(30_000, seed=6)
    t = pa.concat_tables([t.slice(0, 11_111), t.slice(11_111, 7), t.slice(11_118)])
    q = synth.q1_query(t, maintain_order=True)
    got = plugin_loader.call_plugin(t, plugin_loader.plan_to_kwargs(t.schema, q.plan))
    want = oracle.collect(q)
    got = got.set_column(0, "l_returnflag", got["l_returnflag"].cast(pa.string())).set_column(1, "l_linestatus", got["l_linestatus"].cast(pa.string()))
    G.assert_tables_equal(got, want, rtol=1e-12)
