"""High-cardinality tier.  Second form (strategy 10, pw_radix.cuh): rows radix-partitioned by key hash into 32-byte records
(one or two levels, staged in shared memory), one CTA per partition aggregates in a shared-memory table and appends
finished groups to the result table.  First form (strategy 5, pw_partition.cuh): one scatter pass, then the ordinary scan
with its hot table.  FLAG_FORCE_PARTITION (the analogue of POLARS_FORCE_PARTITION) drives unit-sized inputs through them;
the last tests reach the tier through the pilot's own decision."""
import numpy as np
import pyarrow as pa
import pytest

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import engine
from tests import golden_util as G
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["radix", "radix_two_levels", "radix_tiny_tables", "scatter_scan"])
def tier(request, monkeypatch):
    """Which form of the tier the forced runs take (environment switches read per query by pw_engine.cu)."""
    if request.param == "radix_two_levels":
        monkeypatch.setenv("PW_RADIX_LOG2_PARTS", "11")      # 2^6 x 2^5 partitions even for unit-sized inputs
    elif request.param == "radix_tiny_tables":
        monkeypatch.setenv("PW_RADIX_LOG2_PARTS", "1")       # two partitions: shared-memory tables overflow into the HBM region
    elif request.param == "scatter_scan":
        monkeypatch.setenv("PW_NO_RADIX_NOW", "1")
    return 5 if request.param == "scatter_scan" else 10


def run(q, sort_by, expect=5, rtol=1e-12, **opts):
    got = engine.run_group_by(q.table, q.plan, **opts)
    st = engine.last_timings()["strategy"]
    assert st == expect or (expect == 10 and engine.last_timings()["retries"] > 0), st
    G.assert_tables_equal(got, oracle.collect(q), sort_by=sort_by, rtol=rtol)
    return got


ALL = lambda c: [pw.col(c).sum().alias("sum"), pw.col(c).mean().alias("mean"), pw.col(c).min().alias("min"),
                 pw.col(c).max().alias("max"), pw.col(c).count().alias("count"), pw.col(c).first().alias("first"),
                 pw.col(c).last().alias("last"), pw.len().alias("len")]


@pytest.mark.parametrize("n,groups", [(1, 1), (129, 7), (400_000, 40_000), (300_000, 290_000)])
def test_forced_partition_null_aware_aggs(n, groups, tier):
    t = synth.c3_table(n, groups, seed=31)
    q = pw.LazyFrame(t).group_by("key").agg(*ALL("value"))
    run(q, ["key"], expect=tier, flags=engine.FLAG_FORCE_PARTITION)


@pytest.mark.parametrize("vdtype", ["int8", "uint16", "int32", "uint32", "int64", "float32", "float64"])
def test_forced_partition_value_dtypes_with_filter_and_order(vdtype, tier):
    rng = np.random.default_rng(32)
    n = 150_000
    t = pa.table({"k": pa.array(rng.integers(-3, 20_000, n), mask=rng.random(n) < 0.01),     # -1 / -2 alias the key sentinels
                  "p": pa.array(rng.integers(0, 100, n).astype("int32")),
                  "v": pa.array(rng.integers(0, 100, n).astype(vdtype), mask=rng.random(n) < 0.1)})
    q = pw.LazyFrame(t).filter(pw.col("p") < 80).group_by("k", maintain_order=True).agg(*ALL("v"))
    run(q, None, expect=tier, rtol=1e-5 if vdtype == "float32" else 1e-12, flags=engine.FLAG_FORCE_PARTITION)


def test_forced_partition_two_keys_one_nullable(tier):
    rng = np.random.default_rng(33)
    n = 200_000
    t = pa.table({"a": pa.array(rng.integers(0, 300, n).astype("int16"), mask=rng.random(n) < 0.02),
                  "b": pa.array(rng.random(n).round(2)),                                      # float key
                  "v": pa.array(rng.integers(-50, 50, n))})
    q = pw.LazyFrame(t).group_by("a", "b").agg(pw.col("v").sum().alias("s"), pw.len().alias("n"), pw.col("v").min().alias("lo"))
    run(q, ["a", "b"], expect=tier, flags=engine.FLAG_FORCE_PARTITION)


def test_string_keys_are_not_partitioned():
    t = synth.lineitem(50_000, seed=34)
    q = synth.q1_query(t)
    run(q, ["l_returnflag", "l_linestatus"], expect=1, flags=engine.FLAG_FORCE_PARTITION)


def test_pilot_chooses_partitioning_for_many_groups(tier, request):
    if "tiny_tables" in request.node.name:
        pytest.skip("two partitions for 1e6 groups only exercise the retry")
    # 6e6 rows, 1e6 keys in random order: consecutive rows share no groups and the table is far beyond the hot tier
    n, g = 6_000_000, 1_000_000
    t = synth.c3_table(n, g, seed=35)
    q = pw.LazyFrame(t).group_by("key").agg(pw.col("value").sum().alias("sum"), pw.col("value").count().alias("count"),
                                            pw.col("value").min().alias("min"), pw.col("value").max().alias("max"),
                                            pw.col("value").first().alias("first"), pw.col("value").last().alias("last"))
    got = run(q, ["key"], expect=tier)
    # the same answer without partitioning (idempotence across strategies)
    got2 = engine.run_group_by(q.table, q.plan, flags=engine.FLAG_NO_PARTITION)
    assert engine.last_timings()["strategy"] == 2
    G.assert_tables_equal(got.sort_by("key"), got2.sort_by("key"), rtol=1e-12)


def test_skewed_keys_overflow_the_partition_tables():
    # the pilot sees ~1e6 groups, but half of the rows carry keys from a second, denser population the strided sample
    # under-counts: partitions hold more groups than planned; whatever does not fit a CTA's table lands in the overflow
    # region (or the run is repeated on the plain HBM table) — the answer must not change
    rng = np.random.default_rng(36)
    n = 6_000_000
    keys = rng.integers(0, 1_000_000, n, dtype=np.int64)
    keys[::2] = rng.integers(1 << 40, (1 << 40) + 3_000_000, n // 2, dtype=np.int64)
    t = pa.table({"key": keys, "value": rng.random(n)})
    q = pw.LazyFrame(t).group_by("key").agg(pw.col("value").sum().alias("sum"), pw.len().alias("len"), pw.col("value").max().alias("max"))
    got = engine.run_group_by(q.table, q.plan)
    assert engine.last_timings()["strategy"] in (10, 2)
    G.assert_tables_equal(got, oracle.collect(q), sort_by=["key"], rtol=1e-12)
