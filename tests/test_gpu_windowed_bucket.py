"""group_by_dynamic BY ONE DENSE INTEGER KEY (C4: OHLCV one-minute bars per symbol) through the windowed bucket tier
(pw_bucket.cuh, strategy 6): rows bucketed by symbol per tile, accumulators of (symbol, current window) in registers,
published when the window changes.  CUDA vs the oracle's restatement of the reference's window sweep
(polars-time/src/windows/group_by.rs:79-246); first/last/min/max/count/integer sums bit-exact, f64 sums 1e-12."""
import numpy as np
import pyarrow as pa
import pytest

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import engine
from tests import golden_util as G
from tests import synth

pytestmark = pytest.mark.gpu


def run_both(q, rtol=1e-12, expect=6, **opts):
    want = oracle.collect(q)
    got = engine.run_group_by(q.table, q.plan, **opts)
    assert engine.last_timings()["strategy"] == expect, engine.last_timings()
    G.assert_tables_equal(got, want, rtol=rtol)
    # the composite-key hash path answers the same
    o2 = dict(opts)
    o2["flags"] = o2.get("flags", 0) | engine.FLAG_NO_BUCKETS
    got2 = engine.run_group_by(q.table, q.plan, **o2)
    assert engine.last_timings()["strategy"] != 6
    G.assert_tables_equal(got2, want, rtol=rtol)
    return got


@pytest.mark.parametrize("n_symbols", [24, 100, 700])
def test_long_windows_chosen_by_the_pilot(n_symbols):
    # 600 k ticks, ~1 ms apart: a one-minute window holds ~60 k rows = dozens of tiles; window changes inside tiles
    t = synth.ohlcv(600_000, n_symbols=n_symbols, seed=31, mean_gap_us=1000)
    run_both(synth.ohlcv_query(t, by_symbol=True))


def test_a_handful_of_symbols_keeps_the_hash_path():
    # below 16 ids a bucket per id is the wrong shape (hundreds of rows per id per tile on one counter)
    t = synth.ohlcv(600_000, n_symbols=3, seed=38, mean_gap_us=1000)
    q = synth.ohlcv_query(t, by_symbol=True)
    got = engine.run_group_by(q.table, q.plan)
    assert engine.last_timings()["strategy"] != 6
    G.assert_tables_equal(got, oracle.collect(q), rtol=1e-12)


def test_signed_symbols_with_offset_and_filter():
    t = synth.ohlcv(500_000, n_symbols=50, seed=32, mean_gap_us=1500)
    sym = t.column("symbol").to_numpy().astype(np.int64) - 25     # ids -25 .. 24: the table sentinels -1 / -2 are inside
    t = t.set_column(1, "symbol", pa.array(sym))
    c = pw.col
    q = (pw.LazyFrame(t).filter(c("volume") > 300).group_by_dynamic("ts", every="1m", offset="-20s", closed="right", group_by="symbol")
         .agg(c("price").first().alias("open"), c("price").max().alias("high"), c("price").min().alias("low"), c("price").last().alias("close"),
              c("volume").sum().alias("volume"), c("price").mean().alias("mean"), pw.len().alias("n")))
    run_both(q)


@pytest.mark.parametrize("mean_gap_us", [200_000, 5_000_000])
def test_short_windows_under_force_flag(mean_gap_us):
    # windows of a few hundred rows (or a dozen): most tiles span more than two windows, the surplus rows take the HBM
    # path; small inputs reach the tier through the force flag only
    t = synth.ohlcv(40_000, n_symbols=17, seed=33, mean_gap_us=mean_gap_us)
    run_both(synth.ohlcv_query(t, by_symbol=True), flags=engine.FLAG_FORCE_HOT)


def test_index_not_sorted_across_symbols_still_aggregates():
    # sorted inside every symbol but not globally (two concatenated sessions): every row finds its (symbol, window) group
    a = synth.ohlcv(200_000, n_symbols=8, seed=34, mean_gap_us=2000)
    b = synth.ohlcv(200_000, n_symbols=8, seed=35, mean_gap_us=2000)
    b = b.set_column(1, "symbol", pa.array(b.column("symbol").to_numpy() + 8))
    t = pa.concat_tables([a, b])
    run_both(synth.ohlcv_query(t, by_symbol=True))


def test_nullable_values_and_null_aware_aggs():
    rng = np.random.default_rng(36)
    t = synth.ohlcv(400_000, n_symbols=20, seed=36, mean_gap_us=1000)
    n = t.num_rows
    t = t.set_column(2, "price", pa.array(t.column("price").to_numpy(), mask=rng.random(n) < 0.1))
    c = pw.col
    q = (pw.LazyFrame(t).group_by_dynamic("ts", every="1m", group_by="symbol")
         .agg(c("price").first().alias("open"), c("price").max().alias("high"), c("price").min().alias("low"), c("price").last().alias("close"),
              c("price").count().alias("n_valid"), c("price").sum().alias("sum"), c("volume").sum().alias("volume"),
              c("price").first(ignore_nulls=True).alias("open_nn"), c("price").null_count().alias("nulls")))
    run_both(q)


def test_nullable_key_keeps_the_hash_path():
    rng = np.random.default_rng(37)
    t = synth.ohlcv(400_000, n_symbols=20, seed=37, mean_gap_us=1000)
    t = t.set_column(1, "symbol", pa.array(t.column("symbol").to_numpy(), mask=rng.random(t.num_rows) < 0.01))
    q = synth.ohlcv_query(t, by_symbol=True)
    got = engine.run_group_by(q.table, q.plan)
    assert engine.last_timings()["strategy"] != 6
    G.assert_tables_equal(got, oracle.collect(q), rtol=1e-12)
