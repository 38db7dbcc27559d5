"""group_by_dynamic on the GPU vs the oracle's literal restatement of the reference's window sweep (C4 shapes)."""
import numpy as np
import pyarrow as pa
import pytest

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import engine
from tests import golden_util as G
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("by_symbol", [True, False])
@pytest.mark.parametrize("n", [1, 1000, 250_000])
def test_ohlcv_one_minute_bars(n, by_symbol):
    t = synth.ohlcv(n, n_symbols=17, seed=4, mean_gap_us=200_000)
    q = synth.ohlcv_query(t, by_symbol=by_symbol)
    got = engine.run_group_by(q.table, q.plan)
    G.assert_tables_equal(got, oracle.collect(q), rtol=1e-12)   # row order included: keys ascending, windows ascending


@pytest.mark.parametrize("closed", ["left", "right", "both", "none"])
@pytest.mark.parametrize("every,period,offset", [("1m", None, None), ("1m", "3m", None), ("2m", "1m", "30s"), ("1m", "1m", "-90s")])
def test_window_shapes_no_keys(closed, every, period, offset):
    t = synth.ohlcv(60_000, n_symbols=3, seed=8, mean_gap_us=90_000)
    c = pw.col
    q = (pw.LazyFrame(t).group_by_dynamic("ts", every=every, period=period, offset=offset, closed=closed, include_boundaries=True)
         .agg(c("price").first().alias("open"), c("price").max().alias("high"), c("price").min().alias("low"),
              c("price").last().alias("close"), c("volume").sum().alias("volume"), c("price").mean().alias("vwap_ish"),
              pw.len().alias("n")))
    got = engine.run_group_by(q.table, q.plan)
    G.assert_tables_equal(got, oracle.collect(q), rtol=1e-12)


@pytest.mark.parametrize("opts", [{}, {"flags": engine.FLAG_FORCE_SEGMENTED}, {"flags": engine.FLAG_NO_SEGMENTED}, {"flags": engine.FLAG_NO_SEGMENTED | engine.FLAG_FORCE_HOT, "hot_table_slots": 16}])
def test_tumbling_hash_and_segmented_agree(opts):
    t = synth.ohlcv(120_000, n_symbols=5, seed=9, mean_gap_us=50_000)
    q = synth.ohlcv_query(t, by_symbol=False, every="10s")
    got = engine.run_group_by(q.table, q.plan, **opts)
    G.assert_tables_equal(got, oracle.collect(q), rtol=1e-12)


def test_unsorted_index_is_an_error():
    t = pa.table({"ts": pa.array([5, 3, 9, 1], type=pa.int64()), "v": pa.array([1, 2, 3, 4])})
    q = pw.LazyFrame(t).group_by_dynamic("ts", every="2i").agg(pw.col("v").sum())
    with pytest.raises(engine.PolarwayError) as e:
        engine.run_group_by(q.table, q.plan)
    assert e.value.code == -4
    with pytest.raises(ValueError):
        oracle.collect(q)


def test_filter_fused_into_sorted_windows():
    t = synth.ohlcv(150_000, n_symbols=5, seed=10, mean_gap_us=40_000)
    c = pw.col
    q = (pw.LazyFrame(t).filter(c("volume") > 500).group_by_dynamic("ts", every="30s")
         .agg(c("price").first().alias("open"), c("price").last().alias("close"), c("volume").sum().alias("v"), pw.len().alias("n")))
    got = engine.run_group_by(q.table, q.plan)
    G.assert_tables_equal(got, oracle.collect(q), rtol=1e-12)


@pytest.mark.parametrize("spike", [10**12, -10**12])
@pytest.mark.parametrize("n", [3, 50_000])
def test_unsorted_interior_value_outside_first_last_is_a_clean_error(spike, n):
    """First and last index value are in order, an interior value lies far outside [first, last]: the sorted fast path
    sizes its dense window table from the two ends, so this used to write out of bounds before `not_sorted` was seen
    (round-1 advisor finding).  The reference raises (polars-time/src/group_by/dynamic.rs:77-80)."""
    ts = np.arange(5, 5 + n, dtype=np.int64)
    ts[n // 2] = spike
    t = pa.table({"ts": pa.array(ts), "v": pa.array(np.ones(n, dtype=np.int64))})
    q = pw.LazyFrame(t).group_by_dynamic("ts", every="2i").agg(pw.col("v").sum())
    for opts in ({}, {"flags": engine.FLAG_FORCE_SEGMENTED}):
        with pytest.raises(engine.PolarwayError) as e:
            engine.run_group_by(q.table, q.plan, **opts)
        assert e.value.code == -4
    # the library (and the GPU) are still healthy afterwards
    ok = pa.table({"ts": pa.array(np.arange(n, dtype=np.int64)), "v": pa.array(np.ones(n, dtype=np.int64))})
    q2 = pw.LazyFrame(ok).group_by_dynamic("ts", every="2i").agg(pw.col("v").sum())
    G.assert_tables_equal(engine.run_group_by(q2.table, q2.plan), oracle.collect(q2))


@pytest.mark.parametrize("typ", [pa.int8(), pa.int16(), pa.uint32(), pa.duration("us"), pa.float64()])
def test_index_dtype_must_be_date_datetime_int32_int64(typ):
    """polars-time/src/group_by/dynamic.rs:210-258: anything but Date, Datetime, Int32, Int64 is rejected."""
    base = pa.array([1, 2, 3, 4], type=pa.int64())
    t = pa.table({"ts": base.cast(typ), "v": pa.array([1, 2, 3, 4])})
    q = pw.LazyFrame(t).group_by_dynamic("ts", every="2i").agg(pw.col("v").sum())
    with pytest.raises(engine.PolarwayError) as e:
        engine.run_group_by(q.table, q.plan)
    assert e.value.code == -1 and "Date, Datetime, Int32, Int64" in str(e.value)


@pytest.mark.parametrize("typ", [pa.int32(), pa.date32(), pa.timestamp("ms")])
def test_index_dtypes_the_reference_accepts(typ):
    base = pa.array([0, 1, 2, 5, 9, 9, 12], type=pa.int32() if typ in (pa.int32(), pa.date32()) else pa.int64())
    t = pa.table({"ts": base.cast(typ), "v": pa.array([1, 2, 3, 4, 5, 6, 7])})
    every = "2i" if typ == pa.int32() else ("2d" if typ == pa.date32() else "2ms")
    q = pw.LazyFrame(t).group_by_dynamic("ts", every=every).agg(pw.col("v").sum(), pw.len().alias("n"))
    G.assert_tables_equal(engine.run_group_by(q.table, q.plan), oracle.collect(q))


def _keyed_frame(n=60_000, n_keys=6, seed=41):
    rng = np.random.default_rng(seed)
    ts = np.sort(rng.integers(0, 10_000_000, n)).astype(np.int64)
    return ts, rng.integers(0, n_keys, n).astype(np.int64), rng.integers(1, 100, n).astype(np.int64)


def test_index_unsorted_inside_a_key_raises_like_the_reference():
    """polars-time/src/group_by/dynamic.rs:77-80, 327: with `group_by=` keys the frame is sorted by the keys (stable) and
    every key slice must be ascending.  One swapped pair inside ONE key — the rows of the other keys between them keep the
    column far from obviously unsorted."""
    ts, key, v = _keyed_frame()
    rows = np.flatnonzero(key == 3)
    a, b = rows[len(rows) // 2], rows[len(rows) // 2 + 1]
    assert ts[a] < ts[b]
    ts[a], ts[b] = ts[b], ts[a]
    t = pa.table({"ts": pa.array(ts), "k": pa.array(key), "v": pa.array(v)})
    q = pw.LazyFrame(t).group_by_dynamic("ts", every="1000i", group_by="k").agg(pw.col("v").sum().alias("s"))
    with pytest.raises(ValueError):
        oracle.collect(q)
    for opts in ({}, {"flags": engine.FLAG_FORCE_HOT}, {"flags": engine.FLAG_FORCE_GLOBAL}):
        with pytest.raises(engine.PolarwayError) as e:
            engine.run_group_by(q.table, q.plan, **opts)
        assert e.value.code == -4
    # a filter that drops one row of the pair leaves every key slice ascending: no error, same answer as the oracle
    q2 = (pw.LazyFrame(t.append_column("row", pa.array(np.arange(len(ts), dtype=np.int64)))).filter(pw.col("row") != int(a))
          .group_by_dynamic("ts", every="1000i", group_by="k").agg(pw.col("v").sum().alias("s"), pw.len().alias("n")))
    G.assert_tables_equal(engine.run_group_by(q2.table, q2.plan), oracle.collect(q2))
    # ... and one that keeps both still raises
    q3 = (pw.LazyFrame(t).filter(pw.col("v") > 0).group_by_dynamic("ts", every="1000i", group_by="k").agg(pw.col("v").sum().alias("s")))
    with pytest.raises(engine.PolarwayError) as e:
        engine.run_group_by(q3.table, q3.plan)
    assert e.value.code == -4


def test_keys_sorted_first_then_time_is_accepted_on_a_resident_frame():
    # the layout `sort(["k", "ts"])` leaves: ascending inside every key, not over the whole column; the verdict is
    # remembered on the frame (second query: no second check)
    ts, key, v = _keyed_frame(seed=42)
    order = np.lexsort((ts, key))
    t = pa.table({"ts": pa.array(ts[order]), "k": pa.array(key[order]), "v": pa.array(v[order])})
    frame = engine.DeviceFrame(t)
    q = pw.LazyFrame(t).group_by_dynamic("ts", every="1000i", group_by="k").agg(pw.col("v").sum().alias("s"), pw.len().alias("n"))
    want = oracle.collect(q)
    G.assert_tables_equal(frame.group_by(q.plan), want)
    first = engine.last_timings()["kernel_launches"]
    G.assert_tables_equal(frame.group_by(q.plan), want)
    assert engine.last_timings()["kernel_launches"] < first
    frame.free()


@pytest.mark.parametrize("closed", ["left", "right", "both", "none"])
@pytest.mark.parametrize("every,period,offset", [("1m", "3m", None), ("1m", "150s", "20s"), ("2m", "2m", "-30s"), ("30s", "5m", None)])
def test_overlapping_windows_with_group_by_keys(closed, every, period, offset):
    """period > every together with `group_by=` keys.  Every row joins all the windows that contain it
    (pw_overlap.cuh); the non-empty (key, window) groups are the reference's windows, key slice by key slice
    (windows/group_by.rs:79-246 per slice, dynamic.rs:317-362)."""
    t = synth.ohlcv(40_000, n_symbols=7, seed=51, mean_gap_us=400_000)
    c = pw.col
    q = (pw.LazyFrame(t).group_by_dynamic("ts", every=every, period=period, offset=offset, closed=closed, group_by="symbol", include_boundaries=True)
         .agg(c("price").first().alias("open"), c("price").max().alias("high"), c("price").min().alias("low"),
              c("price").last().alias("close"), c("volume").sum().alias("volume"), c("price").mean().alias("mean"), pw.len().alias("n")))
    got = engine.run_group_by(q.table, q.plan)
    G.assert_tables_equal(got, oracle.collect(q), rtol=1e-12)   # row order included: key slices ascending, windows ascending


@pytest.mark.parametrize("keys", [None, "symbol"])
def test_overlapping_windows_with_a_filter(keys):
    t = synth.ohlcv(50_000, n_symbols=5, seed=52, mean_gap_us=300_000)
    c = pw.col
    q = (pw.LazyFrame(t).filter(c("volume") > 300).group_by_dynamic("ts", every="1m", period="4m", closed="right", group_by=keys)
         .agg(c("price").first().alias("open"), c("price").last().alias("close"), c("volume").sum().alias("v"), pw.len().alias("n")))
    got = engine.run_group_by(q.table, q.plan)
    G.assert_tables_equal(got, oracle.collect(q), rtol=1e-12)


def test_overlapping_windows_hash_path_equals_slice_path_without_keys():
    # keys empty, no filter: the closed-form slice path (pw_dynamic.cu) is the default; the hash path must agree
    t = synth.ohlcv(30_000, n_symbols=3, seed=53, mean_gap_us=500_000)
    c = pw.col
    q = (pw.LazyFrame(t).group_by_dynamic("ts", every="1m", period="3m", offset="-45s", closed="left", include_boundaries=True)
         .agg(c("price").sum().alias("s"), c("price").first().alias("open"), pw.len().alias("n")))
    want = oracle.collect(q)
    G.assert_tables_equal(engine.run_group_by(q.table, q.plan), want, rtol=1e-12)
    G.assert_tables_equal(engine.run_group_by(q.table, q.plan, flags=engine.FLAG_NO_SEGMENTED), want, rtol=1e-12)


def test_overlapping_windows_two_keys_one_nullable_and_integer_index():
    rng = np.random.default_rng(54)
    n = 20_000
    ts = np.sort(rng.integers(-5_000, 200_000, n)).astype(np.int64)      # negative index values: windows of negative number
    t = pa.table({"ts": pa.array(ts), "a": pa.array(rng.integers(0, 4, n).astype("int8"), mask=rng.random(n) < 0.05),
                  "b": pa.array(rng.integers(0, 3, n).astype(str)), "v": pa.array(rng.integers(-9, 9, n))})
    q = (pw.LazyFrame(t).group_by_dynamic("ts", every="1000i", period="2500i", group_by=["a", "b"])
         .agg(pw.col("v").sum().alias("s"), pw.col("v").min().alias("lo"), pw.len().alias("n")))
    G.assert_tables_equal(engine.run_group_by(q.table, q.plan), oracle.collect(q))
