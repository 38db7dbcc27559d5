"""Boundary hardening: zero-copy device frames, input ownership through the plugin ABI, concurrent
callers.  Everything goes through the C ABI of libpolarway_b200.so."""
import ctypes as C
import threading

import numpy as np
import pyarrow as pa
import pytest

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import engine
from tests import golden_util as G
from tests import synth

pytestmark = pytest.mark.gpu


def _torch():
    return pytest.importorskip("torch")


def test_frame_from_device_wraps_torch_tensors_zero_copy():
    """pw_b200_frame_from_device: the Arrow buffers are DEVICE pointers (here torch tensors); no upload happens and the
    result equals the uploaded frame's and the oracle's."""
    torch = _torch()
    t = synth.c2_table(300_000, 500, seed=11)
    keys = torch.from_numpy(np.asarray(t.column("key").to_numpy())).cuda()
    vals = torch.from_numpy(np.asarray(t.column("value").to_numpy())).cuda()
    frame = engine.DeviceFrame.from_device([
        ("key", pa.int64(), keys.shape[0], keys.data_ptr(), 0, 0, keys),
        ("value", pa.float64(), vals.shape[0], vals.data_ptr(), 0, 0, vals)])
    q = pw.LazyFrame(t).group_by("key").agg(pw.col("value").sum().alias("s"), pw.col("value").min().alias("lo"),
                                            pw.col("value").max().alias("hi"), pw.col("value").mean().alias("m"), pw.len().alias("n"))
    torch.cuda.synchronize()
    got = frame.group_by(q.plan)
    G.assert_tables_equal(got, oracle.collect(q), sort_by=["key"], rtol=1e-12)
    # zero copy: changing the tensor changes the next result (the frame reads the caller's memory)
    vals.mul_(2.0)
    torch.cuda.synchronize()
    t2 = pa.table({"key": t.column("key"), "value": pa.array(np.asarray(t.column("value").to_numpy()) * 2.0)})
    q2 = pw.LazyFrame(t2).group_by("key").agg(*[pw.col("value").sum().alias("s"), pw.col("value").min().alias("lo"),
                                               pw.col("value").max().alias("hi"), pw.col("value").mean().alias("m"), pw.len().alias("n")])
    G.assert_tables_equal(frame.group_by(q.plan), oracle.collect(q2), sort_by=["key"], rtol=1e-12)
    frame.free()


def test_frame_from_device_validity_bitmap_and_views():
    """nullable values (LSB-first device bitmap) and inline string-view keys built on the device."""
    torch = _torch()
    n = 100_003
    rng = np.random.default_rng(5)
    codes = rng.integers(0, 3, n)
    valid = rng.random(n) > 0.1
    vals = rng.random(n) * 10
    views = np.zeros((n, 4), dtype=np.int32)
    views[:, 0] = 1
    views[:, 1] = np.array([65, 78, 82])[codes]
    bitmap = np.packbits(valid, bitorder="little")
    d_views, d_vals, d_bits = torch.from_numpy(views).cuda(), torch.from_numpy(vals).cuda(), torch.from_numpy(bitmap).cuda()
    frame = engine.DeviceFrame.from_device([
        ("flag", pa.string_view(), n, d_views.data_ptr(), 0, 0, d_views),
        ("v", pa.float64(), n, d_vals.data_ptr(), d_bits.data_ptr(), int(n - valid.sum()), (d_vals, d_bits))])
    host = pa.table({"flag": pa.array(np.array(["A", "N", "R"])[codes]), "v": pa.array(vals, mask=~valid)})
    q = pw.LazyFrame(host).group_by("flag").agg(pw.col("v").sum().alias("s"), pw.col("v").count().alias("c"),
                                                pw.col("v").first().alias("f"), pw.col("v").last().alias("l"))
    torch.cuda.synchronize()
    got = frame.group_by(q.plan)
    got = got.set_column(0, "flag", got.column("flag").cast(pa.string()))
    G.assert_tables_equal(got, oracle.collect(q), sort_by=["flag"], rtol=1e-12)
    frame.free()


def test_concurrent_callers_with_different_shapes():
    """The header promises re-entrancy (every thread has its own stream context, the JIT cache and the frame's pilot
    cache are locked): 8 host threads run different query shapes against shared and private frames at once."""
    tables = [synth.c2_table(150_000 + 1000 * i, 50 + 37 * i, seed=20 + i, int_value=(i % 2 == 1)) for i in range(4)]
    shared = engine.DeviceFrame(tables[0])
    plans = []
    for i in range(8):
        t = tables[i % 4]
        c = pw.col("value")
        aggs = [[c.sum().alias("s")], [c.min().alias("lo"), c.max().alias("hi")], [c.mean().alias("m"), pw.len().alias("n")],
                [c.first().alias("f"), c.last().alias("l"), c.count().alias("c")]][i % 4]
        plans.append(pw.LazyFrame(t).group_by("key", maintain_order=(i % 3 == 0)).agg(*aggs))
    want = [oracle.collect(q) for q in plans]
    errors, results = [], [None] * 8

    def work(i):
        try:
            for rep in range(3):
                if i % 4 == 0:
                    results[i] = shared.group_by(plans[i].plan)      # several threads share one resident frame
                else:
                    results[i] = engine.run_group_by(plans[i].table, plans[i].plan)
        except Exception as ex:  # noqa: BLE001
            errors.append((i, repr(ex)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for i in range(8):
        G.assert_tables_equal(results[i], want[i], sort_by=None if plans[i].plan.maintain_order else ["key"], rtol=1e-12)
    shared.free()


def test_result_count_beyond_the_sample_estimate_retries_with_a_larger_table():
    """Deferred group count (one host sync per query): the table is sized from a key sample; when the real group count
    overflows it, the overflow is only seen after the result copy and the query runs again with a larger table."""
    n = 400_000
    rng = np.random.default_rng(3)
    # the strided sample sees few distinct keys (every 6th row repeats 8 keys), the rest of the rows are all distinct
    keys = np.arange(n, dtype=np.int64) + 1_000_000
    keys[::2] = rng.integers(0, 8, (n + 1) // 2)
    t = pa.table({"key": keys, "value": rng.random(n)})
    q = pw.LazyFrame(t).group_by("key").agg(pw.col("value").sum().alias("s"), pw.len().alias("n"))
    got = engine.run_group_by(q.table, q.plan, initial_table_slots=256)
    G.assert_tables_equal(got, oracle.collect(q), sort_by=["key"], rtol=1e-12)
    assert engine.last_timings()["retries"] >= 1


def test_plugin_callee_releases_every_input_exactly_once():
    # plugin.rs:127-130: the caller forgets the exported inputs; "the inputs get dropped when the ffi side calls the
    # drop callback" — once per input, on success and on failure alike
    from polaroid_b200 import plugin_loader
    rng = np.random.default_rng(81)
    n = 30_000
    t = pa.table({"k": pa.array(rng.integers(0, 50, n)), "v": pa.array(rng.random(n)), "w": pa.array(rng.integers(0, 9, n))})
    q = pw.LazyFrame(t).filter(pw.col("w") < 7).group_by("k").agg(pw.col("v").sum().alias("s"), pw.len().alias("n"))
    log = []
    got = plugin_loader.call_plugin(t, plugin_loader.plan_to_kwargs(t.schema, q.plan), release_log=log)
    assert sorted(log) == [0, 1, 2]
    G.assert_tables_equal(got, oracle.collect(q), sort_by=["k"], rtol=1e-12)
    log = []
    with pytest.raises(engine.PolarwayError):
        plugin_loader.call_plugin(t, {"keys": [0], "aggs": [("x", 0, 7, None)]}, release_log=log)   # column 7 does not exist
    assert sorted(log) == [0, 1, 2]
