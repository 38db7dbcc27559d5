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
