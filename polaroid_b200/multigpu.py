"""Rows sharded over the GPUs of one box: local partial aggregates -> key-hash all-to-all -> merge.

One process per GPU (torchrun); ``torch.distributed`` (NCCL over NVLink/NVSwitch) is only the plumbing for the
single exchange step this path has:

    phase 1  every rank:  pw_b200_frame_groupby_partial   local fused scan -> packed partial rows, counting-sorted
                                                          by owner = hash(key) -> rank (HashPartitioner semantics,
                                                          polars-utils/src/hashing.rs:100-109)
    exchange              one all_to_all of row counts + one all_to_all_single of the packed rows
    phase 2  every rank:  pw_b200_merge_partials          combine rows of equal keys (GroupedReduction::combine,
                                                          polars-expr/src/reduce/mod.rs:94-105) and finalise

Every rank ends up with the groups it owns; the full result is the concatenation over ranks, exactly like the
reference's per-partition DataFrames (polars-stream/src/nodes/group_by.rs:554-563).

``exchange_rows`` is backend-agnostic (works on CPU tensors with gloo) so that the routing logic is covered by the
world_size-2 CPU tests.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List

import torch
import torch.distributed as dist


def exchange_counts(send_counts: List[int], device) -> List[int]:
    """all-to-all of one int64 per peer: how many rows every peer will send me."""
    world = dist.get_world_size()
    send = torch.tensor(send_counts, dtype=torch.int64, device=device)
    recv = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(recv, send)
    return [int(x) for x in recv.tolist()]


def exchange_rows(rows: torch.Tensor, send_counts: List[int], row_words: int) -> torch.Tensor:
    """rows: flat int64 tensor of packed rows already grouped by destination rank (send_counts[r] rows for rank r).
    Returns the rows this rank receives (flat int64), in source-rank order."""
    recv_counts = exchange_counts(send_counts, rows.device)
    in_split = [c * row_words for c in send_counts]
    out_split = [c * row_words for c in recv_counts]
    recv = torch.empty(sum(out_split), dtype=rows.dtype, device=rows.device)
    dist.all_to_all_single(recv, rows, output_split_sizes=out_split, input_split_sizes=in_split)
    return recv


GATHER_CAP_ROWS = int(os.environ.get("PW_MGPU_GATHER_ROWS", "8192"))  # 0 disables the single-collective exchange
GATHER_SMALL_ROWS = 1024   # second capacity: most low-cardinality results fit it (10x less to gather, merge and copy)
_gather_bufs = {}
_stats = {"nccl_ms": 0.0, "bytes": 0, "calls": 0}
_events = []


def reset_stats():
    _stats.update(nccl_ms=0.0, bytes=0, calls=0)
    _events.clear()


def stats() -> dict:
    """NCCL time (CUDA events around the collectives of the general exchange, averaged per call) and bytes sent per rank."""
    if _events:
        torch.cuda.synchronize()
        _stats["nccl_ms"] += sum(a.elapsed_time(b) for a, b in _events)
        _events.clear()
    n = max(1, _stats["calls"])
    return {"nccl_ms": _stats["nccl_ms"] / n, "bytes": _stats["bytes"] // n, "calls": _stats["calls"]}


def _gathered_exchange(L, engine, bq, frame, plan, rank: int, world: int, memo=None):
    """Small results: ONE all-gather of fixed-size buffers, row counts stay on the device, no host synchronisation
    between the local aggregation and the merge (the only one is the result copy at the end of the merge).  The segment
    capacity is chosen from what the first exchange of this (frame, plan) saw — GATHER_SMALL_ROWS when every rank's
    groups fit it twice over, else GATHER_CAP_ROWS; every rank reads the same gathered headers, so every rank decides
    alike.  Returns the result table, or None when some rank had too many groups (the overflow header is seen by all,
    so all fall back to the general exchange together)."""
    import pyarrow as pa
    memo = memo if memo is not None else {}
    rw = memo.get("row_words")
    if rw is None:
        rw = L.pw_b200_partial_row_words(C.byref(bq.q), frame.handle)
        if rw <= 0:
            engine._check(int(rw))
        memo["row_words"] = rw
    cap_rows = memo.get("cap_rows", GATHER_CAP_ROWS)
    words = 1 + cap_rows * (rw + 1)
    dev = torch.device("cuda", torch.cuda.current_device())
    key = (dev.index, world, words)
    if key not in _gather_bufs:
        _gather_bufs[key] = (torch.zeros(words, dtype=torch.int64, device=dev), torch.zeros(world * words, dtype=torch.int64, device=dev))
    send, gathered = _gather_bufs[key]
    # the collective is ordered against torch's current stream: the library's launches must be on that stream too
    L.pw_b200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
    rc = L.pw_b200_frame_groupby_partial_into(C.byref(bq.q), frame.handle, world, C.c_void_p(send.data_ptr()), cap_rows)
    if rc not in (0, 1):
        engine._check(rc)
    dist.all_gather_into_tensor(gathered, send)     # same stream as the library's launches: ordered, no sync
    cap = len(plan.keys) + len(plan.aggs) + 4
    out_arrays = (engine.ArrowArray * cap)()
    out_schemas = (engine.ArrowSchema * cap)()
    n_out = C.c_size_t(cap)
    rc = L.pw_b200_merge_gathered(C.byref(bq.q), frame.handle, C.c_void_p(gathered.data_ptr()), world, cap_rows, rank,
                                  out_arrays, out_schemas, C.byref(n_out))
    if rc == 1:
        # every rank sees the same overflow header, so every rank takes the same decision here
        if cap_rows >= GATHER_CAP_ROWS:
            memo["general"] = True           # too many groups for this exchange: next time straight to the all-to-all
        memo["cap_rows"] = GATHER_CAP_ROWS   # the small capacity was too optimistic (or the data changed)
        return None
    engine._check(rc)
    if "cap_rows" not in memo:
        # first exchange: the headers (one word per rank) tell how many groups every rank had
        heads = gathered.view(world, words)[:, 0].tolist()
        memo["cap_rows"] = GATHER_SMALL_ROWS if 2 * max(heads) <= GATHER_SMALL_ROWS < GATHER_CAP_ROWS else GATHER_CAP_ROWS
    names, cols = engine._import_columns(out_arrays, out_schemas, n_out.value)
    cols = engine._restore_string_types(names, cols, frame.table_schema, plan.keys)
    return pa.Table.from_arrays(cols, names=names)


def group_by_sharded(frame, plan, rank: int, world: int, row_offset: int = 0, force_all_to_all: bool = False, **opts):
    """frame: engine.DeviceFrame holding this rank's shard.  Returns the pyarrow Table of the groups this rank owns.
    force_all_to_all: skip the single-collective exchange for small results (benchmarks of the general path)."""
    import pyarrow as pa
    from . import engine
    L = engine.lib()
    # the lowered PwQuery of a plan object is reused across calls on the same frame (as DeviceFrame.group_by does)
    ck = (id(plan), row_offset, tuple(sorted(opts.items())))
    cache = frame.__dict__.setdefault("_queries_sharded", {})
    hit = cache.get(ck)
    if hit is None or hit[0] is not plan:
        if len(cache) > 64:
            cache.clear()
        hit = (plan, engine._BuiltQuery(frame.table_schema, plan, row_offset=row_offset, **opts), {})
        cache[ck] = hit
    bq, memo = hit[1], hit[2]
    nccl = dist.is_initialized() and dist.get_backend() == "nccl"
    if GATHER_CAP_ROWS > 0 and nccl and not force_all_to_all and not memo.get("general"):
        got = _gathered_exchange(L, engine, bq, frame, plan, rank, world, memo)
        if got is not None:
            return got
    if nccl:
        L.pw_b200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
    part = C.c_void_p()
    engine._check(L.pw_b200_frame_groupby_partial(C.byref(bq.q), frame.handle, world, C.byref(part)))
    try:
        row_words = L.pw_b200_partial_row_bytes(part) // 8
        offs = (C.c_int64 * (world + 1))()
        engine._check(L.pw_b200_partial_offsets(part, offs))
        counts = [offs[i + 1] - offs[i] for i in range(world)]
        n_rows = offs[world]
        dev = torch.device("cuda", torch.cuda.current_device())
        send = torch.empty(max(1, n_rows * row_words), dtype=torch.int64, device=dev)
        engine._check(L.pw_b200_partial_copy_rows(part, C.c_void_p(send.data_ptr())))
    finally:
        L.pw_b200_partial_free(part)
    if nccl:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    recv = exchange_rows(send[: n_rows * row_words], counts, row_words)
    if nccl:
        e1.record()
        _events.append((e0, e1))
        _stats["calls"] += 1
        _stats["bytes"] += int(n_rows * row_words * 8)
    torch.cuda.synchronize()
    n_recv = recv.numel() // row_words
    cap = len(plan.keys) + len(plan.aggs) + 4
    out_arrays = (engine.ArrowArray * cap)()
    out_schemas = (engine.ArrowSchema * cap)()
    n_out = C.c_size_t(cap)
    engine._check(L.pw_b200_merge_partials(C.byref(bq.q), frame.handle, C.c_void_p(recv.data_ptr() if n_recv else 0), n_recv,
                                           out_arrays, out_schemas, C.byref(n_out)))
    names, cols = engine._import_columns(out_arrays, out_schemas, n_out.value)
    cols = engine._restore_string_types(names, cols, frame.table_schema, plan.keys)
    return pa.Table.from_arrays(cols, names=names)
