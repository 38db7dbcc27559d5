"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: variable-size key-hash all-to-all + merge of
partial aggregates.  The CUDA partial/merge kernels are exercised by tests/test_gpu_multigpu.py; here the per-rank
partial and the merge are computed by the CPU oracle so that routing, split sizes and combine semantics are checked
without a GPU."""
import os
import socket

import numpy as np
import pyarrow as pa
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import polaroid_b200 as pw
from oracle import oracle
from polaroid_b200 import multigpu
from tests import golden_util as G


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mix(k):
    k = k.astype(np.uint64)
    k ^= k >> np.uint64(32)
    k *= np.uint64(0xd6e8feb86659fd93)
    k ^= k >> np.uint64(32)
    return k


def _worker(rank, world, port, n, groups, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        keys = rng.integers(0, groups, n, dtype=np.int64)
        vals = rng.integers(-50, 50, n, dtype=np.int64)
        lo, hi = n * rank // world, n * (rank + 1) // world
        shard = pa.table({"key": keys[lo:hi], "value": vals[lo:hi]})
        # phase 1: local partial aggregates (sum, len, min, max, first-row index)
        part = oracle.collect(pw.LazyFrame(shard).group_by("key", maintain_order=True).agg(
            pw.col("value").sum().alias("s"), pw.len().alias("n"), pw.col("value").min().alias("lo"),
            pw.col("value").max().alias("hi")))
        pk = part["key"].to_numpy()
        rows = np.stack([pk, part["s"].to_numpy(), part["n"].to_numpy().astype(np.int64), part["lo"].to_numpy(),
                         part["hi"].to_numpy()], axis=1).astype(np.int64)
        owner = (_mix(pk) % np.uint64(world)).astype(np.int64)
        order = np.argsort(owner, kind="stable")
        rows = rows[order]
        counts = [int((owner == r).sum()) for r in range(world)]
        recv = multigpu.exchange_rows(torch.from_numpy(rows.reshape(-1).copy()), counts, 5).numpy().reshape(-1, 5)
        # phase 2: combine
        t = pa.table({"key": recv[:, 0], "s": recv[:, 1], "n": recv[:, 2], "lo": recv[:, 3], "hi": recv[:, 4]})
        merged = oracle.collect(pw.LazyFrame(t).group_by("key", maintain_order=True).agg(
            pw.col("s").sum().alias("s"), pw.col("n").sum().alias("n"), pw.col("lo").min().alias("lo"),
            pw.col("hi").max().alias("hi")))
        assert all((_mix(merged["key"].to_numpy()) % np.uint64(world)).astype(np.int64) == rank)
        q.put((rank, merged.to_pydict()))
    finally:
        dist.destroy_process_group()


def test_key_hash_all_to_all_merge_world2():
    world, n, groups = 2, 20_000, 257
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, groups, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = pa.concat_tables([pa.table(d) for _, d in sorted(outs)]).sort_by("key")
    rng = np.random.default_rng(5)
    keys = rng.integers(0, groups, n, dtype=np.int64)
    vals = rng.integers(-50, 50, n, dtype=np.int64)
    full = pa.table({"key": keys, "value": vals})
    want = oracle.collect(pw.LazyFrame(full).group_by("key").agg(
        pw.col("value").sum().alias("s"), pw.len().alias("n"), pw.col("value").min().alias("lo"),
        pw.col("value").max().alias("hi"))).sort_by("key")
    assert got["key"].to_pylist() == want["key"].to_pylist()
    assert got["s"].to_pylist() == want["s"].to_pylist()
    assert got["n"].to_pylist() == [int(x) for x in want["n"].to_pylist()]
    assert got["lo"].to_pylist() == want["lo"].to_pylist() and got["hi"].to_pylist() == want["hi"].to_pylist()
