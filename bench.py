#!/usr/bin/env python
"""bench.py — benchmark of the filter -> group_by -> agg hot path on BASELINE.json's configurations.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--configs all|none|c1,c3,...]

A "step" is one pass of the hot path over one batch of synthetic input.

HEADLINE (the JSON line's own keys): BASELINE.json configs[1] = C2, 1e8 rows per GPU, int64 key with 1e3 groups,
f64 sum/mean/min/max.
  value         rows/s with the inputs already resident in HBM (CUDA events on the library's stream, max over ranks)
  e2e           rows/s through the host-facing C ABI call (pinned host Arrow buffers in, host result out; H2D and D2H
                inside the timed region)
  roofline      dominant kernel: algorithmic bytes per launch / its CUDA-event duration vs the measured HBM peak
  cpu_baseline  the CPU oracle port (oracle/, all host threads) timed on the SAME 1e8 rows; its result is also the
                checker: the GPU result of the timed configuration is compared with it ("check")
"configs" (same JSON line): the other BASELINE configurations, generated on the device and wrapped zero-copy with
pw_b200_frame_from_device, each with rows/s, the dominant kernel's roofline fraction against both denominators
(measured copy peak and the nominal 8 TB/s) and a self-check:
  c2_hash  C2 with dense ids disabled (the general hash-index path)
  c1       TPC-H Q1 shape, synthetic lineitem SF1 (6 001 215 rows, string-view keys)           checked: oracle, all rows
  c3       high cardinality: 1e8 rows, 1e7 keys, 5 % null values, null-key group, 7 aggregates  checked: invariants + oracle on a key sample
  c4       OHLCV group_by_dynamic 1m BY SYMBOL over 1e9 sorted ticks                            checked: invariants + oracle on a prefix
  c4e      the same without keys (what polars-timeseries calls)                                 checked: invariants + oracle on a prefix
  c5       TPC-H Q1 SF100 (600 037 902 rows TOTAL, strong scaling: rows split over the GPUs)    checked: per-rank oracle-free closed forms + merged counts
  a2a      (N > 1) C3-shaped 1e7-key run per GPU through the general key-hash all_to_all exchange, NCCL time separate
Under torchrun (N > 1) every rank holds its own shard (C2: weak scaling); partial aggregates are exchanged by key hash
over NCCL and merged; the merged C2 table is gathered to rank 0 and compared with the per-shard oracle
results merged on the host.
--impl reference times the reference-semantics CPU port (oracle/, all host threads) on the same C2 workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NOMINAL_GBS = 8000.0  # north star's "~8 TB/s"
Q1_SF1_ROWS = 6_001_215
Q1_SF100_ROWS = 600_037_902

WORKLOADS = {
    "c2": dict(rows=100_000_000, groups=1_000, bytes_per_row=16.0,
               name="C2 low-cardinality group_by: 1e8 rows, int64 key x 1e3 groups, f64 sum/mean/min/max"),
    "c2_hash": dict(rows=100_000_000, groups=1_000, bytes_per_row=16.0,
                    name="C2 with dense ids disabled (general hash-index hot table)"),
    "c1": dict(rows=Q1_SF1_ROWS, groups=4, bytes_per_row=72.0,
               name="C1 TPC-H Q1 shape, synthetic lineitem SF1: filter shipdate + group_by(returnflag, linestatus) 8 aggregates"),
    "c3": dict(rows=100_000_000, groups=10_000_000, bytes_per_row=16.125,
               name="C3 high-cardinality group_by: 1e8 rows, 1e7 int64 keys, 5% null f64, sum/mean/min/max/count/first/last"),
    "c4e": dict(rows=1_000_000_000, groups=None, bytes_per_row=24.0,
                name="C4 keys-empty variant (what polars-timeseries calls): group_by_dynamic 1m over 1e9 sorted ticks"),
    "c4": dict(rows=1_000_000_000, groups=None, bytes_per_row=28.0,
               name="C4 OHLCV group_by_dynamic 1m by symbol: 1e9 sorted ticks, 100 symbols"),
    "c5": dict(rows=Q1_SF100_ROWS, groups=4, bytes_per_row=72.0,
               name="C5 TPC-H Q1 SF100 (600 037 902 rows total), rows split contiguously over the GPUs"),
    "a2a": dict(rows=100_000_000, groups=10_000_000, bytes_per_row=16.125,
                name="C3-shaped multi-GPU run: 1e8 rows per GPU, 1e7 keys, general key-hash all_to_all exchange"),
}
STRATEGY = {1: "hot table + spill tier", 2: "HBM table", 3: "segmented sorted windows", 4: "hot table, dense ids + spill tier",
            5: "radix partition + hot table", 6: "sorted windows by key (dense ids per window)",
            7: "dense ids bucketed per tile, accumulators in registers + spill tier",
            8: "sorted-key runs in registers", 9: "key-index ids bucketed per tile, accumulators in registers + spill tier",
            10: "two-level radix partition (staged writes) + one shared-memory table per partition"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region.

    NVML in-process (nvidia_ml_py) at ~1 kHz so that even a 20 ms timed region gets samples; falls back to an
    `nvidia-smi -lms` child process when NVML cannot be loaded."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc, self.h, self.stop_flag = gpu_index, [], None, None, False
        self.sm, self.reasons, self.max_mhz = [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    @staticmethod
    def _physical_index(i: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.BAD.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.001)

    def start(self):
        if self.h is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.h is not None:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------
# the headline workload (host-generated: the e2e leg and the CPU leg need host buffers anyway)
# ------------------------------------------------------------------------------------------------------
def make_c2_numpy(rows: int, groups: int, seed: int):
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, groups, rows, dtype=np.int64)
    vals = rng.random(rows) * 100.0
    return keys, vals


def c2_plan():
    import polaroid_b200 as pw
    from polaroid_b200 import plan as P
    aggs = [pw.col("value").sum().alias("sum"), pw.col("value").mean().alias("mean"),
            pw.col("value").min().alias("min"), pw.col("value").max().alias("max")]
    return P.GroupByPlan(keys=["key"], aggs=[a.spec() for a in aggs], maintain_order=False)


def pinned_table(keys: np.ndarray, vals: np.ndarray):
    """pyarrow Table whose buffers live in pinned host memory (so H2D runs at PCIe speed)."""
    import pyarrow as pa
    import torch
    tk = torch.from_numpy(keys).pin_memory()
    tv = torch.from_numpy(vals).pin_memory()
    ak = pa.Array.from_buffers(pa.int64(), len(keys), [None, pa.py_buffer(tk.numpy())])
    av = pa.Array.from_buffers(pa.float64(), len(vals), [None, pa.py_buffer(tv.numpy())])
    return pa.table({"key": ak, "value": av}), (tk, tv)


def cpu_c2(keys, vals, plan, threads: int, reps: int, warmup: int):
    """oracle port over host arrays: (seconds per pass, result table)."""
    import pyarrow as pa
    from oracle import oracle
    from polaroid_b200.plan import LazyResult
    tq = LazyResult(pa.table({"key": keys, "value": vals}), plan)
    out = None
    for _ in range(warmup):
        out = oracle.collect(tq, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(reps):
        out = oracle.collect(tq, n_threads=threads)
    return (time.perf_counter() - t0) / max(reps, 1), out


def compare_c2(got, want, rtol=1e-12) -> str:
    """GPU result vs oracle result of the C2 query (any row order).  Bit-exact key set, min, max; sum/mean within rtol."""
    g = got.sort_by("key").to_pydict()
    w = want.sort_by("key").to_pydict()
    assert g["key"] == w["key"], "group keys differ"
    for name in ("min", "max"):
        assert g[name] == w[name], f"{name} differs"
    for name in ("sum", "mean"):
        a, b = np.array(g[name]), np.array(w[name])
        err = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
        assert err <= rtol, f"{name}: relative error {err:.3e} > {rtol}"
    return f"ok: {len(g['key'])} groups == oracle (keys/min/max exact, sum/mean rel <= {rtol})"


def run_reference(args, rank: int, world: int):
    """The reference-semantics CPU engine (oracle port; the Rust reference cannot be built in this image) with all
    host threads, on the SAME workload as the GPU arm: C2, 1e8 rows per step."""
    if rank != 0:
        return
    from oracle import oracle
    w = WORKLOADS["c2"]
    threads = oracle.max_threads()
    rows = int(os.environ.get("PW_REF_ROWS", w["rows"]))
    keys, vals = make_c2_numpy(rows, w["groups"], seed=2)
    dt, out = cpu_c2(keys, vals, c2_plan(), threads, args.steps, args.warmup)
    assert out.num_rows == w["groups"]
    v = rows / dt
    line = {"impl": "reference", "metric": "rows_per_sec", "value": v, "unit": "rows/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "rows_per_step": rows, "groups": w["groups"]},
            "cpu_baseline": {"value": v, "unit": "rows/s", "cores": threads, "kind": "port",
                             "sample": f"all {rows} rows of the C2 generator per step (oracle/pw_oracle.c: thread-local tables over row "
                                       f"shards + merge, {threads} threads; inputs read in place)"},
            "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# device-generated configurations
# ------------------------------------------------------------------------------------------------------
def _bitmap(valid):
    """bool tensor [n] -> (LSB-first packed uint8 tensor, null count)"""
    import torch
    n = valid.numel()
    pad = (-n) % 8
    v = valid
    if pad:
        v = torch.cat([valid, torch.zeros(pad, dtype=torch.bool, device=valid.device)])
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=valid.device)
    packed = (v.view(-1, 8).to(torch.uint8) * w).sum(dim=1, dtype=torch.uint8)
    return packed, int(n - int(valid.sum().item()))


def _col(name, typ, t, validity=None, nulls=0):
    """one column descriptor for engine.DeviceFrame.from_device out of a torch tensor"""
    length = t.shape[0]
    return (name, typ, length, t.data_ptr(), validity.data_ptr() if validity is not None else 0, nulls, (t, validity))


def gen_lineitem_device(rows: int, seed: int):
    """tests/synth.py::lineitem on the device (same columns, dtypes and proportions); flags as inline Utf8View."""
    import pyarrow as pa
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    u = torch.rand(rows, generator=g, device="cuda")
    u2 = torch.rand(rows, generator=g, device="cuda")
    flag = torch.where(u < 0.25, 65, torch.where(u < 0.50, 82, 78)).to(torch.int32)            # A R N
    status = torch.where(flag == 78, torch.where(u2 < 0.014, 70, 79), 70).to(torch.int32)      # N -> F/O, else F
    del u, u2

    def views(code):
        v = torch.zeros((rows, 4), dtype=torch.int32, device="cuda")
        v[:, 0] = 1
        v[:, 1] = code
        return v

    day = 86_400_000_000
    start = int(np.datetime64("1992-01-02", "us").astype(np.int64))
    span_days = int((np.datetime64("1998-12-01", "us") - np.datetime64("1992-01-02", "us")).astype(np.int64)) // day
    ship = start + torch.randint(0, span_days, (rows,), generator=g, device="cuda", dtype=torch.int64) * day
    qty = torch.randint(1, 51, (rows,), generator=g, device="cuda", dtype=torch.int64)
    price = torch.round((900 + torch.rand(rows, generator=g, device="cuda", dtype=torch.float64) * 104_100) * 100) / 100
    disc = torch.randint(0, 11, (rows,), generator=g, device="cuda", dtype=torch.int64).to(torch.float64) / 100.0
    tax = torch.randint(0, 9, (rows,), generator=g, device="cuda", dtype=torch.int64).to(torch.float64) / 100.0
    tensors = dict(l_shipdate=ship, l_returnflag=views(flag), l_linestatus=views(status), l_quantity=qty,
                   l_extendedprice=price, l_discount=disc, l_tax=tax, _flag=flag, _status=status)
    cols = [_col("l_shipdate", pa.timestamp("us"), ship), _col("l_returnflag", pa.string_view(), tensors["l_returnflag"]),
            _col("l_linestatus", pa.string_view(), tensors["l_linestatus"]), _col("l_quantity", pa.int64(), qty),
            _col("l_extendedprice", pa.float64(), price), _col("l_discount", pa.float64(), disc), _col("l_tax", pa.float64(), tax)]
    return cols, tensors


def lineitem_host_table(tensors, lo=0, hi=None):
    import pyarrow as pa
    sl = slice(lo, hi)
    flag = tensors["_flag"][sl].cpu().numpy().astype(np.uint8).view("S1").astype("U1")
    status = tensors["_status"][sl].cpu().numpy().astype(np.uint8).view("S1").astype("U1")
    return pa.table({
        "l_shipdate": pa.array(tensors["l_shipdate"][sl].cpu().numpy(), type=pa.int64()).cast(pa.timestamp("us")),
        "l_returnflag": pa.array(flag), "l_linestatus": pa.array(status),
        "l_quantity": pa.array(tensors["l_quantity"][sl].cpu().numpy()),
        "l_extendedprice": pa.array(tensors["l_extendedprice"][sl].cpu().numpy()),
        "l_discount": pa.array(tensors["l_discount"][sl].cpu().numpy()), "l_tax": pa.array(tensors["l_tax"][sl].cpu().numpy())})


def q1_plan():
    import pyarrow as pa
    from tests import synth
    empty = pa.table({"l_shipdate": pa.array([], type=pa.timestamp("us")), "l_returnflag": pa.array([], type=pa.string()),
                      "l_linestatus": pa.array([], type=pa.string()), "l_quantity": pa.array([], type=pa.int64()),
                      "l_extendedprice": pa.array([], type=pa.float64()), "l_discount": pa.array([], type=pa.float64()),
                      "l_tax": pa.array([], type=pa.float64())})
    return synth.q1_query(empty).plan


def gen_c3_device(rows: int, groups: int, seed: int):
    import pyarrow as pa
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    keys = torch.randint(0, groups, (rows,), generator=g, device="cuda", dtype=torch.int64)
    kvalid = torch.ones(rows, dtype=torch.bool, device="cuda")
    kvalid[torch.randint(0, rows, (max(1, rows // 1000),), generator=g, device="cuda")] = False   # one null-key group
    vals = torch.rand(rows, generator=g, device="cuda", dtype=torch.float64) * 100.0
    vvalid = torch.rand(rows, generator=g, device="cuda") >= 0.05
    kb, kn = _bitmap(kvalid)
    vb, vn = _bitmap(vvalid)
    cols = [_col("key", pa.int64(), keys, kb, kn), _col("value", pa.float64(), vals, vb, vn)]
    return cols, dict(key=keys, kvalid=kvalid, value=vals, vvalid=vvalid)


def c3_plan():
    import polaroid_b200 as pw
    from polaroid_b200 import plan as P
    c = pw.col("value")
    aggs = [c.sum().alias("sum"), c.mean().alias("mean"), c.min().alias("min"), c.max().alias("max"), c.count().alias("count"),
            c.first().alias("first"), c.last().alias("last")]
    return P.GroupByPlan(keys=["key"], aggs=[a.spec() for a in aggs], maintain_order=False)


def gen_ohlcv_device(rows: int, seed: int, chunk: int = 125_000_000):
    """tests/synth.py::ohlcv on the device: ts sorted (exponential gaps, mean 1 ms), 100 symbols, random-walk price."""
    import pyarrow as pa
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    ts = torch.empty(rows, dtype=torch.int64, device="cuda")
    price = torch.empty(rows, dtype=torch.float64, device="cuda")
    t0 = int(np.datetime64("2024-01-02T09:30:00", "us").astype(np.int64))
    p0 = 100.0
    for lo in range(0, rows, chunk):   # chunked: the float temporaries of a 1e9-row cumsum stay small
        hi = min(rows, lo + chunk)
        gaps = torch.empty(hi - lo, dtype=torch.float64, device="cuda").exponential_(1.0 / 1000.0, generator=g).to(torch.int64)
        torch.cumsum(gaps, 0, out=ts[lo:hi])
        ts[lo:hi] += t0
        t0 = int(ts[hi - 1].item())
        steps = torch.empty(hi - lo, dtype=torch.float64, device="cuda").normal_(0.0, 0.01, generator=g)
        torch.cumsum(steps, 0, out=price[lo:hi])
        price[lo:hi] += p0
        p0 = float(price[hi - 1].item())
        del gaps, steps
    sym = torch.randint(0, 100, (rows,), generator=g, device="cuda", dtype=torch.int32)   # read as uint32 (values < 100)
    vol = torch.randint(1, 1001, (rows,), generator=g, device="cuda", dtype=torch.int64)
    cols = [_col("ts", pa.timestamp("us"), ts), _col("symbol", pa.uint32(), sym), _col("price", pa.float64(), price),
            _col("volume", pa.int64(), vol)]
    return cols, dict(ts=ts, symbol=sym, price=price, volume=vol)


def ohlcv_plan(by_symbol: bool):
    import pyarrow as pa
    from tests import synth
    empty = pa.table({"ts": pa.array([], type=pa.timestamp("us")), "symbol": pa.array([], type=pa.uint32()),
                      "price": pa.array([], type=pa.float64()), "volume": pa.array([], type=pa.int64())})
    return synth.ohlcv_query(empty, by_symbol=by_symbol).plan


def ohlcv_host_table(tensors, n):
    import pyarrow as pa
    return pa.table({"ts": pa.array(tensors["ts"][:n].cpu().numpy(), type=pa.int64()).cast(pa.timestamp("us")),
                     "symbol": pa.array(tensors["symbol"][:n].cpu().numpy().view(np.uint32)),
                     "price": pa.array(tensors["price"][:n].cpu().numpy()), "volume": pa.array(tensors["volume"][:n].cpu().numpy())})


def slice_cols(cols, n):
    """the first n rows of a device frame description (same buffers, shorter length; no validity here)"""
    return [(name, typ, n, vptr, nptr, nulls, extra) for (name, typ, _len, vptr, nptr, nulls, extra) in cols]


# ---- checks -----------------------------------------------------------------------------------------------
def _assert_tables(got, want, rtol, sort_by=None):
    from tests import golden_util as G
    G.assert_tables_equal(got, want, sort_by=sort_by, rtol=rtol)


def check_c1(res, tensors, plan) -> str:
    from oracle import oracle
    from polaroid_b200.plan import LazyResult
    host = lineitem_host_table(tensors)
    want = oracle.collect(LazyResult(host, plan), n_threads=oracle.max_threads())
    import pyarrow as pa
    for name in ("l_returnflag", "l_linestatus"):   # Utf8View keys come back as the library's view -> large_string export
        res = res.set_column(res.column_names.index(name), name, res.column(name).cast(pa.string()))
    _assert_tables(res, want, 1e-12, sort_by=["l_returnflag", "l_linestatus"])
    return f"ok: all {host.num_rows} rows vs oracle (counts and integer sums exact, f64 sums/means rel <= 1e-12)"


def check_c5_shard(res, tensors, plan, rows) -> str:
    """Q1 over a 6e8/N-row shard: closed forms computed with torch on the device (independent library code), exact for
    the integer columns, 1e-9 for the f64 totals (torch's own summation order)."""
    import torch
    d = res.to_pydict()
    cutoff = int(np.datetime64("1998-09-02", "us").astype(np.int64))
    keep = tensors["l_shipdate"] <= cutoff
    n_keep = int(keep.sum().item())
    assert sum(d["count_order"]) == n_keep, f"count {sum(d['count_order'])} != {n_keep}"
    assert sum(d["sum_qty"]) == int(tensors["l_quantity"][keep].sum().item()), "sum_qty differs"
    tot = float(tensors["l_extendedprice"][keep].sum().item())
    got = float(np.sum(d["sum_base_price"]))
    assert abs(got - tot) <= 1e-9 * abs(tot), f"sum_base_price {got} vs {tot}"
    return f"ok: {rows} rows, {len(d['count_order'])} groups; kept-row count and sum_qty exact, price total rel <= 1e-9 (torch)"


def check_c3(res, tensors, plan, rows) -> str:
    import pyarrow as pa
    import torch
    from oracle import oracle
    from polaroid_b200.plan import LazyResult
    key, kvalid, value, vvalid = tensors["key"], tensors["kvalid"], tensors["value"], tensors["vvalid"]
    d_count = np.asarray(res.column("count").to_numpy())
    n_valid_vals = int(vvalid.sum().item())
    assert int(d_count.astype(np.int64).sum()) == n_valid_vals, "sum of per-group counts != non-null values"
    n_groups = int(torch.unique(key[kvalid]).numel()) + (1 if int((~kvalid).sum().item()) else 0)
    assert res.num_rows == n_groups, f"{res.num_rows} groups != {n_groups}"
    tot = float(value[vvalid].sum().item())
    got = float(np.sum(res.column("sum").to_numpy(zero_copy_only=False)))
    assert abs(got - tot) <= 1e-9 * abs(tot), "sum of sums differs"
    # oracle on every row of a sample of groups (the null-key group included)
    sample = torch.randint(0, 10_000_000, (2000,), device="cuda", dtype=torch.int64)
    pick = (torch.isin(key, sample) & kvalid) | (~kvalid)
    hk = key[pick].cpu().numpy(); hkv = kvalid[pick].cpu().numpy()
    hv = value[pick].cpu().numpy(); hvv = vvalid[pick].cpu().numpy()
    sub = pa.table({"key": pa.array(hk, mask=~hkv), "value": pa.array(hv, mask=~hvv)})
    want = oracle.collect(LazyResult(sub, plan))
    keys = pa.array(np.unique(hk[hkv]))
    import pyarrow.compute as pc
    got_sub = res.filter(pc.or_kleene(pc.is_in(res.column("key"), value_set=keys), pc.is_null(res.column("key"))))
    _assert_tables(got_sub, want, 1e-12, sort_by=["key"])
    return (f"ok: {rows} rows -> {n_groups} groups; count/group-count invariants exact, total rel <= 1e-9; "
            f"{want.num_rows} sampled groups ({sub.num_rows} rows) == oracle incl. first/last and the null-key group")


def check_ohlcv(res_big, tensors, plan, frame_prefix_fn, by_symbol: bool, rows: int, prefix: int = 3_000_000) -> str:
    import pyarrow.compute as pc
    from oracle import oracle
    from polaroid_b200.plan import LazyResult
    tot_vol = int(tensors["volume"].sum().item())
    assert int(np.asarray(res_big.column("volume").to_numpy()).astype(np.int64).sum()) == tot_vol, "volume total differs"
    # oracle on a prefix, through the same kernels ...
    host = ohlcv_host_table(tensors, prefix)
    want = oracle.collect(LazyResult(host, plan))
    got = frame_prefix_fn(prefix)
    _assert_tables(got, want, 1e-12)
    # ... and the full run must agree with the prefix run on every window that ends inside the prefix
    last_ts = want.column("ts")[want.num_rows - 1]
    done = pc.less(want.column("ts"), last_ts)
    big_head = res_big.filter(pc.less(res_big.column("ts"), last_ts))
    if by_symbol:
        big_head = big_head.sort_by([("symbol", "ascending"), ("ts", "ascending")])
    _assert_tables(big_head, want.filter(done), 1e-12)
    return (f"ok: {rows} rows -> {res_big.num_rows} windows; volume total exact; first {prefix} rows == oracle "
            f"({want.num_rows} windows, row order included) and the full run agrees on them")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU of the headline workload (default 1e8)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="all", help="'all', 'none' or a comma list out of c2_hash,c1,c3,c4,c4e,c5,a2a")
    ap.add_argument("--config-steps", type=int, default=5)
    ap.add_argument("--scale", type=float, default=1.0, help="scale the row counts of the extra configurations (smoke runs)")
    ap.add_argument("--workload", default="c2", help="run ONE configuration as the headline (profiling): c2|c2_hash|c1|c3|c4|c4e|c5")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import pyarrow as pa
    import torch
    import torch.distributed as dist
    from polaroid_b200 import engine, multigpu

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = engine.lib()
    engine._check(L.pw_b200_set_device(local_rank))
    stream = torch.cuda.Stream()
    L.pw_b200_set_stream(stream.cuda_stream)
    peak, peak_src = measured_peak_gbs()
    warm = max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step, steps, warmup):
        """-> (ms per step [device events, max over ranks], kernel ms [mean], last result, timings of the last step)"""
        out = None
        for _ in range(warmup):
            out = step()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_ms = []
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for _ in range(steps):
                out = step()
                k_ms.append(engine.last_timings()["scan_kernel_ms"])
            ev1.record(stream)
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1) / max(1, steps))
        return ms, float(np.mean(k_ms)), out, engine.last_timings()

    def roofline(rows, bytes_per_row, out_rows, out_bytes_per_row, k_ms, step_ms, tm):
        algo = rows * bytes_per_row + out_rows * out_bytes_per_row
        ach = algo / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        return {"kernel": {3: "pw_seg_jit", 6: "pw_bucket_jit", 7: "pw_bucket_jit", 8: "pw_runs_jit", 9: "pw_bucket_jit", 10: "pw_radix_jit_m0..m3"}.get(tm["strategy"], "pw_scan_jit") if tm["reserved"] else "ahead-of-time kernel",
                "kernel_ms": k_ms, "algorithmic_bytes_per_launch": algo, "achieved_gbs": ach,
                "frac_measured": ach / peak, "frac_nominal_8tbs": ach / NOMINAL_GBS,
                "whole_step_frac_measured": (algo / (step_ms * 1e-3) / 1e9) / peak}

    # =====================================================================================================
    # headline: C2 (or the configuration named by --workload, for profiling one of the others in isolation)
    # =====================================================================================================
    if args.workload != "c2":
        args.configs = args.workload
    w = WORKLOADS["c2"]
    rows = args.rows or w["rows"]
    line = None
    if args.workload == "c2":
        keys, vals = make_c2_numpy(rows, w["groups"], seed=2 + rank)
        host_table, _pins = pinned_table(keys, vals)
        plan = c2_plan()
        frame = engine.DeviceFrame(host_table)
        if world > 1:
            step = lambda: multigpu.group_by_sharded(frame, plan, rank, world, row_offset=rank * rows)
        else:
            step = lambda: frame.group_by(plan)
        for _ in range(warm):
            out = step()
        launches = engine.last_timings()["kernel_launches"]
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        ms, k_ms, out, tm = timed(step, args.steps, 0)
        clocks = sampler.stop()
        value = rows * world / (ms * 1e-3)

        # ---------------- e2e: host Arrow buffers through the public C ABI call ----------------
        if world > 1:
            e2e_step = lambda: multigpu.group_by_sharded(engine.DeviceFrame(host_table), plan, rank, world, row_offset=rank * rows)
        else:
            e2e_step = lambda: engine.run_group_by(host_table, plan)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            res = e2e_step()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / max(1, args.e2e_steps))
        h2d = int(rows * w["bytes_per_row"])
        d2h = sum(b.size for c in res.columns for ch in c.chunks for b in ch.buffers() if b is not None)

        # ---------------- roofline of the dominant kernel ----------------
        algo_bytes = rows * w["bytes_per_row"] + w["groups"] * 40
        achieved = algo_bytes / (k_ms * 1e-3) / 1e9
        traffic = None   # DRAM traffic of the same kernel on the same workload from the committed ncu --set full capture
        for name in ("r02_c2_traffic.json", "r01_c2_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as f:
                    tj = json.load(f)
                if tj["workload"] == "c2" and rows == tj["rows"]:
                    traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                    break
            except Exception:
                pass

        # ---------------- CPU baseline + self-check (outside the timed regions) ----------------
        cpu, check = None, "skipped (--no-cpu-baseline)"
        if not args.no_cpu_baseline:
            from oracle import oracle
            threads = oracle.max_threads() if world == 1 else max(1, oracle.max_threads() // world)
            reps = 3 if world == 1 else 1
            cdt, ref = cpu_c2(keys, vals, plan, threads, reps, 1 if world == 1 else 0)
            if world == 1:
                cpu = {"value": rows / cdt, "unit": "rows/s", "cores": threads, "kind": "port", "ms_per_pass": cdt * 1e3,
                       "sample": f"all {rows} rows of the same input, oracle/pw_oracle.c with {threads} threads, {reps} passes "
                                 "(inputs read in place)"}
                check = compare_c2(out, ref) + "; e2e result: " + compare_c2(res, ref)
            else:
                # every rank owns a slice of the groups; rank 0 gathers them and the per-shard oracle tables, merges the
                # latter on the host (sum of sums, sum of counts, min of mins, max of maxes) and compares
                mine = (out.to_pydict(), ref.append_column("len", pa.array(
                    np.bincount(keys, minlength=w["groups"])[np.asarray(ref.column("key").to_numpy())])).to_pydict())
                gathered = [None] * world
                dist.all_gather_object(gathered, mine)
                if rank == 0:
                    got = {k: sum((g[0][k] for g in gathered), []) for k in gathered[0][0]}
                    assert len(set(got["key"])) == len(got["key"]), "a group is owned by two ranks"
                    acc = {}
                    for _g, r in gathered:
                        for k, s, mn, mx, ln in zip(r["key"], r["sum"], r["min"], r["max"], r["len"]):
                            a = acc.setdefault(k, [0.0, np.inf, -np.inf, 0])
                            a[0] += s; a[1] = min(a[1], mn); a[2] = max(a[2], mx); a[3] += ln
                    ks = sorted(acc)
                    want = pa.table({"key": ks, "sum": [acc[k][0] for k in ks], "mean": [acc[k][0] / acc[k][3] for k in ks],
                                     "min": [acc[k][1] for k in ks], "max": [acc[k][2] for k in ks]})
                    check = compare_c2(pa.table(got), want, rtol=1e-11) + f" (merged over {world} ranks; per-shard oracle tables merged on the host)"

        if rank == 0:
            line = {
                "metric": "rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["name"], "rows_per_gpu": rows, "groups": w["groups"],
                           "l2": f"inputs ({rows * w['bytes_per_row'] / 1e9:.2f} GB per step) are larger than the 126 MB L2; no explicit flush",
                           "strategy": STRATEGY.get(tm["strategy"]),
                           "parallelism": f"rows sharded over {world} GPU(s); partial aggregates merged by key hash"},
                "clocks": clocks,
                "e2e": {"value": rows * world / e2e_s, "unit": "rows/s", "h2d_bytes_per_step": h2d * world,
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3},
                "gpu_launches": int(launches) * args.steps,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "kernel": {7: "pw_bucket_jit", 9: "pw_bucket_jit"}.get(tm["strategy"], "pw_scan_jit") if tm["reserved"] else "pw::scan_kernel",
                             "kernel_ms": k_ms, "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                             "frac_nominal_8tbs": achieved / NOMINAL_GBS,
                             "whole_step_frac": (algo_bytes / (ms * 1e-3) / 1e9) / peak},
                "cpu_baseline": cpu,
                "check": check,
                "phases_ms": {k: tm[k] for k in ("estimate_ms", "scan_ms", "scan_kernel_ms", "finalize_ms", "d2h_ms", "total_device_ms", "host_ms", "partition_ms")},
                "spilled_rows": tm["spilled_rows"], "table_slots": tm["table_slots"], "n_groups": tm["n_groups"],
                "jit": bool(tm["reserved"]), "retries": tm["retries"],
            }
        # the hash-index variant of the same frame (dense ids off)
        want_cfg = [] if args.configs == "none" else (["c2_hash", "c1", "c3", "c4", "c4e", "c5", "a2a"] if args.configs == "all" else args.configs.split(","))
        configs = {}
        if "c2_hash" in want_cfg and world == 1:
            ms2, k2, out2, tm2 = timed(lambda: frame.group_by(plan, flags=engine.FLAG_NO_DENSE_IDS), args.config_steps, 3)
            e = {"workload": WORKLOADS["c2_hash"]["name"], "rows": rows, "ms_per_step": ms2, "rows_per_s": rows / (ms2 * 1e-3),
                 "strategy": STRATEGY.get(tm2["strategy"])}
            e.update(roofline(rows, 16.0, w["groups"], 40, k2, ms2, tm2))
            e["check"] = compare_c2(out2, ref) if not args.no_cpu_baseline else "skipped"
            configs["c2_hash"] = e
        frame.free()
        del frame, host_table, _pins, keys, vals
    else:
        want_cfg = [args.workload]
        configs = {}

    # =====================================================================================================
    # the other BASELINE configurations (device-generated, zero-copy frames)
    # =====================================================================================================
    S = args.scale
    cs, cw = args.config_steps, 2

    def entry(name, n_rows, ms, k_ms, tm, out_rows, out_bpr, bpr=None, **extra):
        wl = WORKLOADS[name]
        e = {"workload": wl["name"], "rows": n_rows, "ms_per_step": ms, "rows_per_s": n_rows / (ms * 1e-3),
             "strategy": STRATEGY.get(tm["strategy"]), "n_groups": int(tm["n_groups"]), "phases_ms": {k: tm[k] for k in ("estimate_ms", "scan_ms", "scan_kernel_ms", "partition_ms", "finalize_ms", "d2h_ms", "total_device_ms", "host_ms") if k in tm}}
        e.update(roofline(n_rows, bpr or wl["bytes_per_row"], out_rows, out_bpr, k_ms, ms, tm))
        e.update(extra)
        return e

    def guarded(name, fn):
        try:
            fn()
        except Exception as ex:  # a failing extra configuration must not take the headline line down with it
            configs[name] = {"workload": WORKLOADS[name]["name"], "error": f"{type(ex).__name__}: {ex}"[:600]}
        torch.cuda.empty_cache()

    if "c1" in want_cfg and world == 1:
        def run_c1():
            n = max(1000, int(Q1_SF1_ROWS * S))
            cols, tensors = gen_lineitem_device(n, seed=1)
            f = engine.DeviceFrame.from_device(cols)
            plan1 = q1_plan()
            ms1, k1, out1, tm1 = timed(lambda: f.group_by(plan1), max(cs, 20), 3)
            configs["c1"] = entry("c1", n, ms1, k1, tm1, out1.num_rows, 100, check=check_c1(out1, tensors, plan1))
            f.free()
        guarded("c1", run_c1)

    if "c3" in want_cfg and world == 1:
        def run_c3():
            n = max(100_000, int(100_000_000 * S))
            cols, tensors = gen_c3_device(n, max(1, n // 10), seed=3)
            f = engine.DeviceFrame.from_device(cols)
            plan3 = c3_plan()
            ms3, k3, out3, tm3 = timed(lambda: f.group_by(plan3), max(cs, 5), 3)   # warm-up 3: the pinned result pool grows by one 0.6 GB block in each of the first two calls
            dev_ms = tm3["scan_ms"]   # partition + scan: the whole device phase of the group-by
            configs["c3"] = entry("c3", n, ms3, dev_ms, tm3, out3.num_rows, 60, check=check_c3(out3, tensors, plan3, n),
                                  kernel_note="kernel_ms = histogram + two scatter passes + per-partition aggregation (whole device phase)")
            f.free()
        guarded("c3", run_c3)

    if ("c4" in want_cfg or "c4e" in want_cfg) and world == 1:
        def run_c4():
            n = max(200_000, int(1_000_000_000 * S))
            cols, tensors = gen_ohlcv_device(n, seed=4)
            f = engine.DeviceFrame.from_device(cols)
            for name, by_symbol in (("c4", True), ("c4e", False)):
                if name not in want_cfg:
                    continue
                planx = ohlcv_plan(by_symbol)
                msx, kx, outx, tmx = timed(lambda: f.group_by(planx), max(cs, 5), 3)

                def prefix_run(m, planx=planx):
                    fp = engine.DeviceFrame.from_device(slice_cols(cols, m))
                    try:
                        return fp.group_by(planx)
                    finally:
                        fp.free()
                configs[name] = entry(name, n, msx, kx, tmx, outx.num_rows, 48,
                                      check=check_ohlcv(outx, tensors, planx, prefix_run, by_symbol, n, prefix=min(3_000_000, n // 2)))
            f.free()
        guarded("c4", run_c4)

    if "c5" in want_cfg:
        def run_c5():
            total = max(1000 * world, int(Q1_SF100_ROWS * S))
            lo, hi = total * rank // world, total * (rank + 1) // world
            n = hi - lo
            cols, tensors = gen_lineitem_device(n, seed=100 + rank)
            f = engine.DeviceFrame.from_device(cols)
            plan5 = q1_plan()
            if world > 1:
                step5 = lambda: multigpu.group_by_sharded(f, plan5, rank, world, row_offset=lo)
            else:
                step5 = lambda: f.group_by(plan5)
            ms5, k5, out5, tm5 = timed(step5, cs, 2)
            local = f.group_by(plan5)   # this shard alone, for the closed-form check
            chk = check_c5_shard(local, tensors, plan5, n)
            if world > 1:
                gathered = [None] * world
                dist.all_gather_object(gathered, (out5.to_pydict(), local.to_pydict()))
                if rank == 0:
                    merged = {}
                    for g, _l in gathered:
                        for i in range(len(g["count_order"])):
                            merged[(g["l_returnflag"][i], g["l_linestatus"][i])] = (g["count_order"][i], g["sum_qty"][i])
                    want = {}
                    for _g, l in gathered:
                        for i in range(len(l["count_order"])):
                            k = (l["l_returnflag"][i], l["l_linestatus"][i])
                            c0, s0 = want.get(k, (0, 0))
                            want[k] = (c0 + l["count_order"][i], s0 + l["sum_qty"][i])
                    assert merged == want, "merged Q1 table != sum of the per-rank tables"
                    chk += f"; merged table over {world} ranks == sum of per-rank tables (counts, sum_qty exact)"
            e = entry("c5", total, ms5, k5, tm5, 4, 100, bpr=72.0 / world, check=chk, scaling="strong", rows_per_gpu=n,
                      per_gpu_kernel_gbs=n * 72.0 / (k5 * 1e-3) / 1e9 if k5 > 0 else None)
            e["rows_per_s"] = total / (ms5 * 1e-3)
            configs["c5"] = e
            f.free()
        guarded("c5", run_c5)

    if "a2a" in want_cfg and world > 1:
        def run_a2a():
            n = max(100_000, int(100_000_000 * S))
            cols, tensors = gen_c3_device(n, max(1, n // 10), seed=30 + rank)
            f = engine.DeviceFrame.from_device(cols)
            plan3 = c3_plan()
            a2a_step = lambda: multigpu.group_by_sharded(f, plan3, rank, world, row_offset=rank * n, force_all_to_all=True)
            for _ in range(3):   # the pinned result pool grows by one block in each of the first two calls; NCCL connects its peers in the first
                a2a_step()
            multigpu.reset_stats()
            msa, ka, outa, tma = timed(a2a_step, 3, 0)
            st = multigpu.stats()
            # every rank owns a disjoint slice of the keys: counts add up to the non-null values of ALL shards
            cnt = torch.tensor([int(np.asarray(outa.column("count").to_numpy()).astype(np.int64).sum()), int(tensors["vvalid"].sum().item()),
                                outa.num_rows], device="cuda", dtype=torch.int64)
            dist.all_reduce(cnt)
            assert int(cnt[0]) == int(cnt[1]), "all_to_all merge lost or duplicated rows"
            configs["a2a"] = entry("a2a", n * world, msa, tma["scan_ms"], tma, int(cnt[2].item()), 60, bpr=16.125 / world,
                                   check=f"ok: {int(cnt[2])} groups over {world} ranks; merged counts == non-null values of all shards ({int(cnt[1])})",
                                   nccl_ms_per_step=st.get("nccl_ms"), exchanged_bytes_per_rank=st.get("bytes"), rows_per_gpu=n)
            f.free()
        guarded("a2a", run_a2a)

    if rank == 0:
        if line is None:   # --workload X: the single configuration is the line
            name = args.workload
            e = configs.get(name, {"error": "configuration did not run"})
            line = {"metric": "rows_per_sec", "value": e.get("rows_per_s"), "unit": "rows/s", "n_gpus": world, "steps": args.config_steps,
                    "warmup": cw, "ms_per_step": e.get("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOADS[name]["name"]}, "detail": e}
        else:
            line["configs"] = configs
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
