#!/usr/bin/env python
"""bench.py — headline benchmark of the filter -> group_by -> agg hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c1|c3|c4]

A "step" is one pass of the hot path over one batch of synthetic input.  At N=1 the workload is
BASELINE.json configs[1]: 1e8 rows, int64 key with 1e3 groups, f64 sum/mean/min/max (C2).
  value     rows/s with the inputs already resident in HBM (CUDA events on the library's stream)
  e2e       rows/s through the host-facing C ABI call (pinned host Arrow buffers in, host result out;
            H2D and D2H inside the timed region)
  roofline  dominant kernel: algorithmic bytes per launch / its CUDA-event duration vs the measured HBM peak
  cpu_baseline  the CPU oracle port timed on this box's host cores on a bounded sample
Under torchrun (N>1) every rank holds its own shard (weak scaling); partial aggregates are exchanged by
key hash with one NCCL all-to-all and merged.
--impl reference times the reference-semantics CPU port (oracle/, all host threads) on the same workload.
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

WORKLOADS = {
    "c2": dict(rows=100_000_000, groups=1_000, bytes_per_row=16.0,
               name="C2 low-cardinality group_by: 1e8 rows, int64 key x 1e3 groups, f64 sum/mean/min/max"),
    # the other BASELINE configs: parity-test shapes, runnable here for the per-config table in profiles/
    "c1": dict(rows=6_001_215, groups=4, bytes_per_row=72.0,
               name="C1 TPC-H Q1 shape, synthetic lineitem SF1: filter shipdate + group_by(returnflag, linestatus) 8 aggregates"),
    "c3": dict(rows=100_000_000, groups=10_000_000, bytes_per_row=16.125,
               name="C3 high-cardinality group_by: 1e8 rows, 1e7 int64 keys, 5% null f64, sum/mean/min/max/count/first/last"),
    "c4e": dict(rows=200_000_000, groups=None, bytes_per_row=24.0,
                name="C4 keys-empty variant (what polars-timeseries calls): group_by_dynamic 1m over 2e8 sorted ticks, no group_by"),
    "c4": dict(rows=200_000_000, groups=None, bytes_per_row=28.0,
               name="C4 OHLCV group_by_dynamic 1m by symbol: 2e8 sorted ticks (1e9 in BASELINE; reduced for host RAM), 100 symbols"),
}


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


def make_workload(name: str, rows: int, rank: int):
    """-> (pyarrow table in pageable host memory, plan)"""
    import pyarrow as pa
    import polaroid_b200 as pw
    from tests import synth
    if name == "c1":
        t = synth.lineitem(rows, seed=1 + rank)
        return t, synth.q1_query(t).plan
    if name == "c3":
        t = synth.c3_table(rows, 10_000_000 if rows >= 10_000_000 else max(rows // 10, 1), seed=3 + rank)
        q = pw.LazyFrame(t).group_by("key").agg(
            pw.col("value").sum().alias("sum"), pw.col("value").mean().alias("mean"), pw.col("value").min().alias("min"),
            pw.col("value").max().alias("max"), pw.col("value").count().alias("count"), pw.col("value").first().alias("first"),
            pw.col("value").last().alias("last"))
        return t, q.plan
    if name in ("c4", "c4e"):
        t = synth.ohlcv(rows, n_symbols=100, seed=4 + rank, mean_gap_us=1000)
        return t, synth.ohlcv_query(t, by_symbol=(name == "c4")).plan
    raise ValueError(name)


def pinned_table(keys: np.ndarray, vals: np.ndarray):
    """pyarrow Table whose buffers live in pinned host memory (so H2D runs at PCIe speed)."""
    import pyarrow as pa
    import torch
    tk = torch.from_numpy(keys).pin_memory()
    tv = torch.from_numpy(vals).pin_memory()
    ak = pa.Array.from_buffers(pa.int64(), len(keys), [None, pa.py_buffer(tk.numpy())])
    av = pa.Array.from_buffers(pa.float64(), len(vals), [None, pa.py_buffer(tv.numpy())])
    return pa.table({"key": ak, "value": av}), (tk, tv)


def run_reference(args, rank: int, world: int):
    """The reference-semantics CPU engine (oracle port; the Rust reference cannot be built in this image)
    with all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    import pyarrow as pa
    from oracle import oracle
    from polaroid_b200.plan import LazyResult
    w = WORKLOADS["c2"]
    threads = oracle.max_threads()
    sample = int(os.environ.get("PW_REF_SAMPLE_ROWS", 20_000_000))
    keys, vals = make_c2_numpy(sample, w["groups"], seed=2)
    t = pa.table({"key": keys, "value": vals})
    q = LazyResult(t, c2_plan())
    for _ in range(args.warmup):
        oracle.collect(q, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = oracle.collect(q, n_threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    assert out.num_rows == w["groups"]
    v = sample / dt
    line = {"impl": "reference", "metric": "rows_per_sec", "value": v, "unit": "rows/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "rows_per_step": sample, "groups": w["groups"]},
            "cpu_baseline": {"value": v, "unit": "rows/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} rows of the C2 generator per step (oracle/pw_oracle.c, {threads} threads)"},
            "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default: the workload's)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
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
    from polaroid_b200 import engine

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = engine.lib()
    engine._check(L.pw_b200_set_device(local_rank))
    stream = torch.cuda.Stream()
    L.pw_b200_set_stream(stream.cuda_stream)

    w = WORKLOADS[args.workload]
    rows = args.rows or w["rows"]
    if args.workload == "c2":
        keys, vals = make_c2_numpy(rows, w["groups"], seed=2 + rank)
        host_table, _pins = pinned_table(keys, vals)
        plan = c2_plan()
    else:
        host_table, plan = make_workload(args.workload, rows, rank)
        args.no_cpu_baseline = True   # the CPU baseline leg is defined for the headline workload
    warm = max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM ----------------
    frame = engine.DeviceFrame(host_table)
    if world > 1:
        from polaroid_b200 import multigpu
        step = lambda: multigpu.group_by_sharded(frame, plan, rank, world, row_offset=rank * rows)
    else:
        step = lambda: frame.group_by(plan)
    for _ in range(warm):
        out = step()
    launches = engine.last_timings()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms = []
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            out = step()
            scan_ms.append(engine.last_timings()["scan_kernel_ms"])
        ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1) / max(1, args.steps)
    if world > 1:
        tms = torch.tensor([ms], device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = rows * world / (ms * 1e-3)
    tm = engine.last_timings()

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
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    if world > 1:
        te = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
    h2d = int(rows * w["bytes_per_row"])
    d2h = sum(b.size for c in res.columns for ch in c.chunks for b in ch.buffers() if b is not None)

    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = measured_peak_gbs()
    k_ms = float(np.mean(scan_ms))
    algo_bytes = rows * w["bytes_per_row"] + (w["groups"] or tm["n_groups"]) * 40
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    # DRAM traffic of the same kernel on the same workload from the committed ncu --set full capture (per launch)
    traffic = None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_c2_traffic.json")) as f:
            tj = json.load(f)
        if args.workload == tj["workload"] and rows == tj["rows"]:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        traffic = None

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        from polaroid_b200.plan import LazyResult
        threads = oracle.max_threads()
        sample = min(rows, int(os.environ.get("PW_REF_SAMPLE_ROWS", 20_000_000)))
        tq = LazyResult(pa.table({"key": keys[:sample], "value": vals[:sample]}), plan)
        oracle.collect(tq, n_threads=threads)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            ref = oracle.collect(tq, n_threads=threads)
        cdt = (time.perf_counter() - t0) / reps
        cpu = {"value": sample / cdt, "unit": "rows/s", "cores": threads, "kind": "port",
               "sample": f"first {sample} rows of the same input, oracle/pw_oracle.c with {threads} threads, {reps} reps"}

    if rank == 0:
        line = {
            "metric": "rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "rows_per_gpu": rows, "groups": w["groups"] or int(tm["n_groups"]),
                       "l2": f"inputs ({rows * w['bytes_per_row'] / 1e9:.2f} GB per step) are larger than the 126 MB L2; no explicit flush",
                       "strategy": {1: "hot table + spill tier", 2: "HBM table", 3: "segmented", 4: "hot table, dense ids + spill tier", 5: "radix partition + hot table"}.get(tm["strategy"]),
                       "parallelism": f"rows sharded over {world} GPU(s); partial aggregates merged by key hash"},
            "clocks": clocks,
            "e2e": {"value": rows * world / e2e_s, "unit": "rows/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(launches) * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": ("pw_seg_jit" if tm["strategy"] == 3 else "pw_scan_jit") if tm["reserved"] else "pw::scan_kernel",
                         "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                         "whole_step_frac": (algo_bytes / (ms * 1e-3) / 1e9) / peak},
            "cpu_baseline": cpu,
            "phases_ms": {k: tm[k] for k in ("estimate_ms", "scan_ms", "scan_kernel_ms", "finalize_ms", "d2h_ms", "total_device_ms", "host_ms", "partition_ms")},
            "spilled_rows": tm["spilled_rows"], "table_slots": tm["table_slots"], "n_groups": tm["n_groups"],
            "jit": bool(tm["reserved"]), "retries": tm["retries"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
