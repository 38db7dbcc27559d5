"""Run bench.py for several workloads and print one compact line each (development helper)."""
import json, subprocess, sys

def main():
    workloads = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2", "c1", "c4", "c4e", "c3"]
    extra = sys.argv[2:]
    for w in workloads:
        p = subprocess.run([sys.executable, "bench.py", "--workload", w, "--steps", "10", "--warmup", "3", "--no-cpu-baseline", *extra],
                           capture_output=True, text=True)
        line = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(w, "FAILED", p.stderr[-800:])
            continue
        j = json.loads(line[-1])
        r, ph = j["roofline"], j["phases_ms"]
        print(f"{w:4s} step {j['ms_per_step']:8.3f} ms  kernel {r['kernel_ms']:8.3f} ms  frac {r['frac']:.3f}  e2e {j['e2e']['ms_per_step']:8.2f} ms  "
              f"call {ph['host_ms']:.3f} scan {ph['scan_ms']:.3f} part {ph.get('partition_ms', 0):.3f} est {ph['estimate_ms']:.3f} fin {ph['finalize_ms']:.3f} d2h {ph['d2h_ms']:.3f}  spilled {j['spilled_rows']} groups {j['n_groups']} "
              f"jit {j['jit']} strat {j['config']['strategy']} clocks {j['clocks'].get('sm_mhz')}", flush=True)

if __name__ == "__main__":
    main()
