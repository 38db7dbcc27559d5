"""Builds libpolarway_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libpolarway_b200.so")
SOURCES = [*[f"pw_launch_nc{nc}_kw{kw}.cu" for nc in (12, 4) for kw in (6, 4, 2, 1)], "pw_launch_nc4.cu", "pw_launch_nc12.cu", "pw_jit.cu", "pw_dynamic.cu", "pw_partial.cu", "pw_filter.cu", "pw_views.cu", "pw_engine.cu", "pw_capi.cu", "pw_plugin.cu", "pw_arrow.cpp"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
              "-Xptxas", "-v" if os.environ.get("PW_PTXAS_V") else "-O3"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "polarway_b200.h")]
    if any(os.path.getmtime(d) > t for d in deps):
        return True
    # a header edited WHILE a build was running leaves objects older than the header behind a library that is newer
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f not in SOURCES] + [os.path.join(HERE, "..", "include", "polarway_b200.h")]
    t_hdr = max(os.path.getmtime(h) for h in headers)
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.rsplit(".", 1)[0] + ".o")
        if os.path.exists(obj) and os.path.getmtime(obj) < max(t_hdr, os.path.getmtime(os.path.join(CSRC, src))):
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    # an object is rebuilt when its own source or any header (everything that is not one of SOURCES) is newer
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f not in SOURCES] + [os.path.join(HERE, "..", "include", "polarway_b200.h")]
    t_hdr = max(os.path.getmtime(h) for h in headers)
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(t_hdr, os.path.getmtime(os.path.join(CSRC, src))):
            continue
        cmd = ["nvcc", *NVCC_FLAGS, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
        if verbose and out:
            sys.stderr.write(out)
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
