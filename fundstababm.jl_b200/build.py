"""Builds csrc/libfsb_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(CSRC, "libfsb_b200.so")
SOURCES = ["fsb_abi.cu", "fsb_step_big.cu", "fsb_step_fixed.cu", "fsb_facts.cu"]  # (the second one: the same step kernels built for big shapes)
HEADERS = ["fsb_device.cuh", "fsb_state.h", "fsb_step.cuh", "fsb_passx.cuh", "fsb_aux.cuh", "../../include/fsb.h"]

# -fmad=false: no implicit FMA contraction -- every fused operation in the kernels is an explicit
# fma(), which is what makes results bit-identical to the CPU oracle.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        extra = os.environ.get("FSB_NVCC_EXTRA", "").split()  # e.g. -DFSB_PHASE_TIMING for scripts/phase_cycles.py
        # one object per translation unit, compiled in parallel (each takes minutes), then one link
        flags = [f for f in NVCC_FLAGS if f != "-shared"]
        objs, procs = [], []
        for src in SOURCES:
            obj = os.path.join(CSRC, src.replace(".cu", ".o"))
            objs.append(obj)
            procs.append(subprocess.Popen([nvcc] + flags + extra + ["-c", "-o", obj, os.path.join(CSRC, src)],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate()[0] for p in procs]
        ok = all(p.returncode == 0 for p in procs)
        # link to a temporary name and move it over SO only on success: a failed build must never leave (or
        # silently keep) a stale library that tests and benches would then load
        tmp = SO + ".tmp"
        cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-ldl"]
        res = subprocess.run(cmd, capture_output=True, text=True) if ok else subprocess.CompletedProcess(cmd, 1, "", "compile failed")
        log = " ".join(cmd) + "\n" + "\n".join(outs) + (res.stdout or "") + (res.stderr or "")
        for obj in objs:
            if os.path.exists(obj):
                os.remove(obj)
        with open(os.path.join(CSRC, "build.log"), "w") as f:
            f.write(log)
        if res.returncode != 0 or not os.path.exists(tmp):
            if os.path.exists(tmp):
                os.remove(tmp)
            tail = "\n".join(l for l in log.splitlines() if "error" in l.lower() or "fatal" in l.lower())[-4000:] or log[-4000:]
            raise RuntimeError(f"building {SO} failed (full log: {os.path.join(CSRC, 'build.log')}):\n{tail}")
        os.replace(tmp, SO)
        if verbose:
            print(log)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
