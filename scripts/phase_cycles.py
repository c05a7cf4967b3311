"""Per-phase cycle accounting of the step kernel (needs a build with FSB_NVCC_EXTRA=-DFSB_PHASE_TIMING).
Prints the share of CTA time (thread 0's clock) spent before each mark of fsb_step.cuh."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

NAMES = {0: "queue/idle between steps", 19: "T0a prefetch + market + stock move", 20: "T0b maps + flags to shared",
         1: "T0c reviewer draw + wait at B1", 2: "T1 review pass (to B2)", 16: "T1b compaction + counters",
         3: "no-divestment day: open pass + end", 17: "T2a column prefetch issue", 18: "T2b stake sequence (own warp)",
         4: "T2c wait at B4", 5: "T2d sale orders + impact (to B5)", 6: "T2e cash-outs + flow hist (to B6)",
         7: "T3 respawn block", 8: "T4 cash list + horizons (to B7)", 9: "T4b column staging wait (to B8)",
         10: "T5 full pass (to B9)", 23: "T6a rank sort of own columns", 11: "T6b choice scans",
         12: "T6c wait for the other warps' choices", 13: "T7a injections (to B13)", 24: "T7b stakes loop",
         25: "T8 buy orders + impact", 26: "X set-up (queue, hand_i, hdr)", 27: "X re-pack marked rows", 28: "X stage columns + row order", 29: "X rows: fill + dense + packed loop", 30: "X rows: wait for the stage", 31: "X rows: epilogue", 14: "T8b wait at B14", 15: "T9 final revalue + price store (to end)"}
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=4096)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--t0", type=int, default=300)
ap.add_argument("--blocks-per-sm", type=int, default=0)
ap.add_argument("--config4", action="store_true", help="BASELINE configs[3] (needs -DFSB_BIG_TIMING)")
a = ap.parse_args()
fsb = load_package()
p = fsb.Params.default()
if a.config4:
    p = fsb.Params.default(bigk=2000, bigm=5000, bign=8000, bigt=1101, portfsizerange=(2500, 5000))
with fsb.Engine(p, a.reps, blocks_per_sm=a.blocks_per_sm) as eng:
    eng.init_device(7)
    eng.run_async(102, a.t0 - 1)
    eng.sync()
    out = np.zeros(32, dtype=np.uint64)
    lib = fsb._lib.lib()

    def read_marks():
        if a.config4:
            fn = C.CDLL(fsb._lib.so_path()).fsb_debug_big_phase_cycles
            fn.argtypes = [C.c_void_p, C.c_int]
            fn(out.ctypes.data, 1)
        else:
            lib.fsb_debug_phase_cycles(eng.handle, out.ctypes.data, 1)

    read_marks()
    eng.run_async(a.t0, a.t0 + a.steps - 1)
    eng.sync()
    read_marks()
    ms, _, n = eng.last_run_timing()
tot = float(out[:32].sum())
print(f"reps={a.reps} steps={a.steps} ms={ms:.3f} fund_steps_per_s={a.reps * p.bigk * a.steps / (ms * 1e-3):.4e}")
per = a.reps * a.steps
for i in range(32):
    if out[i]:
        print(f"{100 * out[i] / tot:5.1f}%  {out[i] / per:9.0f} cycles/rep-step  [{i:2d}] {NAMES.get(i, '')}")
print(f"total {tot / per:.0f} cycles per replication-step (thread 0 of the CTA)")
