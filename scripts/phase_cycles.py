"""Per-phase cycle accounting of the step kernel (needs a build with FSB_NVCC_EXTRA=-DFSB_PHASE_TIMING).
Prints the share of CTA time (thread 0's clock) spent before each mark of fsb_step.cuh."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

NAMES = {0: "queue/idle between steps", 1: "P0 shocks+maps+reviewer draw (to B1)", 2: "review pass (to B2)",
         3: "no-divestment day: open pass + end", 4: "compaction + stake sequence (to B4)", 5: "sale orders + impact (to B5)",
         6: "cash-outs + flow hist (to B6)", 7: "respawn block", 8: "cash list + horizons (to B7)",
         9: "column staging wait (to B8)", 10: "full pass (to B9)", 11: "rank sort (to its barrier)",
         12: "fund choice (to its barrier)", 13: "injections (to B13)", 14: "stakes + buy orders + impact (to B14)",
         15: "final revalue + price store (to end)", 16: "  (4a) compaction + counters, warp 0",
         17: "  (4b) column prefetch issue", 18: "  (4c) stake sequence, warp 0", 4: "  (4d) wait at B4",
         19: "  (1a) prefetch + market + stock move", 20: "  (1b) maps + stakes to shared", 1: "  (1c) reviewer draw + wait at B1",
         23: "  (11a) rank sort of warp 0 columns", 11: "  (11b) choice scans of warp 0", 24: "  (14a) stakes loop", 25: "  (14b) buy orders + impact", 14: "  (14c) wait at B14"}
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=4096)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--t0", type=int, default=300)
ap.add_argument("--blocks-per-sm", type=int, default=0)
a = ap.parse_args()
fsb = load_package()
p = fsb.Params.default()
with fsb.Engine(p, a.reps, blocks_per_sm=a.blocks_per_sm) as eng:
    eng.init_device(7)
    eng.run_async(102, a.t0 - 1)
    eng.sync()
    out = np.zeros(32, dtype=np.uint64)
    lib = fsb._lib.lib()
    lib.fsb_debug_phase_cycles(eng.handle, out.ctypes.data, 1)
    eng.run_async(a.t0, a.t0 + a.steps - 1)
    eng.sync()
    lib.fsb_debug_phase_cycles(eng.handle, out.ctypes.data, 1)
    ms, _, n = eng.last_run_timing()
tot = float(out[:26].sum())
print(f"reps={a.reps} steps={a.steps} ms={ms:.3f} fund_steps_per_s={a.reps * p.bigk * a.steps / (ms * 1e-3):.4e}")
per = a.reps * a.steps
for i in range(26):
    if out[i]:
        print(f"{100 * out[i] / tot:5.1f}%  {out[i] / per:9.0f} cycles/rep-step  [{i:2d}] {NAMES.get(i, '')}")
print(f"total {tot / per:.0f} cycles per replication-step (thread 0 of the CTA)")
