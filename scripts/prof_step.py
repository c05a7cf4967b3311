"""Small driver for ncu: a few launches of the per-step kernel at the default shape."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=4096)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--t0", type=int, default=102)
ap.add_argument("--blocks-per-sm", type=int, default=0)
ap.add_argument("--block-threads", type=int, default=0)
ap.add_argument("--digest", action="store_true", help="also print a digest of the first and last replication's state (kernel variants must agree)")
ap.add_argument("--config4", action="store_true", help="BASELINE.json configs[3]: 2000 funds x 5000 assets, N = 8000, dense portfolios")
a = ap.parse_args()
fsb = load_package()
p = fsb.Params.default()
if a.config4:
    p = fsb.Params.default(bigk=2000, bigm=5000, bign=8000, bigt=1101, portfsizerange=(2500, 5000))
with fsb.Engine(p, a.reps, blocks_per_sm=a.blocks_per_sm,
                block_threads=a.block_threads) as eng:
    eng.init_device(7)
    if a.t0 > 102:
        eng.run_async(102, a.t0 - 1)
        eng.sync()
    eng.run_async(a.t0, a.t0 + a.steps - 1)  # replications the model itself stops (Q6) stay frozen, not fatal
    eng.sync()
    ms, pass_ms, n = eng.last_run_timing()
    nerr = int((eng.rep_status() != 0).sum())
    print(f"reps={a.reps} steps={a.steps} launches={n} total_ms={ms:.3f} per_launch_ms={ms / n:.4f} "
          f"fund_steps_per_s={a.reps * p.bigk * a.steps / (ms * 1e-3):.4e} frozen_reps={nerr} pass_ms={pass_ms:.3f}")
    if a.digest:
        import hashlib
        h = hashlib.sha256()
        for r in (0, a.reps - 1):
            st = eng.download_compact(r)
            for k in sorted(st):
                h.update(st[k].tobytes())
        print("digest", h.hexdigest()[:16])
