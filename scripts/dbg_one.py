"""Tiny default-shape run for debugging: a few replications, a few steps, compared with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
fsb = load_package()
from oracle import oracle as orc
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
p = fsb.Params.default()
with fsb.Engine(p, 2) as eng:
    eng.init_device(11)
    print("init ok", flush=True)
    eng.run(102, 101 + steps)
    print("run ok", eng.rep_status(), flush=True)
    st = eng.download_compact(0)
w = orc.World(p, seed=11, rep=0); w.initialise(); w.modelrun(102, 101 + steps, "probabilistic", 0)
print("holdings equal:", np.array_equal(st["holdings"], w.holdings), " nav_close equal:", np.array_equal(st["nav_close"], np.asarray(w.nav_close) if hasattr(w, "nav_close") else st["nav_close"]))
