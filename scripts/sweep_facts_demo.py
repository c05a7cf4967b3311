"""BASELINE.json configs[4] in small (a calibration sweep over the flow-performance and price-impact parameters)
with the "Evaluation" stage of src/calibrate.jl:14-16 on the device: per sweep point the integer ensemble statistics
and the ensemble means of Func.calc_stylisedfacts.  One GPU, default Params shape, full 1249-day runs.
Usage: python scripts/sweep_facts_demo.py [--reps 128] > gpurun_out/sweep_facts_<tag>.json"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=128)
a = ap.parse_args()
fsb = load_package()
pts = fsb.grid(thresholdstd=(0.025, 0.1), reviewprobability=(1 / 126, 1 / 21), impactscale=(0.5, 4.0))
t0 = time.perf_counter()
stats = fsb.calibrate(pts, a.reps, seed=0x5EED, evaluation="stylisedfacts")
wall = time.perf_counter() - t0
rows = []
for p, st in zip(pts, stats):
    sf = st["stylisedfacts"]
    rows.append(dict(thresholdstd=p.thresholdstd, reviewprobability=p.reviewprobability,
                     impact_max=float(max(p.impactrange.values() if hasattr(p.impactrange, "values") else p.impactrange)),
                     replications=st["replications"], liquidations=st["liquidations"], divestments=st["divestments"],
                     error_replications=st["error_replications"], nan_replications=st["nan_replications"],
                     mean_kurtosis=sf["stockkurtoses"], kurtosis_n=sf["stockkurtoses_n"], mean_lossgain=sf["lossgainratios"],
                     mean_volume_corr=sf["corrs"], stock_pacf_lag1=float(sf["stockpacf"][0]),
                     stock_volacluster_lag1=float(sf["stockvolacluster"][0])))
print(json.dumps(dict(what="calibrate(evaluation='stylisedfacts'), default shape, 8 points", replications_per_point=a.reps,
                      wall_s=wall, fund_steps=len(pts) * a.reps * 250 * 1249, points=rows), indent=1))
