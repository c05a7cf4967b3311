#!/usr/bin/env python
"""bench.py -- fund-steps/sec of the B200 engine on the FundStabABM simulation step.

    python bench.py --gpus N --steps K --warmup W            (N=1; N>1 under torch.distributed.run)
    python bench.py --impl reference ...                     (CPU oracle port on the host cores)

A "step" is one simulated day for the whole batch of replications (five kernel launches per GPU and
replication batch: events-1, holdings pass, events-2).  Workload at N GPUs: the default Params shape (K=M=250, N=1000, W=100;
BASELINE.json configs[2] at N=1) with --reps-per-gpu replications resident in HBM on every GPU
(weak scaling: replications are independent, there is no data-path collective), Philox shocks.
`value` is fund-steps/s with state resident in HBM; `e2e` is the same metric through the C-ABI
in equivalence mode (configs[1]: 10^4 replications, every step's shocks copied from pinned host
buffers and the NAVs copied back inside the timed region).
Prints ONE JSON line on rank 0.
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

METRIC = "fund_steps_per_sec"
UNIT = "fund-steps/s"
SEED = 20261019


def b_alg(K, M, N):
    """Algorithmic bytes per fund-step of the WHOLE step: 8M + 40M/K + 8 + 16N/K."""
    return 8.0 * M + 40.0 * M / K + 8.0 + 16.0 * N / K


def b_alg_pass(K, M, N):
    """Algorithmic bytes per fund-step of the dominant kernel alone (fsb_pass_kernel): the
    holdings row (8M), the two price columns of the day (16M/K) and the two NAVs written (16)."""
    return 8.0 * M + 16.0 * M / K + 16.0


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU side: the oracle port on the host cores (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------
def cpu_run(params, n_threads, reps_per_thread, t_first, t_last, flags=0):
    """Every thread owns `reps_per_thread` replications and advances each over [t_first,t_last]
    (ctypes releases the GIL, so the threads run on separate cores).  Returns seconds."""
    from oracle import oracle as orc
    worlds = []
    for i in range(n_threads * reps_per_thread):
        w = orc.World(params, seed=SEED, rep=i)
        w.initialise()
        w.modelrun(t_first, t_first - 1)  # builds the log outside the timed region
        worlds.append(w)

    def work(tid):
        for j in range(reps_per_thread):
            worlds[tid * reps_per_thread + j].modelrun(t_first, t_last, "probabilistic", flags)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    t0 = time.perf_counter()
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    dt = time.perf_counter() - t0
    for w in worlds:
        w.close()
    return dt


def cpu_baseline(params, budget_s=12.0):
    """The oracle port on the host cores, bounded samples of the same workload:
      value           lazy-NAV oracle, one replication per host thread (all cores)
      lazy_1core      the same on one core
      faithful_1core  ORC_FAITHFUL_HISTORY on one core: fundrevalue! rewrites the whole K x T value history at
                      functions.jl:460, :479 and after every buy order (:260) -- the work the reference itself does"""
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    t_first = params.perfwindow[-1] + 2
    dt = cpu_run(params, 1, 1, t_first, t_first + 39)          # calibrate: one replication, 40 days, one thread
    per_day = dt / 40.0
    lazy1 = params.bigk / per_day
    days = int(max(40, min(params.bigt - t_first + 1, budget_s / per_day)))
    dt = cpu_run(params, cores, 1, t_first, t_first + days - 1)
    value = cores * params.bigk * days / dt
    fdays = 6
    dtf = cpu_run(params, 1, 1, t_first, t_first + fdays - 1, flags=orc.FAITHFUL_HISTORY)
    if dtf < 3.0:  # (a fast host: lengthen the sample)
        fdays = int(min(60, max(fdays, fdays * 6.0 / max(dtf, 1e-3))))
        dtf = cpu_run(params, 1, 1, t_first, t_first + fdays - 1, flags=orc.FAITHFUL_HISTORY)
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} replications (1 per thread) x {days} days, default Params shape, lazy-NAV oracle "
                      f"(oracle/fsb_oracle.c)",
            "lazy_1core": {"value": lazy1, "unit": UNIT, "sample": "1 replication x 40 days, 1 thread"},
            "faithful_1core": {"value": params.bigk * fdays / dtf, "unit": UNIT,
                               "sample": f"1 replication x {fdays} days, 1 thread, ORC_FAITHFUL_HISTORY: the full-history "
                                         "fundrevalue! of functions.jl:266-271 at :460, :479 and :260"}}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  Julia is not in this image, so this is the
    oracle port (kind="port"), one replication per host thread on PERSISTENT threads (a barrier per bench step), each
    step a bounded sample: `days` simulated days of every replication."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fsb = load()
    params = fsb.Params.default()
    cores = os.cpu_count() or 1
    t_first = params.perfwindow[-1] + 2
    total_steps = args.warmup + args.steps
    days = max(1, min(12, (params.bigt - t_first + 1) // max(1, total_steps)))  # simulated days per bench step and replication
    from oracle import oracle as orc
    orc.build()
    worlds = []
    for i in range(cores):
        w = orc.World(params, seed=SEED, rep=i)
        w.initialise()
        w.modelrun(t_first, t_first - 1)
        worlds.append(w)
    gate = threading.Barrier(cores + 1)

    def worker(w):
        t = t_first
        for _ in range(total_steps):
            gate.wait()
            w.modelrun(t, t + days - 1, "probabilistic", 0)
            t += days
            gate.wait()

    threads = [threading.Thread(target=worker, args=(w,), daemon=True) for w in worlds]
    for th in threads:
        th.start()

    def step():
        gate.wait()  # release the workers
        gate.wait()  # ... and wait for all of them

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    for th in threads:
        th.join()
    value = cores * params.bigk * days * args.steps / dt
    sample = f"{cores} replications (1 per persistent host thread) x {days} days per step, lazy-NAV oracle port"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(params, cores, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Julia is not installed in this image; the reference's own runmodel.jl cannot be timed here "
                "(julia/time_reference.jl is the script for a box that has it)"}))


def workload_config(params, reps, where):
    return {"workload": "FundStabABM default Params shape (BASELINE.json configs[2] at 1 GPU; configs[1] for e2e)",
            "K_funds": params.bigk, "M_stocks": params.bigm, "N_investors": params.bign,
            "W_perfwindow": params.perfwindow[-1], "selector": "probabilistic", "replications_per_gpu": reps,
            "device": where, "l2": "state of all replications (>= 9 GB) far exceeds the 126 MB L2; no flush needed"}


def load():
    from __graft_entry__ import load_package
    return load_package()


# ---------------------------------------------------------------------------------------------
def timed_days(eng, t0, warmup, steps, barrier, dist=None, world=1, clocks_on=None):
    """W untimed + EXACTLY K timed simulated days on one engine; device time (CUDA events inside the library, max over
    ranks), the stats collective included in the wall clock.  Returns a dict of raw measurements."""
    import torch
    if warmup:
        eng.run_async(t0, t0 + warmup - 1)
        eng.sync()
    sampler = None
    if clocks_on is not None:
        sampler = ClockSampler(clocks_on)
        sampler.start()
    barrier()
    w0 = time.perf_counter()
    # a replication the model itself stops (Q6: functions.jl:282 would throw) is frozen and counted in
    # checks.error_replications, it does not abort the batch
    eng.run_async(t0 + warmup, t0 + warmup + steps - 1)
    eng.sync()
    stats = eng.stats_raw()                                       # includes the single NCCL allreduce
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - w0)
    clocks = sampler.stop() if sampler else None
    dev_ms, pass_ms, launches = eng.last_run_timing()             # CUDA events: whole loop / summed pass-kernel launches
    if world > 1:
        tmax = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(tmax[0]), float(tmax[1])
    return dict(dev_ms=dev_ms, wall_ms=wall_ms, pass_ms=pass_ms, launches=int(launches), n_pass=max(1, eng.last_pass_launches()),
                stats=stats, clocks=clocks)


def roofline(params, R, steps, m, world, value):
    K, M, N = params.bigk, params.bigm, params.bign
    peak, peak_src = measured_peak()
    per_launch_ms = m["pass_ms"] / m["n_pass"]
    reps_per_launch = R * steps / m["n_pass"]
    achieved = reps_per_launch * K * b_alg_pass(K, M, N) / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic_per_launch.json")
    if os.path.exists(tr_path) and (K, M) in ((250, 250), (2000, 5000)):
        try:
            with open(tr_path) as f:
                tj = json.load(f)
            traffic = (tj if K == 250 else tj["config4"])["dram_bytes_per_replication_step"] * reps_per_launch
        except Exception:
            traffic = None
    whole = value / world * b_alg(K, M, N) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_fund_step": b_alg_pass(K, M, N),
            "kernel": "fsb_pass_kernel", "avg_launch_ms": per_launch_ms, "launches_timed": m["n_pass"],
            "share_of_step_time": m["pass_ms"] / m["dev_ms"],
            "whole_step": {"algorithmic_bytes_per_fund_step": b_alg(K, M, N), "achieved": whole, "frac": whole / peak,
                           "what": "per GPU: all kernels of a step (draws, events-1, pass, choice, events-2) against SURVEY 8d's bytes"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5],
                    help="BASELINE.json configs[]: 3 = default shape, 10^5 replications (the bench line the driver runs); "
                         "4 = 2000 funds x 5000 assets; 5 = the calibrate sweep, 48 points x 10^3 replications")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --reps-per-gpu replications on every GPU; strong: --reps-per-gpu replications in total")
    ap.add_argument("--reps-per-gpu", type=int, default=100_000)
    ap.add_argument("--e2e-reps", type=int, default=10_000)
    ap.add_argument("--full-horizon-reps", type=int, default=4096, help="N=1: replications of the full 1249-day run (0 = skip)")
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--t0", type=int, default=0, help="first timed-region day minus warm-up (default: perfwindow[end] + 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the strong-scaling measurement (10^5 replications in total)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fsb = load()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == 5:
        return bench_config5(args, fsb, dist if world > 1 else None, rank, world, local, barrier)
    params = fsb.Params.default() if args.config == 3 else \
        fsb.Params.default(bigk=2000, bigm=5000, bign=8000, bigt=1101, portfsizerange=(2500, 5000))
    K, M, N = params.bigk, params.bigm, params.bign
    if args.config == 4 and args.reps_per_gpu == 100_000:
        args.reps_per_gpu = 1000  # BASELINE.json configs[3]: 10^3 replications (86 MB each; the pass kernel's work units are (replication, part, chunk))
    total = args.reps_per_gpu * (world if args.scaling == "weak" else 1)
    lo, hi = (rank * total) // world, ((rank + 1) * total) // world
    R = hi - lo
    t_first = params.perfwindow[-1] + 2
    total_days = args.warmup + args.steps
    assert t_first + total_days - 1 <= params.bigt

    def engine(nreps, offset, **kw):
        e = fsb.Engine(params, nreps, rep_offset=offset, device=local, blocks_per_sm=args.blocks_per_sm, **kw)
        if world > 1:
            from fundstababm_jl_b200.calibrate import share_unique_id
            e.comm_init_rank(world, rank, share_unique_id(dist, local))
        e.init_device(SEED)
        e.stats_raw()  # warms NCCL up (the one collective) outside the timed region
        return e

    eng = engine(R, lo)
    m = timed_days(eng, t_first, args.warmup, args.steps, barrier, dist, world, clocks_on=local)
    status_bad = int((eng.rep_status() != 0).sum())
    fund_steps = float(total) * K * args.steps
    value = fund_steps / (m["dev_ms"] * 1e-3)
    checks = {"error_replications": status_bad, "replication_steps_counted": int(m["stats"][0][8])}
    if world > 1:
        # the library's own ncclAllReduce against an independent sum of the per-rank statistics (torch.distributed)
        loc = torch.from_numpy(eng.stats_raw(local=True).astype(np.int64)).cuda(local)
        dist.all_reduce(loc)
        checks["allreduce_equals_sum_of_rank_statistics"] = bool((loc.cpu().numpy().astype(np.uint64) == m["stats"]).all())
        checks["replications_in_statistics"] = int(m["stats"][0][0])
    roof = roofline(params, R, args.steps, m, world, value)
    eng.close()

    # ---- strong scaling (N > 1): BASELINE.json configs[2] literally -- 10^5 replications in total, sharded -----------------
    strong = None
    if world > 1 and args.scaling == "weak" and args.config == 3 and not args.no_strong:
        tot_s = args.reps_per_gpu
        lo_s, hi_s = (rank * tot_s) // world, ((rank + 1) * tot_s) // world
        eng_s = engine(hi_s - lo_s, lo_s)
        ms = timed_days(eng_s, t_first, args.warmup, args.steps, barrier, dist, world)
        strong = {"value": float(tot_s) * K * args.steps / (ms["dev_ms"] * 1e-3), "unit": UNIT, "replications_total": tot_s,
                  "replications_per_gpu": hi_s - lo_s, "ms_per_step": ms["dev_ms"] / args.steps,
                  "what": "the same days with 10^5 replications in TOTAL sharded over the GPUs (strong scaling; compare with the "
                          "N=1 line's value)"}
        eng_s.close()

    # ---- e2e: equivalence mode through the C ABI with host buffers (configs[1]) -----------------
    e2e = None
    if not args.no_e2e and args.config == 3:
        Re = args.e2e_reps
        eng2 = fsb.Engine(params, Re, rep_offset=rank * Re, device=local)
        eng2.init_device(SEED)
        rng = np.random.default_rng(1234 + rank)
        pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory().numpy()
        nav = pin(Re, K)
        nsteps = args.steps
        # the producer side (a Julia recorder in equivalence mode) owns pinned buffers holding each step's shocks;
        # four rotating sets, filled before the timed region
        tapes = []
        for _ in range(min(4, nsteps)):
            a, b, c = pin(Re), pin(Re, M), pin(Re, N)
            a[...] = rng.standard_normal(Re)
            b[...] = rng.standard_normal((Re, M))
            c[...] = rng.random((Re, N))
            tapes.append((a, b, c))

        def e2e_step(t, i):
            a, b, c = tapes[i % len(tapes)]
            eng2.run_step_hosttape(t, a, b, c, nav)

        for i in range(args.warmup):
            e2e_step(t_first + i, i)
        barrier()
        t0 = time.perf_counter()
        for i in range(nsteps):
            e2e_step(t_first + args.warmup + i, i)
        barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            tm = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            e2e_s = float(tm[0])
        assert np.isfinite(nav).all() and not eng2.rep_status().any()
        e2e = {"value": world * Re * K * nsteps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(8 * Re * (1 + M + N)), "d2h_bytes_per_step": int(8 * Re * K),
               "replications_per_gpu": Re, "ms_per_step": 1e3 * e2e_s / nsteps,
               "what": "fsb_run_step_hosttape: per step H2D of the step's shocks from pinned host buffers, "
                       "five kernel launches per replication batch (8 batches, transfers overlap kernels), D2H of all NAVs; wall clock; "
                       "the exogenous shocks come from the host tape, event draws (anchors, picks, choices) from device Philox"}
        eng2.close()

    # ---- the whole horizon once (N = 1): 1249 simulated days, with the share of replications that end alive -----------
    full = None
    if world == 1 and args.config == 3 and args.full_horizon_reps > 0:
        Rf = args.full_horizon_reps
        with fsb.Engine(params, Rf, device=local) as ef:
            ef.init_device(SEED)
            barrier()
            ef.run_async()
            ef.sync()
            ms_f, _, _ = ef.last_run_timing()
            st = ef.stats()[0]
        days_f = params.bigt - t_first + 1
        full = {"value": Rf * K * days_f / (ms_f * 1e-3), "unit": UNIT, "replications": Rf, "days": days_f,
                "nan_replications": st["nan_replications"], "frozen_replications_q6": st["error_replications"],
                "live_fraction": 1.0 - (st["nan_replications"] + st["error_replications"]) / Rf,
                "floor_hits_per_replication": st["floor_hits"] / Rf, "liquidations_per_replication": st["liquidations"] / Rf,
                "what": "fsb_run over days 102..1350 (device time).  NaN replications are the reference's own Q12 (reinvestment "
                        "into a zero-value fund, functions.jl:367-368); their fund-steps are counted like live ones"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config == 3:
        from oracle import oracle as orc
        orc.build()
        cpu = cpu_baseline(params)
    if rank == 0:
        cfg = workload_config(params, R, "B200")
        if args.config == 4:
            cfg["workload"] = "BASELINE.json configs[3]: 2000 funds x 5000 assets x 8000 investors, portfolios of 2500..5000 stocks"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["dev_ms"] / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(cfg, global_replications=total, first_timed_day=t_first + args.warmup,
                           timing="CUDA events on the engine stream around the K simulated days, max over ranks; "
                                  f"wall clock incl. stats allreduce {m['wall_ms']:.2f} ms"),
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": m["launches"], "clocks": m["clocks"],
            "checks": checks,
        }
        if strong:
            line["strong"] = strong
        if full:
            line["full_horizon"] = full
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_config5(args, fsb, dist, rank, world, local, barrier):
    """BASELINE.json configs[4]: the calibrate sweep -- thresholdstd x reviewprobability x impact scale = 48 points x 10^3
    replications of the default shape over the full horizon, sharded over the GPUs, one allreduce of the per-point
    statistics.  A "step" here is the whole sweep; --steps / --warmup count sweeps."""
    import hashlib
    import numpy as np
    from fundstababm_jl_b200.calibrate import calibrate, grid
    cal_mod = sys.modules["fundstababm_jl_b200.calibrate"]  # (the package exports the function under the module's name)
    base = fsb.Params.default()
    points = grid(base, thresholdstd=(0.01, 0.025, 0.05, 0.1), reviewprobability=(1 / 126, 1 / 63, 1 / 21),
                  impactscale=(0.5, 1.0, 2.0, 4.0))
    reps = 1000 if args.reps_per_gpu == 100_000 else args.reps_per_gpu
    days = base.bigt - base.perfwindow[-1] - 1
    times, stats = [], None
    for i in range(max(0, args.warmup) + max(1, args.steps)):
        barrier()
        t0 = time.perf_counter()
        stats = calibrate(points, reps, seed=SEED, device=local, engine_kwargs=dict(log_capacity=base.bigk + 4096))
        barrier()
        times.append(time.perf_counter() - t0)
    dt = float(np.mean(times[max(0, args.warmup):]))
    if rank == 0:
        h = hashlib.sha256()
        for st in stats:
            for k in sorted(st):
                h.update(np.asarray(st[k]).tobytes())
        fs = len(points) * reps * base.bigk * days
        print(json.dumps({
            "metric": METRIC, "value": fs / dt, "unit": UNIT, "n_gpus": world, "steps": max(1, args.steps), "warmup": max(0, args.warmup),
            "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[4]: calibrate sweep, thresholdstd x reviewprobability x impact scale",
                       "points": len(points), "replications_per_point": reps, "days": days, "K_funds": base.bigk,
                       "timing": "wall clock of calibrate(): engine creation, device initialisation, 1249 days, statistics allreduce",
                       "rank0_split_of_the_last_sweep_s": {k: round(v, 3) for k, v in cal_mod.last_timing.items()}},
            "checks": {"statistics_sha256": h.hexdigest(), "replications": int(sum(st["replications"] for st in stats)),
                       "error_replications_q6": int(sum(st["error_replications"] for st in stats)),
                       "nan_replications": int(sum(st["nan_replications"] for st in stats)),
                       "liquidations_by_point": [int(st["liquidations"]) for st in stats]},
            "gpu_launches": 5 * days * max(1, args.steps)}))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
