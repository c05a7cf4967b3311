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
    cores = os.cpu_count() or 1
    t_first = params.perfwindow[-1] + 2
    # calibrate: one replication, 40 days, one thread
    dt = cpu_run(params, 1, 1, t_first, t_first + 39)
    per_day = dt / 40.0
    days = int(max(40, min(params.bigt - t_first + 1, budget_s / per_day)))
    dt = cpu_run(params, cores, 1, t_first, t_first + days - 1)
    value = cores * params.bigk * days / dt
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} replications (1 per thread) x {days} days, default Params shape, lazy-NAV oracle "
                      f"(oracle/fsb_oracle.c); single-thread rate {params.bigk / per_day:.3e} {UNIT}"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  Julia is not in this
    image, so this is the oracle port (kind="port"), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fsb = load()
    params = fsb.Params.default()
    cores = os.cpu_count() or 1
    t_first = params.perfwindow[-1] + 2
    days = 12  # simulated days per bench step and replication
    from oracle import oracle as orc
    orc.build()
    worlds = []
    for i in range(cores):
        w = orc.World(params, seed=SEED, rep=i)
        w.initialise()
        w.modelrun(t_first, t_first - 1)
        worlds.append(w)
    t = t_first

    def step(t0):
        ths = [threading.Thread(target=lambda w=w: w.modelrun(t0, t0 + days - 1, "probabilistic", 0)) for w in worlds]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    for _ in range(args.warmup):
        step(t)
        t += days
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(t)
        t += days
    dt = time.perf_counter() - t0
    value = cores * params.bigk * days * args.steps / dt
    sample = f"{cores} replications (1 per host thread) x {days} days per step, lazy-NAV oracle port"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(params, cores, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Julia is not installed in this image; the reference's own runmodel.jl cannot be timed here"}))


def workload_config(params, reps, where):
    return {"workload": "FundStabABM default Params shape (BASELINE.json configs[2] at 1 GPU; configs[1] for e2e)",
            "K_funds": params.bigk, "M_stocks": params.bigm, "N_investors": params.bign,
            "W_perfwindow": params.perfwindow[-1], "selector": "probabilistic", "replications_per_gpu": reps,
            "device": where, "l2": "state of all replications (>= 9 GB) far exceeds the 126 MB L2; no flush needed"}


def load():
    from __graft_entry__ import load_package
    return load_package()


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reps-per-gpu", type=int, default=100_000)
    ap.add_argument("--e2e-reps", type=int, default=10_000)
    ap.add_argument("--steps-per-launch", type=int, default=1)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    params = fsb.Params.default()
    K, M, N = params.bigk, params.bigm, params.bign
    R = args.reps_per_gpu
    t_first = params.perfwindow[-1] + 2
    spl = max(1, args.steps_per_launch)
    # one bench "step" = one simulated day for every replication
    total_days = args.warmup + args.steps
    assert t_first + total_days - 1 <= params.bigt

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = fsb.Engine(params, R, rep_offset=rank * R, device=local, steps_per_launch=spl,
                     blocks_per_sm=args.blocks_per_sm)
    if world > 1:
        from fundstababm_jl_b200.calibrate import share_unique_id
        eng.comm_init_rank(world, rank, share_unique_id(dist, local))
    eng.init_device(SEED)
    eng.stats_raw()  # warms NCCL up (the one collective) outside the timed region
    # ---- warm-up (W untimed steps) ------------------------------------------------------------
    if args.warmup:
        eng.run_async(t_first, t_first + args.warmup - 1)
        eng.sync()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    # EXACTLY K steps, device-timed inside; a replication the model itself stops (Q6: functions.jl:282
    # would throw) is frozen and counted in checks.error_replications, it does not abort the batch
    eng.run_async(t_first + args.warmup, t_first + total_days - 1)
    eng.sync()
    stats = eng.stats_raw()                                       # includes the single NCCL allreduce
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    dev_ms, pass_ms, launches = eng.last_run_timing()             # CUDA events: whole loop / summed pass-kernel launches
    status_bad = int((eng.rep_status() != 0).sum())
    if world > 1:
        tmax = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(tmax[0]), float(tmax[1])
    fund_steps = float(world) * R * K * args.steps
    value = fund_steps / (dev_ms * 1e-3)
    peak, peak_src = measured_peak()
    # the dominant kernel: the holdings pass, one launch per simulated day and replication batch, timed launch by
    # launch with CUDA events on the launching streams (fsb_last_run_timing)
    n_pass = max(1, eng.last_pass_launches())
    per_launch_ms = pass_ms / n_pass
    reps_per_launch = R * args.steps / n_pass
    alg_bytes_per_launch = reps_per_launch * K * b_alg_pass(K, M, N)
    achieved = alg_bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic_per_launch.json")
    if os.path.exists(tr_path):
        try:
            with open(tr_path) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_per_replication_step"] * reps_per_launch
        except Exception:
            traffic = None
    eng.close()

    # ---- e2e: equivalence mode through the C ABI with host buffers (configs[1]) -----------------
    e2e = None
    if not args.no_e2e:
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
                       "five kernel launches per replication batch (8 batches, transfers overlap kernels), D2H of all NAVs; wall clock"}
        eng2.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        orc.build()
        cpu = cpu_baseline(params)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(params, R, "B200"), steps_per_launch=spl, global_replications=world * R,
                           timing="CUDA events on the engine stream around the K simulated days, max over ranks; "
                                  f"wall clock incl. stats allreduce {wall_ms:.2f} ms"),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_fund_step": b_alg_pass(K, M, N),
                         "kernel": "fsb_pass_kernel", "avg_launch_ms": per_launch_ms, "launches_timed": n_pass,
                         "share_of_step_time": pass_ms / dev_ms,
                         "whole_step": {"algorithmic_bytes_per_fund_step": b_alg(K, M, N),
                                        "achieved": value / world * b_alg(K, M, N) / 1e9, "frac": value / world * b_alg(K, M, N) / 1e9 / peak,
                                        "what": "per GPU: all kernels of a step (draws, events-1, pass, choice, events-2) against SURVEY 8d's bytes"}},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "checks": {"error_replications": status_bad, "replication_steps_counted": int(stats[0][8])},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
