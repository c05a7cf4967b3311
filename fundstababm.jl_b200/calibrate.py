"""calibrate -- the batch driver that realises the sketch in src/calibrate.jl:1-30:

    Sampler -> for step in budget -> for rep in reps -> Model -> Evaluation -> Result averaging

as ONE device-resident batch: every (parameter point, replication) pair is a replication of the
engine, sharded over the ranks of ``torch.distributed`` (if initialised) by contiguous blocks of
the flattened index; the per-point integer ensemble statistics are summed with a single NCCL
allreduce inside the library (fsb_stats).  The "Evaluation" stage can also be the stylised facts of
every replication (``evaluation="stylisedfacts"``: Func.calc_stylisedfacts on the device),
averaged per point.  The surrogate / selector stages of the sketch are host-side model fitting and stay
out of scope.
"""
from __future__ import annotations

import itertools

import numpy as np

from . import params as Params
from . import _lib as L
from .engine import Engine, comm_unique_id


def grid(base=None, *, thresholdstd=(0.05,), reviewprobability=(1 / 63,), impactscale=(1.0,), thresholdmean=(0.0,)):
    """Sampler: full factorial grid over the flow-performance parameters (src/params.jl:23-27)
    and a scale on the price-impact range (src/params.jl:16)."""
    base = Params.default() if base is None else base
    pts = []
    for ts, rp, sc, tm in itertools.product(thresholdstd, reviewprobability, impactscale, thresholdmean):
        imp = base.impactrange if sc == 1.0 else Params.ValueTable(tuple(v * sc for v in base.impactrange.values()))
        pts.append(base(thresholdstd=ts, reviewprobability=rp, thresholdmean=tm, impactrange=imp))
    return pts


def shard(total, rank, world):
    """Contiguous block [lo, hi) of the flattened (point, replication) index for this rank."""
    lo = (total * rank) // world
    hi = (total * (rank + 1)) // world
    return lo, hi


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def share_unique_id(dist, device):
    """Rank 0 creates the NCCL unique id; shipped to the others through torch.distributed."""
    import torch
    rank = dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device("cuda", device) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf = torch.tensor(list(comm_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().tolist())


FACT_SCALARS = ("stockkurtoses", "lossgainratios", "corrs")
FACT_LAGGED = ("stockpacf", "stockvolacluster", "marketpacf", "marketvolacluster")


def facts_moments(facts, point_of_rep, n_points):
    """Result averaging, step 1 (per rank): per point the sums and counts of the FINITE entries of every
    stylised fact over this rank's replications (and stocks) -- NaN trajectories (Q12) and stocks that never
    traded (NaN correlation) are left out, and counted.  Returns [n_points][n_fields][2] float64 where the
    fields are the three per-stock scalars followed by maxlag entries for each of the four PACF families."""
    maxlag = facts["marketpacf"].shape[-1]
    cols = []
    for k in FACT_SCALARS:                       # [rep][stock]
        cols.append(facts[k][:, :, None])
    for k in FACT_LAGGED[:2]:                    # [rep][stock][lag]
        cols.append(facts[k])
    for k in FACT_LAGGED[2:]:                    # [rep][lag]
        cols.append(facts[k][:, None, :])
    out = np.zeros((n_points, len(FACT_SCALARS) + 4 * maxlag, 2))
    pt = np.asarray(point_of_rep)
    f0 = 0
    for c in cols:
        ok = np.isfinite(c)
        v = np.where(ok, c, 0.0)
        for p in np.unique(pt):
            m = pt == p
            out[p, f0:f0 + c.shape[2], 0] += v[m].sum(axis=(0, 1))
            out[p, f0:f0 + c.shape[2], 1] += ok[m].sum(axis=(0, 1))
        f0 += c.shape[2]
    return out


def facts_means(moments):
    """Result averaging, step 2 (after the sum over ranks): one dict per point, keyed like calc_stylisedfacts."""
    maxlag = (moments.shape[1] - len(FACT_SCALARS)) // 4
    res = []
    for m in moments:
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = m[:, 0] / m[:, 1]
        d = {k: float(mean[i]) for i, k in enumerate(FACT_SCALARS)}
        d.update({k + "_n": int(m[i, 1]) for i, k in enumerate(FACT_SCALARS)})
        for j, k in enumerate(FACT_LAGGED):
            a = len(FACT_SCALARS) + j * maxlag
            d[k] = mean[a:a + maxlag].copy()
        res.append(d)
    return res


def calibrate_single_process(points, reps, devices, *, seed=0, fundselector="probabilistic", t_last=None, engine_kwargs=None):
    """The sweep driven by ONE process over several devices -- the form a Julia caller uses (src/runmodel.jl is one task;
    julia/FundStabB200.jl `calibrate`): one engine per device on a contiguous shard of the flattened (point, replication)
    index, the days enqueued asynchronously on every device, one grouped allreduce of the statistics (fsb_comm_init_all +
    fsb_stats_all).  Same statistics as calibrate() under torch.distributed with len(devices) ranks."""
    from .engine import comm_init_all, stats_all, decode_stats
    total, world = len(points) * reps, len(devices)
    if total < world:
        raise ValueError(f"{total} (point, replication) pairs cannot be sharded over {world} devices")
    engines = []
    try:
        for g, dev in enumerate(devices):
            lo, hi = shard(total, g, world)
            engines.append(Engine(points, hi - lo, rep_offset=lo, reps_per_point=reps, device=dev, **dict(engine_kwargs or {})))
        if world > 1:
            comm_init_all(engines)
        for e in engines:
            e.init_device(seed)
        for e in engines:
            e.run_async(t_last=t_last, selector=fundselector)
        for e in engines:
            e.sync()
        for e in engines:  # capacity / tape flags are errors, the model's own Q6 stops are counted (see calibrate)
            bad = e.rep_status() & ~(L.REP_Q6_NO_ANCHOR | L.REP_TRACE_OVERFLOW)
            if bad.any():
                raise L.FsbError(L.ERR_REPLICATION, f"device {e.device}: {int((bad != 0).sum())} replications stopped on engine capacity / tape flags")
        raw = stats_all(engines)
        return [decode_stats(row, engines[0].T) for row in raw]
    finally:
        for e in engines:
            e.close()


last_timing = {}  # wall-clock split of the last plain calibrate() call on this rank (bench.py --config 5 reports it)


def history_bytes_per_replication(p):
    """Device bytes a replication needs on top of the ring state when the engine keeps the price and volume histories
    (what the stylised-facts kernels read): 2 x T x Mp doubles."""
    mp = (p.bigm + 127) // 128 * 128
    return 2 * 8 * p.bigt * mp


def calibrate(points, reps, *, seed=0, fundselector="probabilistic", device=0, t_last=None,
              engine_kwargs=None, evaluation=None, maxlag=10, returnspctile=5, history_batch=None,
              history_budget_bytes=32 << 30, library_collective=False):
    """Runs ``reps`` replications of every point and returns one statistics dict per point
    (identical on every rank).

    ``evaluation="stylisedfacts"`` adds the per-point means (over replications and stocks) of Func.calc_stylisedfacts
    under ``"stylisedfacts"`` (float sums over ranks: equal across world sizes to rounding, unlike the integer
    statistics).  The facts need every return of a series three times over (their mean, then the percentile of the
    demeaned |returns|, then the counts against it: functions.jl:517-586), so they are computed by the history-based
    device kernels (fsb_stylised_facts) -- but the histories (5.4 MB per replication at the default shape) never have to
    exist for the whole ensemble at once: the rank's replications run in HISTORY BATCHES of ``history_batch``
    replications (default: what fits ``history_budget_bytes`` of device memory), each a history-keeping engine that is
    simulated, evaluated and released, while the integer statistics accumulate.  Replications are independent and the
    Philox counters use global replication ids, so the results do not depend on the batch size: a 10^5-replication sweep
    that only fits as a ring engine gets the same facts as ``keep_history`` would give.

    The one collective of a sweep -- the sum of the per-point integer statistics over the ranks -- goes through the
    process group the caller already has (torch.distributed, NCCL over NVLink on GPUs); ``library_collective=True`` uses the
    library's own communicator instead (fsb_comm_init_rank + fsb_stats, the path bench.py and the Julia glue exercise),
    which costs a communicator creation and its first-use connection set-up per call (about 1.5 s at 2 ranks)."""
    if evaluation not in (None, "stylisedfacts"):
        raise ValueError(f"unknown evaluation {evaluation!r}")
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    total = len(points) * reps
    if total < world:  # (decided identically on every rank, before any collective: no rank is left waiting in NCCL)
        raise ValueError(f"{total} (point, replication) pairs cannot be sharded over {world} ranks")
    lo, hi = shard(total, rank, world)
    kw = dict(engine_kwargs or {})

    def check_flags(eng):
        # Replications the MODEL stops (Q6: functions.jl:282 would throw) are frozen and counted in
        # error_replications.  Anything else is a capacity artefact of this engine (event / log rows) or a tape
        # problem: it would bias the point's statistics silently, so it is an error (raise the capacities).
        bad = eng.rep_status() & ~(L.REP_Q6_NO_ANCHOR | L.REP_TRACE_OVERFLOW)
        return int((bad != 0).sum()), (int(np.flatnonzero(bad)[0]), int(bad[np.flatnonzero(bad)[0]])) if bad.any() else None

    def allreduce(a):
        if world == 1:
            return a
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a))
        if dist.get_backend() == "nccl":
            t = t.cuda(device)
        dist.all_reduce(t)
        return t.cpu().numpy()

    def run(eng):
        try:
            eng.run(t_last=t_last, selector=fundselector)
        except L.FsbError as ex:
            if ex.code != L.ERR_REPLICATION:
                raise

    def fail_on(nbad, first):
        nbad = int(allreduce(np.array([nbad], dtype=np.int64))[0])
        if nbad:
            raise L.FsbError(L.ERR_REPLICATION, f"{nbad} replications stopped on engine capacity / tape flags"
                             + (f" (this rank: first local replication {first[0]}, flags 0x{first[1]:x})" if first else "")
                             + "; raise event_capacity / log_capacity in engine_kwargs")

    if not evaluation:
        import time as _time
        tm = {}
        t0 = _time.perf_counter()
        with Engine(points, hi - lo, rep_offset=lo, reps_per_point=reps, device=device, **kw) as eng:
            tm["create_s"] = _time.perf_counter() - t0
            t0 = _time.perf_counter()
            if world > 1 and library_collective:
                eng.comm_init_rank(world, rank, share_unique_id(dist, device))
            tm["comm_init_s"] = _time.perf_counter() - t0
            t0 = _time.perf_counter()
            eng.init_device(seed)
            tm["init_device_s"] = _time.perf_counter() - t0
            t0 = _time.perf_counter()
            run(eng)
            tm["run_s"] = _time.perf_counter() - t0
            t0 = _time.perf_counter()
            fail_on(*check_flags(eng))
            if library_collective or world == 1:
                out = eng.stats()   # the single NCCL allreduce, inside the library
            else:
                from .engine import decode_stats
                raw = allreduce(eng.stats_raw(local=True).astype(np.int64))   # the single allreduce, on the caller's process group
                out = [decode_stats(row.astype(np.uint64), points[0].bigt) for row in raw]
            tm["flags_and_stats_s"] = _time.perf_counter() - t0
        last_timing.clear()
        last_timing.update(tm)
        return out

    # ---- evaluation on the device, history batch by history batch ------------------------------------------------------
    kw["keep_history"] = True
    if history_batch is None:
        history_batch = max(1, int(history_budget_bytes // (history_bytes_per_replication(points[0]) + (1 << 20))))
    raw = mom = None
    nbad, first, T = 0, None, points[0].bigt
    for b_lo in range(lo, hi, history_batch):
        b_hi = min(hi, b_lo + history_batch)
        with Engine(points, b_hi - b_lo, rep_offset=b_lo, reps_per_point=reps, device=device, **kw) as eng:
            eng.init_device(seed)
            run(eng)
            nb, fb = check_flags(eng)
            nbad, first = nbad + nb, first or fb
            r = eng.stats_raw(local=True).astype(np.int64)
            facts = eng.stylised_facts(maxlag=maxlag, returnspctile=returnspctile)
            m = facts_moments(facts, (b_lo + np.arange(b_hi - b_lo)) // reps, len(points))
        raw = r if raw is None else raw + r
        mom = m if mom is None else mom + m
    fail_on(nbad, first)
    raw = allreduce(raw)   # (the same words fsb_stats would have summed with its own allreduce)
    mom = allreduce(mom)
    from .engine import decode_stats
    stats = [decode_stats(row.astype(np.uint64), T) for row in raw]
    for st, f in zip(stats, facts_means(mom)):
        st["stylisedfacts"] = f
    return stats
