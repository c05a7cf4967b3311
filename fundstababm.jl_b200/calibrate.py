"""calibrate -- the batch driver that realises the sketch in src/calibrate.jl:1-30:

    Sampler -> for step in budget -> for rep in reps -> Model -> Evaluation -> Result averaging

as ONE device-resident batch: every (parameter point, replication) pair is a replication of the
engine, sharded over the ranks of ``torch.distributed`` (if initialised) by contiguous blocks of
the flattened index; the per-point integer ensemble statistics are summed with a single NCCL
allreduce inside the library (fsb_stats).  The surrogate / selector stages of the sketch are
host-side model fitting and stay out of scope.
"""
from __future__ import annotations

import itertools

from . import params as Params
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


def calibrate(points, reps, *, seed=0, fundselector="probabilistic", device=0, t_last=None, steps_per_launch=1,
              engine_kwargs=None):
    """Runs ``reps`` replications of every point and returns one statistics dict per point
    (identical on every rank)."""
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    total = len(points) * reps
    lo, hi = shard(total, rank, world)
    if hi <= lo:
        raise ValueError("fewer (point, replication) pairs than ranks")
    kw = dict(engine_kwargs or {})
    with Engine(points, hi - lo, rep_offset=lo, reps_per_point=reps, device=device,
                steps_per_launch=steps_per_launch, **kw) as eng:
        if world > 1:
            eng.comm_init_rank(world, rank, share_unique_id(dist, device))
        eng.init_device(seed)
        eng.run(t_last=t_last, selector=fundselector)
        return eng.stats()
