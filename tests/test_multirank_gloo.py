"""world_size-2 `gloo` tests (CPU) of the multi-rank host logic: how replications are sharded over
ranks, how the NCCL unique id travels from rank 0, and that the integer ensemble statistics are
invariant to the sharding (the property that makes the single allreduce order independent).
The data path itself has no collective: replications are independent."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from __graft_entry__ import load_package
        fsb = load_package()
        cal, eng_mod = fsb._calibrate, fsb.engine

        # 1. shards of the flattened (point, replication) index tile [0, total) exactly once
        total = 48 * 1000 + 7
        lo, hi = cal.shard(total, rank, world)
        spans = [None] * world
        dist.all_gather_object(spans, (lo, hi))
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))

        # 2. the 128-byte communicator id is created on rank 0 only and reaches every rank intact
        calls = []

        def fake_unique_id():
            calls.append(rank)
            return bytes((7 * i + 3) % 256 for i in range(128))

        cal.comm_unique_id = fake_unique_id
        uid = cal.share_unique_id(dist, 0)
        assert uid == bytes((7 * i + 3) % 256 for i in range(128))
        assert calls == ([0] if rank == 0 else [])

        # 3. what the library's allreduce does to the statistics words: a sum of per-shard integer
        #    words equals the words of the unsharded batch, whatever the split
        L = fsb._lib
        T = 30
        words = L.STATS_HEADER + (T + 1) + L.STATS_LIQ_BINS + L.STATS_DD_BINS + L.STATS_FLOW_BINS
        rng = np.random.default_rng(5)
        per_rep = rng.integers(0, 50, size=(total % 97 + 40, words)).astype(np.int64)  # one row per replication
        n = per_rep.shape[0]
        a, b = cal.shard(n, rank, world)
        mine = torch.from_numpy(per_rep[a:b].sum(axis=0))
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        assert np.array_equal(mine.numpy(), per_rep.sum(axis=0))
        st = eng_mod.decode_stats(mine.numpy().astype(np.uint64), T)
        assert st["replications"] == int(per_rep[:, L.ST_REPS].sum())

        # 3b. the float "result averaging" of calibrate(evaluation="stylisedfacts"): per-point sums and counts of
        #     the finite stylised facts, summed over ranks, equal the unsharded ones to rounding
        R, M, Lg, reps = 22, 6, 3, 11
        frng = np.random.default_rng(9)
        facts = dict(stockpacf=frng.standard_normal((R, M, Lg)), stockvolacluster=frng.standard_normal((R, M, Lg)),
                     marketpacf=frng.standard_normal((R, Lg)), marketvolacluster=frng.standard_normal((R, Lg)),
                     stockkurtoses=frng.standard_normal((R, M)), lossgainratios=frng.random((R, M)),
                     corrs=frng.random((R, M)))
        facts["corrs"][3, 2] = np.nan
        pts = np.arange(R) // reps
        a, b = cal.shard(R, rank, world)
        mom = torch.from_numpy(cal.facts_moments({k: v[a:b] for k, v in facts.items()}, pts[a:b], 2))
        dist.all_reduce(mom)
        np.testing.assert_allclose(mom.numpy(), cal.facts_moments(facts, pts, 2), rtol=1e-13)
        means = cal.facts_means(mom.numpy())
        assert means[0]["corrs_n"] == reps * M - 1 and means[1]["corrs_n"] == reps * M

        # 4. bench.py's timing reduction: max over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t[0]) == 10.0 + world - 1
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_reference_arm_runs_on_rank0_only(tmp_path):
    """bench.py --impl reference under torchrun: rank != 0 exits 0 without work or output."""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""
