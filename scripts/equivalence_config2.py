#!/usr/bin/env python
"""BASELINE.json configs[1] -- the default model shape in EQUIVALENCE MODE at scale: R replications x the full 1249-day
horizon through the C ABI, every replication compared with the oracle.

Two legs (VERDICT r1 "Next round" 2a):
  hosttape   fsb_run_step_hosttape: every day's exogenous shocks (market move, stock moves, reviewer uniforms) of ALL
             replications come from host buffers -- here filled with the oracle's Philox draws of those sites
             (oracle.fill_step_tape), the values a recorder around functions.jl:15, :32, :149 would capture;
  events     fsb_set_dense_tape + fsb_set_event_tape per replication: ALSO the respawn anchors / portfolio sizes / picks
             and the fund choices are forced from a recording (the oracle's trace).
What is compared: a 16-byte digest per replication over the final holdings (K x M), prices, investor positions / amounts /
stakes, NAVs, market value and the whole liquidation log (NaN payloads and the sign of zero canonicalised, everything else bit
for bit).  The oracle side can be computed anywhere (it needs no GPU): `--oracle-out` writes the digests, `--oracle-in`
reads them back, so the 10^4-replication run on the GPU box does not spend GPU time on 40 core-minutes of CPU work.

    python scripts/equivalence_config2.py --reps 10000 --oracle-out tests/golden/config2_oracle_digests.npz   (CPU only)
    python scripts/equivalence_config2.py --reps 10000 --oracle-in  tests/golden/config2_oracle_digests.npz --out profiles/equivalence_config2_r2.json
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED = 20261020
REP0 = 500_000  # global id of the first replication of this study


def canon(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return (np.where(np.isnan(a), np.nan, a) + 0.0).tobytes()


def digest(holdings, price_now, inv_fund, inv_amount, inv_stake, nav_close, market_now, log):
    """32 bytes: [0:16] over the whole final state; [16:32] over the CORE state only -- holdings, investor positions /
    amounts / stakes and the liquidation log.  A replication the model itself stops (Q6: `shuffle(...)[1:neggs]` throws at
    functions.jl:282, in the middle of a day) is compared on the core state: that is what exists on both sides at the
    throw (the day's sales executed and disbursed); prices / NAVs of the unfinished day are not part of any result the
    reference would have returned."""
    full, core = hashlib.blake2b(digest_size=16), hashlib.blake2b(digest_size=16)
    for a in (holdings, inv_amount, inv_stake):
        b = canon(a)
        full.update(b); core.update(b)
    for a in (price_now, nav_close, market_now):
        full.update(canon(a))
    fb = np.ascontiguousarray(inv_fund, dtype=np.int64).tobytes()
    full.update(fb); core.update(fb)
    for k in ("liquidated", "replacedby", "failtime", "ninvestors", "nstocks"):
        b = np.ascontiguousarray(log[k], dtype=np.int64).tobytes()
        full.update(b); core.update(b)
    for k in ("initialsize", "liquidity"):
        b = canon(log[k])
        full.update(b); core.update(b)
    return full.digest() + core.digest()


def oracle_state(w, p, t_last):
    """The oracle's final state in the shape of Engine.download_compact."""
    K = p.bigk
    pos = np.argmax(w.inv_assets != 0, axis=1)
    has = (w.inv_assets != 0).any(axis=1)
    inv_fund = np.where(has, pos, K).astype(np.int64)
    amount = np.where(has, w.inv_assets[np.arange(p.bign), pos], 0.0)
    stake = np.where(inv_fund < K, w.stakes[np.minimum(inv_fund, K - 1), np.arange(p.bign)], 0.0)
    return dict(holdings=w.holdings, price_now=w.stock_value[:, t_last - 1], inv_fund=inv_fund, inv_amount=amount, inv_stake=stake,
                nav_close=w.nav_close, market_now=w.market[t_last - 1:t_last])


def run_oracle(orc, p, reps, t_last, selector, threads, want_events=False):
    """R oracle replications on host threads.  Returns digests [R,16] (+ per-replication event tapes when asked)."""
    out = np.zeros((reps, 32), dtype=np.uint8)
    info = np.zeros((reps, 3), dtype=np.int64)  # rc, floor hits, NaN flag
    tapes = [None] * reps
    nxt = [0]
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                r = nxt[0]
                nxt[0] += 1
            if r >= reps:
                return
            w = orc.World(p, seed=SEED, rep=REP0 + r)
            if want_events:
                w.enable_trace()
            w.initialise()
            rc = w.modelrun(t_last=t_last, selector=selector)
            st = oracle_state(w, p, t_last)
            out[r] = np.frombuffer(digest(log=w.log(), **st), dtype=np.uint8)
            info[r] = (rc, w.floor_hits(), int(np.isnan(st["price_now"]).any()))
            if want_events:
                tapes[r] = w.trace()
            w.close()

    ths = [threading.Thread(target=work) for _ in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    return out, info, tapes


def events_from_trace_fast(L, orc, tr):
    """helpers.events_from_trace on a structured array: (site, t, a, b, value) records of the recording's event draws."""
    recs, ordinal, pick = [], {}, {}
    for t, kind, a, b in zip(tr["t"].tolist(), tr["kind"].tolist(), tr["a"].tolist(), tr["b"].tolist()):
        if kind == orc.EV_SPAWN:
            i = ordinal.get(t, 0)
            ordinal[t] = i + 1
            recs.append((L.TAPE_ANCHOR, t, i, 0, float(b)))
            pick[t] = (i, 0)
        elif kind == orc.EV_SPAWN_N:
            recs.append((L.TAPE_RPSIZE, t, pick[t][0], 0, float(b)))
        elif kind == orc.EV_PICK:
            i, j = pick[t]
            recs.append((L.TAPE_RPPICK, t, i, j, float(b)))
            pick[t] = (i, j + 1)
        elif kind == orc.EV_CHOICE:
            recs.append((L.TAPE_CHOICE_IDX, t, a, 0, float(b)))
    return recs


def engine_digests(eng, reps):
    out = np.zeros((reps, 32), dtype=np.uint8)
    for r in range(reps):
        c = eng.download_compact(r)
        out[r] = np.frombuffer(digest(c["holdings"], c["price_now"], c["inv_fund"], c["inv_amount"], c["inv_stake"], c["nav_close"],
                                      c["market_now"], eng.download_log(r)), dtype=np.uint8)
    return out


def run_hosttape(fsb, orc, p, reps, t_last, selector, threads):
    import torch
    K, M, N = p.bigk, p.bigm, p.bign
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory().numpy()
    bufs = [(pin(reps), pin(reps, M), pin(reps, N)) for _ in range(2)]
    nav = pin(reps, K)
    t0 = time.perf_counter()
    with fsb.Engine(p, reps, rep_offset=REP0) as eng:
        eng.init_device(SEED)
        t_first = p.perfwindow[-1] + 2
        orc.fill_step_tape(SEED, REP0, t_first, *bufs[0], threads=threads)
        for t in range(t_first, t_last + 1):
            a, b, c = bufs[(t - t_first) & 1]
            nxt = bufs[(t - t_first + 1) & 1]
            filler = threading.Thread(target=orc.fill_step_tape, args=(SEED, REP0, t + 1) + nxt, kwargs=dict(threads=threads))
            if t < t_last:
                filler.start()  # the next day's tape is drawn while this day runs on the device
            eng.run_step_hosttape(t, a, b, c, nav, selector=selector)
            if t < t_last:
                filler.join()
        flags = eng.rep_status()
        st = eng.stats()[0]
        dig = engine_digests(eng, reps)
    return dig, flags, st, time.perf_counter() - t0


def run_events(fsb, orc, p, reps, t_last, selector, tapes, threads):
    """Dense per-replication tapes (the oracle's Philox shocks) + the recording's event draws."""
    L = fsb._lib
    T, M, N = p.bigt, p.bigm, p.bign
    t0 = time.perf_counter()
    with fsb.Engine(p, reps, rep_offset=REP0) as eng:
        eng.init_device(SEED)
        eng.set_seed(SEED ^ 0x5A5A5A5A)  # a different key: every draw that is NOT on a tape would show up as a mismatch
        zm, zs, ur = np.zeros(1), np.zeros((1, M)), np.zeros((1, N))
        for r in range(reps):
            z_mkt, z_stock, u_rev = np.zeros(T), np.zeros((M, T), order="F"), np.zeros((N, T), order="F")
            for t in range(p.perfwindow[-1] + 2, t_last + 1):
                orc.fill_step_tape(SEED, REP0 + r, t, zm, zs, ur)
                z_mkt[t - 1], z_stock[:, t - 1], u_rev[:, t - 1] = zm[0], zs[0], ur[0]
            eng.set_dense_tape(r, z_mkt, z_stock, u_rev)
            eng.set_event_tape(r, events_from_trace_fast(L, orc, tapes[r]))
        try:
            eng.run(t_last=t_last, selector=selector)
        except L.FsbError as ex:
            if ex.code != L.ERR_REPLICATION:
                raise
        flags = eng.rep_status()
        dig = engine_digests(eng, reps)
    return dig, flags, time.perf_counter() - t0


def mismatches(dig, odig, q6):
    """Replication ids whose digests differ: the whole state, or the core state for the replications the model stopped."""
    full_bad = (dig[:, :16] != odig[:, :16]).any(axis=1)
    core_bad = (dig[:, 16:] != odig[:, 16:]).any(axis=1)
    return np.flatnonzero(np.where(q6, core_bad, full_bad))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=256)
    ap.add_argument("--event-reps", type=int, default=0, help="replications of the event-tape leg (0 = skip)")
    ap.add_argument("--days", type=int, default=0, help="simulated days (0 = the full horizon)")
    ap.add_argument("--selector", default="probabilistic")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--oracle-out", default="")
    ap.add_argument("--oracle-in", default="")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from __graft_entry__ import load_package
    fsb = load_package()
    from oracle import oracle as orc
    orc.build()
    p = fsb.Params.default()
    t_last = p.bigt if a.days <= 0 else p.perfwindow[-1] + 1 + a.days
    res = {"config": "BASELINE.json configs[1]: default Params shape, equivalence mode", "replications": a.reps,
           "days": t_last - p.perfwindow[-1] - 1, "selector": a.selector, "seed": SEED, "first_global_replication": REP0}
    if a.oracle_in:
        z = np.load(a.oracle_in)
        assert int(z["t_last"]) == t_last and z["digests"].shape[0] >= a.reps and str(z["selector"]) == a.selector
        odig, oinfo = z["digests"][:a.reps], z["info"][:a.reps]
        res["oracle"] = f"digests read from {os.path.relpath(a.oracle_in, ROOT)}"
    else:
        t0 = time.perf_counter()
        odig, oinfo, _ = run_oracle(orc, p, a.reps, t_last, a.selector, a.threads)
        res["oracle"] = f"computed here on {a.threads} host threads in {time.perf_counter() - t0:.1f} s"
    res["oracle_counts"] = {"q6_stops": int((oinfo[:, 0] != 0).sum()), "nan_replications": int(oinfo[:, 2].sum()),
                            "replications_with_floor_hits": int((oinfo[:, 1] > 0).sum()), "floor_hits": int(oinfo[:, 1].sum())}
    if a.oracle_out:
        np.savez_compressed(a.oracle_out, digests=odig, info=oinfo, t_last=t_last, selector=a.selector, seed=SEED, rep0=REP0)
        print(json.dumps(res))
        return
    dig, flags, st, secs = run_hosttape(fsb, orc, p, a.reps, t_last, a.selector, a.threads)
    q6 = oinfo[:, 0] != 0
    q6_same = bool((((flags & fsb._lib.REP_Q6_NO_ANCHOR) != 0) == q6).all())   # the model's own stops hit the same replications
    bad = mismatches(dig, odig, q6)
    res["hosttape"] = {"mismatching_replications": int(len(bad)), "first_mismatches": bad[:8].tolist(), "seconds": round(secs, 1),
                       "replications_compared_on_the_whole_state": int((~q6).sum()),
                       "q6_stopped_replications_compared_on_the_core_state": int(q6.sum()), "q6_stops_on_the_same_replications": q6_same,
                       "engine_counts": {"q6_stops": int(((flags & fsb._lib.REP_Q6_NO_ANCHOR) != 0).sum()),
                                         "other_flags": int(((flags & ~(fsb._lib.REP_Q6_NO_ANCHOR | fsb._lib.REP_TRACE_OVERFLOW)) != 0).sum()),
                                         "nan_replications": st["nan_replications"], "floor_hits": st["floor_hits"],
                                         "divestments": st["divestments"], "liquidations": st["liquidations"]}}
    ok = len(bad) == 0 and q6_same and res["hosttape"]["engine_counts"]["other_flags"] == 0
    if a.event_reps > 0:
        ne = min(a.event_reps, a.reps)
        edig_o, einfo, tapes = run_oracle(orc, p, ne, t_last, a.selector, a.threads, want_events=True)
        assert (edig_o == odig[:ne]).all()
        edig, eflags, esecs = run_events(fsb, orc, p, ne, t_last, a.selector, tapes, a.threads)
        # a replication the model stops (Q6) freezes at the day of the stop on both sides; tape flags would be a harness bug
        ebad = mismatches(edig, edig_o, q6[:ne])
        res["events"] = {"replications": ne, "mismatching_replications": int(len(ebad)), "first_mismatches": ebad[:8].tolist(),
                         "tape_flags": int(((eflags & (fsb._lib.REP_TAPE_MISSING | fsb._lib.REP_TAPE_INVALID)) != 0).sum()),
                         "seconds": round(esecs, 1)}
        ok = ok and len(ebad) == 0 and res["events"]["tape_flags"] == 0
    res["equivalent"] = bool(ok)
    line = json.dumps(res)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
