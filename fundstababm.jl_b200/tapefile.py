"""The file julia/record_tapes.jl writes (SURVEY.md 8c "tape"): one replication of the reference --
state after Func.initialise, every draw of the time loop as dense / event tapes, state and liquidation
log after the last day.  Little endian; Float64 / Int64 arrays in Julia's column-major order.

    magic "FSBTAPE1" | int64[12]: K M N T W selector n_events n_log has_final seed bits(reviewprobability) 0
    state: market[T] stock_value[MxT] beta[M] vol[M] impact[M] volume[MxT] inv_assets[Nx(K+1)]
           horizon(int64)[N] threshold[N] holdings[KxM] stakes[KxN] fund_value[KxT]
    dense tapes: z_mkt[T] z_stock[MxT] u_review[NxT]
    events: n_events x (int64 site, t, a, b ; float64 value)      ids 0-based, sites as include/fsb.h
    final state (as above) | log: liquidated replacedby (int64) initialsize liquidity (f64)
                                   failtime ninvestors nstocks (int64), n_log rows each
"""
from __future__ import annotations

import numpy as np

MAGIC = b"FSBTAPE1"
SELECTOR_NAMES = ("best", "goodenough", "probabilistic", "random")
STATE_FIELDS = ("market", "stock_value", "beta", "vol", "impact", "volume", "inv_assets", "horizon", "threshold",
                "holdings", "stakes", "fund_value")
LOG_FIELDS = (("liquidated", np.int64), ("replacedby", np.int64), ("initialsize", np.float64), ("liquidity", np.float64),
              ("failtime", np.int64), ("ninvestors", np.int64), ("nstocks", np.int64))


def _state_shapes(K, M, N, T):
    return {"market": (T,), "stock_value": (M, T), "beta": (M,), "vol": (M,), "impact": (M,), "volume": (M, T),
            "inv_assets": (N, K + 1), "horizon": (N,), "threshold": (N,), "holdings": (K, M), "stakes": (K, N),
            "fund_value": (K, T)}


def _write_state(f, st, shapes):
    for name in STATE_FIELDS:
        a = np.asarray(st[name], dtype=np.int64 if name == "horizon" else np.float64)
        assert a.shape == shapes[name], (name, a.shape, shapes[name])
        f.write(np.asfortranarray(a).tobytes(order="F"))


def _read_state(buf, off, shapes):
    st = {}
    for name in STATE_FIELDS:
        dt = np.int64 if name == "horizon" else np.float64
        n = int(np.prod(shapes[name]))
        st[name] = np.frombuffer(buf, dtype=dt, count=n, offset=off).reshape(shapes[name], order="F").copy()
        off += 8 * n
    return st, off


def write_tape_file(path, *, W, reviewprobability, selector, seed, init, dense, events, final, log):
    """init/final: dicts of STATE_FIELDS; dense: z_mkt, z_stock, u_review; events: (site,t,a,b,value) tuples."""
    K, M = init["holdings"].shape
    N, T = init["inv_assets"].shape[0], init["market"].shape[0]
    shapes = _state_shapes(K, M, N, T)
    n_log = len(log["liquidated"])
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(np.array([K, M, N, T, int(W), SELECTOR_NAMES.index(selector), len(events), n_log, 1, seed,
                          int(np.array([reviewprobability], dtype="<f8").view("<i8")[0]), 0], dtype="<i8").tobytes())
        _write_state(f, init, shapes)
        for name, shp in (("z_mkt", (T,)), ("z_stock", (M, T)), ("u_review", (N, T))):
            a = np.asarray(dense[name], dtype=np.float64)
            assert a.shape == shp, (name, a.shape)
            f.write(np.asfortranarray(a).tobytes(order="F"))
        for site, t, a, b, v in events:
            f.write(np.array([site, t, a, b], dtype="<i8").tobytes())
            f.write(np.array([v], dtype="<f8").tobytes())
        _write_state(f, final, shapes)
        for name, dt in LOG_FIELDS:
            f.write(np.asarray(log[name], dtype=dt).tobytes())


def read_tape_file(path):
    buf = open(path, "rb").read()
    if buf[:8] != MAGIC:
        raise ValueError(f"{path}: not a FSBTAPE1 file")
    hdr = np.frombuffer(buf, dtype="<i8", count=12, offset=8)
    K, M, N, T, W, sel, n_ev, n_log, has_final, seed = (int(x) for x in hdr[:10])
    reviewprobability = float(hdr[10:11].view("<f8")[0])
    shapes = _state_shapes(K, M, N, T)
    off = 8 + 96
    init, off = _read_state(buf, off, shapes)
    dense = {}
    for name, shp in (("z_mkt", (T,)), ("z_stock", (M, T)), ("u_review", (N, T))):
        n = int(np.prod(shp))
        dense[name] = np.frombuffer(buf, dtype="<f8", count=n, offset=off).reshape(shp, order="F").copy()
        off += 8 * n
    events = []
    for _ in range(n_ev):
        k = np.frombuffer(buf, dtype="<i8", count=4, offset=off)
        v = np.frombuffer(buf, dtype="<f8", count=1, offset=off + 32)
        events.append((int(k[0]), int(k[1]), int(k[2]), int(k[3]), float(v[0])))
        off += 40
    final, log = None, None
    if has_final:
        final, off = _read_state(buf, off, shapes)
        log = {}
        for name, dt in LOG_FIELDS:
            log[name] = np.frombuffer(buf, dtype=dt, count=n_log, offset=off).copy()
            off += 8 * n_log
    if off != len(buf):
        raise ValueError(f"{path}: {len(buf) - off} trailing bytes")
    return dict(K=K, M=M, N=N, T=T, W=W, reviewprobability=reviewprobability, selector=SELECTOR_NAMES[sel], seed=seed, init=init, dense=dense, events=events,
                final=final, log=log)
