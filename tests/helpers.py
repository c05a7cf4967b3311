"""Shared helpers for the parity tests: build the same world in the oracle and in the engine."""
import numpy as np


def eq(a, b):
    """Bit-level parity up to NaN payloads: equal values, NaN where the other side is NaN (the
    reference model itself produces NaN trajectories, Q12: reinvestment into a zero-value fund)."""
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def small_params(fsb, **kw):
    base = dict(bigk=6, bign=40, bigm=9, bigt=60, perfwindow=(1, 7), portfsizerange=(2, 6),
                reviewprobability=0.2, invcaprange=(10, 1000))
    base.update(kw)
    return fsb.Params.default(**base)


def oracle_world(orc, params, seed, rep, trace=True):
    w = orc.World(params, seed=seed, rep=rep)
    if trace:
        w.enable_trace()
    w.initialise()
    return w


def structs_from_world(fsb, w):
    T = fsb.Types
    market = T.MarketIndex(w.market.copy())
    stocks = T.Equity(w.stock_value.copy(), w.beta.copy(), w.vol.copy(), w.impact.copy(), w.volume.copy())
    investors = T.RetailInvestor(w.inv_assets.copy(), w.horizon.copy(), w.threshold.copy())
    funds = T.EquityFund(w.holdings.copy(), w.stakes.copy(), w.fund_value.copy())
    return market, stocks, investors, funds


def assert_state_equal(st, w, *, fund_value=True, exact=True):
    """st: Engine.download_state dict; w: oracle World.  Bit-exact unless exact=False (1e-12 rel)."""
    pairs = [("market", w.market), ("stock_value", w.stock_value), ("beta", w.beta), ("vol", w.vol),
             ("impact", w.impact), ("volume", w.volume), ("inv_assets", w.inv_assets), ("threshold", w.threshold),
             ("holdings", w.holdings), ("stakes", w.stakes)]
    if fund_value:
        pairs.append(("fund_value", w.fund_value))
    for name, ref in pairs:
        got = st[name]
        if exact:
            bad = ~((got == ref) | (np.isnan(got) & np.isnan(ref)))
            assert not bad.any(), f"{name}: {int(bad.sum())} mismatching entries, first at {np.argwhere(bad)[0]}: " \
                                  f"{got[tuple(np.argwhere(bad)[0])]!r} vs {ref[tuple(np.argwhere(bad)[0])]!r}"
        else:
            np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0, err_msg=name)
    assert np.array_equal(st["horizon"], w.horizon), "horizon"


def assert_log_equal(log, wlog):
    for k in ("liquidated", "replacedby", "failtime", "ninvestors", "nstocks"):
        assert eq(log[k], wlog[k]), k
    for k in ("initialsize", "liquidity"):
        assert eq(log[k], wlog[k]), k


def assert_trace_equal(tr, w, t_range):
    ev, oe = tr["events"], w.trace()
    assert len(ev) == len(oe), f"event count {len(ev)} vs {len(oe)}"
    for f in ("t", "kind", "a", "b"):
        bad = np.nonzero(ev[f] != oe[f])[0]
        assert len(bad) == 0, f"event field {f} differs first at record {bad[0]}: {ev[bad[0]]} vs {oe[bad[0]]}"
    assert eq(ev["v"], oe["v"]), "event values"
    cols = slice(t_range[0] - 1, t_range[1])
    assert eq(tr["nav_open"][:, cols], w.trace_nav_open[:, cols]), "nav_open"
    assert eq(tr["p_sell"][:, cols], w.trace_p_sell[:, cols]), "p_sell"
    assert eq(tr["p_resp"][:, cols], w.trace_p_resp[:, cols]), "p_resp"


STATE_FIELDS = ("market", "stock_value", "beta", "vol", "impact", "volume", "inv_assets", "horizon", "threshold",
                "holdings", "stakes", "fund_value")


def world_state(w):
    return {k: np.array(getattr(w, k), copy=True) for k in STATE_FIELDS}


def events_from_trace(fsb, orc, w):
    """The oracle's own anchors / portfolio sizes and picks / fund choices as event-tape records
    (site, t, a, b, value), ids 0-based -- what julia/record_tapes.jl captures from the reference."""
    L = fsb._lib
    recs, ordinal, pick_j = [], {}, {}
    for e in w.trace():
        t = int(e["t"])
        if e["kind"] == orc.EV_SPAWN:
            i = ordinal.get(t, 0)
            ordinal[t] = i + 1
            recs.append((L.TAPE_ANCHOR, t, i, 0, float(e["b"])))
            pick_j[t] = (i, 0)
        elif e["kind"] == orc.EV_SPAWN_N:
            recs.append((L.TAPE_RPSIZE, t, pick_j[t][0], 0, float(e["b"])))
        elif e["kind"] == orc.EV_PICK:
            i, j = pick_j[t]
            recs.append((L.TAPE_RPPICK, t, i, j, float(e["b"])))
            pick_j[t] = (i, j + 1)
        elif e["kind"] == orc.EV_CHOICE:
            recs.append((L.TAPE_CHOICE_IDX, t, int(e["a"]), 0, float(e["b"])))
    return recs


def oracle_recording(fsb, orc, p, seed, rep, selector):
    """A recording in the format of julia/record_tapes.jl, with the oracle standing in for the reference:
    numpy shocks as dense tapes, the oracle's event draws read back from its trace."""
    rng = np.random.default_rng(seed)
    dense = dict(z_mkt=rng.standard_normal(p.bigt), z_stock=rng.standard_normal((p.bigm, p.bigt)),
                 u_review=rng.random((p.bign, p.bigt)))
    w = oracle_world(orc, p, seed, rep)
    init = world_state(w)
    w.set_dense_tape(dense["z_mkt"], dense["z_stock"], dense["u_review"])
    assert w.modelrun(selector=selector) == 0
    w.finalise_fund_value(p.bigt)
    rec = dict(W=p.perfwindow[-1], reviewprobability=p.reviewprobability, selector=selector, seed=seed, init=init,
               dense=dense, events=events_from_trace(fsb, orc, w), final=world_state(w), log=w.log())
    return rec, w
