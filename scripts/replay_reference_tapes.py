"""Equivalence mode against a recording of the real reference (julia/record_tapes.jl):
uploads the recorded post-initialise state, feeds the recorded draws as tapes, runs the engine's
Func.modelrun and compares the final state and the liquidation log with what the reference produced.
Discrete outcomes must be identical; FP64 arrays within 1e-12 relative (the reference sums with
BLAS / Base.sum orders the engine does not share, so bit equality is not expected here).

    python scripts/replay_reference_tapes.py recording.fsbtape [--use-uniforms]

--use-uniforms: drive the "probabilistic" choices by the recorded uniforms (tests the restatement of
Distributions.Categorical sampling) instead of forcing the recorded choice indices."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402


def replay(fsb, rec, use_uniforms=False, rtol=1e-12):
    """Runs the engine on a recording (fundstababm.jl_b200.tapefile.read_tape_file) and checks it against the
    recorded outcome.  Returns {array name: worst relative error}; raises AssertionError on a failure."""
    L, T = fsb._lib, fsb.Types
    K, M, N, Tn = rec["K"], rec["M"], rec["N"], rec["T"]
    p = fsb.Params.default(bigk=K, bigm=M, bign=N, bigt=Tn, perfwindow=(1, rec["W"]), reviewprobability=rec["reviewprobability"])
    i = rec["init"]
    market = T.MarketIndex(i["market"].copy())
    stocks = T.Equity(i["stock_value"].copy(), i["beta"].copy(), i["vol"].copy(), i["impact"].copy(), i["volume"].copy())
    investors = T.RetailInvestor(i["inv_assets"].copy(), i["horizon"].copy(), i["threshold"].copy())
    funds = T.EquityFund(i["holdings"].copy(), i["stakes"].copy(), i["fund_value"].copy())
    drop = L.TAPE_CHOICE_IDX if use_uniforms else L.TAPE_CHOICE_U
    events = [e for e in rec["events"] if e[0] != drop]
    log = fsb.Func.modelrun(market, stocks, investors, funds, p, rec["selector"], seed=rec["seed"], rep=0,
                            dense_tape=rec["dense"], event_tape=events, as_dataframe=False)
    f = rec["final"]
    got = {"market": market.value, "stock_value": stocks.value, "volume": stocks.volume, "inv_assets": investors.assets,
           "holdings": funds.holdings, "stakes": funds.stakes, "fund_value": funds.value}
    worst = {}
    for name, g in got.items():
        want = f[name]
        assert np.array_equal(np.isnan(g), np.isnan(want)), f"{name}: NaN pattern differs"
        assert np.array_equal(g == 0, want == 0), f"{name}: zero pattern differs (a discrete outcome)"
        with np.errstate(invalid="ignore", divide="ignore"):
            rel = np.abs(g - want) / np.abs(want)
        rel = rel[np.isfinite(rel)]
        worst[name] = float(rel.max()) if rel.size else 0.0
        assert worst[name] <= rtol, f"{name}: relative error {worst[name]:.3e} > {rtol:g}"
    for name in ("liquidated", "replacedby", "failtime", "ninvestors", "nstocks"):
        assert np.array_equal(log[name], rec["log"][name]), f"log column {name} differs"
    for name in ("initialsize", "liquidity"):
        np.testing.assert_allclose(log[name], rec["log"][name], rtol=rtol, atol=0, err_msg=name)
    return worst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("recording")
    ap.add_argument("--use-uniforms", action="store_true")
    a = ap.parse_args()
    fsb = load_package()
    from importlib import import_module
    rec = import_module(fsb.__name__ + ".tapefile").read_tape_file(a.recording)
    worst = replay(fsb, rec, a.use_uniforms)
    print(f"{a.recording}: K={rec['K']} M={rec['M']} N={rec['N']} T={rec['T']} selector={rec['selector']}: "
          f"{len(rec['events'])} event draws replayed; discrete outcomes identical; worst relative errors: "
          + ", ".join(f"{k} {v:.1e}" for k, v in worst.items()))
