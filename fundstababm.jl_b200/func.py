"""Func -- host-side mirror of the hot-path entry points of the reference's ``module Func``.

Only the seam of SURVEY.md 8b is mirrored: ``initialise`` (src/functions.jl:417-430),
``modelrun`` (:446-491) and the runtime invariant check ``boundstest`` (:432-444); and, from SURVEY.md
8 f2, ``calc_stylisedfacts`` (:604-648).  All of them run on the GPU through the C ABI; there is no
CPU implementation here.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine, LOG_COLUMNS

VALID_SELECTORS = ("best", "goodenough", "probabilistic", "random")  # src/functions.jl:344-353


def _fill(market, stocks, investors, funds, st):
    market.value[...] = st["market"]
    stocks.value[...] = st["stock_value"]
    stocks.beta[...] = st["beta"]
    stocks.vol[...] = st["vol"]
    stocks.impact[...] = st["impact"]
    stocks.volume[...] = st["volume"]
    investors.assets[...] = st["inv_assets"]
    investors.horizon[...] = st["horizon"]
    investors.threshold[...] = st["threshold"]
    funds.holdings[...] = st["holdings"]
    funds.stakes[...] = st["stakes"]
    funds.value[...] = st["fund_value"]


def initialise(market, stocks, investors, funds, params, *, seed=0, rep=0, device=0):
    """``Func.initialise(market, stocks, investors, funds, params)`` (src/functions.jl:417-430):
    fills the structs in place.  Draws come from Philox4x32-10 keyed by ``seed`` with the global
    replication id ``rep`` (the reference uses Julia's unseeded global RNG)."""
    with Engine(params, 1, rep_offset=rep, device=device, keep_history=True) as eng:
        eng.init_device(seed)
        _fill(market, stocks, investors, funds, eng.download_state(0))


def modelrun(market, stocks, investors, funds, params, fundselector="bestperformer", *, seed=0, rep=0, device=0,
             dense_tape=None, event_tape=None, flags=0, as_dataframe=True, stylefacts=None):
    """``Func.modelrun(market, stocks, investors, funds, params, fundselector)``
    (src/functions.jl:446-491): runs t = perfwindow[end]+2 : bigt on the device, mutates the
    structs in place (runmodel reads market.value, stocks.value, stocks.volume afterwards,
    src/runmodel.jl:47) and returns the liquidation log (columns of :452-456).

    As in the reference (Q9) the default ``"bestperformer"`` is not a valid selector.
    ``stylefacts``: a dict that receives ``calc_stylisedfacts`` of the finished run, computed on the
    device before the engine is released (``{"_kw": {...}}`` passes maxlag / returnspctile)."""
    if fundselector not in VALID_SELECTORS:
        raise ValueError("Please specify a valid fund selection method in reinvest! "
                         f"(one of {VALID_SELECTORS}, got {fundselector!r})")
    with Engine(params, 1, rep_offset=rep, device=device, keep_history=True) as eng:
        eng.upload_state(0, market, stocks, investors, funds)
        if dense_tape is not None:
            eng.set_dense_tape(0, **dense_tape)
        if event_tape is not None:
            eng.set_event_tape(0, event_tape)
        # the engine's Philox key is set by init_device; for an uploaded state set it explicitly
        eng.set_seed(seed)
        eng.run(selector=fundselector, flags=flags)
        _fill(market, stocks, investors, funds, eng.download_state(0))
        log = eng.download_log(0)
        if stylefacts is not None:  # src/runmodel.jl:47, while the histories are still in HBM
            stylefacts.update(_facts_of(eng, 0, **stylefacts.pop("_kw", {})))
    if as_dataframe:
        import pandas as pd
        return pd.DataFrame({k: log[k] for k in LOG_COLUMNS})
    return log


def calc_zScore(sample, muZero):  # noqa: N802,N803  (reference names, src/functions.jl:650-655)
    sample = np.asarray(sample, dtype=np.float64)
    return (sample.mean() - muZero) / (sample.std(ddof=1) / np.sqrt(len(sample)))


def _facts_of(eng, rep, maxlag=10, returnspctile=5):
    f = {k: v[0] for k, v in eng.stylised_facts(rep, 1, maxlag, returnspctile).items()}
    n_returns = eng.T - eng.W - 1
    # confidence bands of calc_pacf (:588-602): the vector method divides by sqrt(#returns), the matrix method by
    # sqrt(size(.,1)) = sqrt(bigm) -- kept as the reference computes them
    f["marketpacf"] = (f["marketpacf"], 1.96 / np.sqrt(n_returns))
    f["marketvolacluster"] = (f["marketvolacluster"], 1.96 / np.sqrt(n_returns))
    f["stockpacf"] = (f["stockpacf"], 1.96 / np.sqrt(eng.M))
    f["stockvolacluster"] = (f["stockvolacluster"], 1.96 / np.sqrt(eng.M))
    # the z statistics of the OneSampleZTests of :629-634 (calc_zScore :650-655)
    f["zScoreKurtoses"] = calc_zScore(f["stockkurtoses"], 0.0)
    f["zScoreLossGain"] = calc_zScore(f["lossgainratios"], 1.0)
    f["zScoreVolumeVola"] = calc_zScore(f["corrs"], 0.0)
    return f


def calc_stylisedfacts(marketval, stocksval, stocksvolume, params, *, maxlag=10, returnspctile=5, device=0):
    """``Func.calc_stylisedfacts(marketval, stocksval, stocksvolume, params; maxlag, returnspctile)``
    (src/functions.jl:604-648) on the device: the histories are uploaded to a one-replication engine and only
    the per-stock vectors come back.  Returns a dict keyed by the reference's local names (``stockpacf`` ...
    ``corrs``; the PACF entries are ``(coefficients, confinterval)`` pairs as calc_pacf returns them) plus the
    z statistics of the three z-tests.  The exact KS test of :635-641 needs every return on the host and is
    not computed.  After a ``modelrun`` use its ``stylefacts=`` argument instead: no upload at all."""
    from . import types as Types  # noqa: N812
    market, stocks, investors, funds = Types.allocate(params)
    market.value[...] = marketval
    stocks.value[...] = stocksval
    stocks.volume[...] = stocksvolume
    investors.horizon[...] = 1  # a formally valid (empty) investor / fund state: only the histories are read
    with Engine(params, 1, device=device, keep_history=True) as eng:
        eng.upload_state(0, market, stocks, investors, funds)
        return _facts_of(eng, 0, maxlag, returnspctile)


def boundstest(market, stocks, investors, funds):
    """The "Variable Bounds" checks of src/functions.jl:432-444; returns the list of failures."""
    checks = {
        "stocks.value >= 0": np.all(stocks.value >= 0),
        "stocks.vol >= 0": np.all(stocks.vol >= 0),
        "stocks.impact >= 0": np.all(stocks.impact >= 0),
        "funds.holdings >= 0": np.all(funds.holdings >= 0),
        "funds.stakes >= 0": np.all(funds.stakes >= 0),
        "funds.value >= 0": np.all(funds.value >= 0),
        "investors.assets >= 0": np.all(investors.assets >= 0),
        "investors.horizon >= 0": np.all(investors.horizon >= 0),
        "1 >= investors.threshold >= -1": np.all((investors.threshold <= 1) & (investors.threshold >= -1)),
    }
    return [name for name, ok in checks.items() if not ok]
