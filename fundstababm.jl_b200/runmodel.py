"""runmodel -- mirror of the reference's driver (src/runmodel.jl:6-68) for the hot path.

Allocates the zeroed structs (:13-28), calls ``Func.initialise`` (:30) and ``Func.modelrun``
(:36) -- both on the GPU -- and returns what the reference only prints, including the stylised
facts of :47 (computed on the device from the histories the run leaves in HBM).  Post-processing
that is out of scope (logistic regression :44, plots :50-67) is not performed here; the returned
arrays are exactly what those functions consume.
"""
from __future__ import annotations

from dataclasses import dataclass

from . import func as Func
from . import params as Params
from . import types as Types


@dataclass
class RunResult:
    liquidationlog: object
    market: Types.MarketIndex
    stocks: Types.Equity
    investors: Types.RetailInvestor
    funds: Types.EquityFund
    boundsfailures: list
    stylefacts: dict


def runmodel(params=None, *, fundselector="probabilistic", doplot=False, boundstest=False, seed=0, device=0,
             verbose=True):
    params = Params.default() if params is None else params
    if doplot:
        raise NotImplementedError("plots and parameters.txt (src/runmodel.jl:50-67) are out of scope")
    if verbose:  # banner of src/runmodel.jl:9-11 (labels as printed by the reference)
        print("Run for ", params.bigt, " periods.", sep="")
        print("Sizes: ", params.bign, " investors, ", params.bigm, " funds, ", params.bigk, " stocks.", sep="")
        print("Stock impact range:", params.impactrange, ".", sep="")
    market, stocks, investors, funds = Types.allocate(params)
    Func.initialise(market, stocks, investors, funds, params, seed=seed, device=device)
    failures = []
    if boundstest:
        failures += Func.boundstest(market, stocks, investors, funds)
    stylefacts = {}
    log = Func.modelrun(market, stocks, investors, funds, params, fundselector, seed=seed, device=device,
                        stylefacts=stylefacts)
    if verbose:
        print(log)
    if boundstest:
        failures += Func.boundstest(market, stocks, investors, funds)
    return RunResult(log, market, stocks, investors, funds, failures, stylefacts)
