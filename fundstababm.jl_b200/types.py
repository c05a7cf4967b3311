"""Types -- host-side mirror of the reference's ``module Types`` (src/types.jl:15-50).

Same struct and field names, same shapes (pinned by test/types_test.jl:37-47,65-75,88-94,114-122);
arrays are numpy, Float64/Int64, column-major (``order="F"``) so they cross the C ABI unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def _mat(a):
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if a.ndim != 2:
        raise TypeError("expected a Matrix{Float64}")
    return a


def _vec(a, dtype=np.float64):
    a = np.ascontiguousarray(np.asarray(a, dtype=dtype))
    if a.ndim != 1:
        raise TypeError("expected a Vector")
    return a


@dataclass
class Equity:  # src/types.jl:15-21
    value: np.ndarray   # M x T value history of each stock
    beta: np.ndarray    # M
    vol: np.ndarray     # M
    impact: np.ndarray  # M symmetric price impact for 1 money unit bought/sold
    volume: np.ndarray  # M x T money value traded in a day

    def __post_init__(self):
        self.value, self.volume = _mat(self.value), _mat(self.volume)
        self.beta, self.vol, self.impact = _vec(self.beta), _vec(self.vol), _vec(self.impact)


@dataclass
class MarketIndex:  # src/types.jl:26-28
    value: np.ndarray  # T

    def __post_init__(self):
        self.value = _vec(self.value)


@dataclass
class RetailInvestor:  # src/types.jl:30-34
    assets: np.ndarray     # N x (K+1): stakes in K funds and cash (last column)
    horizon: np.ndarray    # N, Int64 (test/types_test.jl:89)
    threshold: np.ndarray  # N

    def __post_init__(self):
        self.assets = _mat(self.assets)
        self.horizon = _vec(self.horizon, np.int64)
        self.threshold = _vec(self.threshold)


@dataclass
class EquityFund:  # src/types.jl:36-40
    holdings: np.ndarray  # K x M units of each stock
    stakes: np.ndarray    # K x N share of the fund each investor owns
    value: np.ndarray     # K x T fund value history

    def __post_init__(self):
        self.holdings, self.stakes, self.value = _mat(self.holdings), _mat(self.stakes), _mat(self.value)


@dataclass
class BuyMarketOrder:  # src/types.jl:42-45
    values: np.ndarray  # R x M money spent on each stock
    funds: np.ndarray   # R fund that the stocks are bought for

    def __post_init__(self):
        self.values, self.funds = _mat(self.values), _vec(self.funds, np.int64)


@dataclass
class SellMarketOrder:  # src/types.jl:47-50
    values: np.ndarray     # R x M (<= 0)
    investors: np.ndarray  # R investor that stocks are liquidated for

    def __post_init__(self):
        self.values, self.investors = _mat(self.values), _vec(self.investors, np.int64)


def allocate(params):
    """The zeroed structs of runmodel (src/runmodel.jl:13-28)."""
    K, M, N, T = params.bigk, params.bigm, params.bign, params.bigt
    market = MarketIndex(np.zeros(T))
    stocks = Equity(np.zeros((M, T)), np.zeros(M), np.zeros(M), np.zeros(M), np.zeros((M, T)))
    investors = RetailInvestor(np.zeros((N, K + 1)), np.zeros(N, dtype=np.int64), np.zeros(N))
    funds = EquityFund(np.zeros((K, M)), np.zeros((K, N)), np.zeros((K, T)))
    return market, stocks, investors, funds
