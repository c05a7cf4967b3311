"""Params -- host-side mirror of the reference's ``module Params``.

Follows /root/reference/src/params.jl:4-31 (``Params.default(; overrides...)``, a
``Parameters.@with_kw`` NamedTuple constructor with 19 fields).  Field names, defaults and the
meaning of every field are the reference's; Julia ranges become :class:`UnitRange` /
:class:`StepRange`, whose elements are computed in exact decimal arithmetic so that they equal
what Julia's ``StepRangeLen`` (twice-precision) yields, e.g. ``(0.00075:0.0001:0.0075)[2] ==
0.00085``.
"""
from __future__ import annotations

from dataclasses import dataclass, field, fields, replace
from fractions import Fraction
from typing import Tuple


@dataclass(frozen=True)
class UnitRange:
    """Julia ``lo:hi`` (inclusive integer range)."""
    lo: int
    hi: int

    def __len__(self):
        return max(0, self.hi - self.lo + 1)

    def __getitem__(self, i):
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        return self.lo + i

    def __iter__(self):
        return iter(range(self.lo, self.hi + 1))

    def __repr__(self):
        return f"{self.lo}:{self.hi}"


@dataclass(frozen=True)
class StepRange:
    """Julia ``start:step:stop`` over Float64 (``StepRangeLen``); decimal-exact elements."""
    start: float
    step: float
    stop: float

    def _fr(self):
        return Fraction(repr(self.start)), Fraction(repr(self.step)), Fraction(repr(self.stop))

    def __len__(self):
        a, s, b = self._fr()
        if s == 0 or (b - a) / s < 0:
            return 0
        return int((b - a) // s) + 1

    def values(self):
        a, s, _ = self._fr()
        return [float(a + i * s) for i in range(len(self))]

    def __getitem__(self, i):
        return self.values()[i]

    def __iter__(self):
        return iter(self.values())

    def scaled(self, factor: float) -> "ValueTable":
        return ValueTable(tuple(v * factor for v in self.values()))

    def __repr__(self):
        return f"{self.start}:{self.step}:{self.stop}"


@dataclass(frozen=True)
class ValueTable:
    """An explicit table of values used where the reference takes a range (calibration sweeps
    scale ``impactrange``)."""
    table: Tuple[float, ...]

    def __len__(self):
        return len(self.table)

    def values(self):
        return list(self.table)

    def __getitem__(self, i):
        return self.table[i]

    def __iter__(self):
        return iter(self.table)


@dataclass(frozen=True)
class ParamSet:
    # src/params.jl:5-8
    bigm: int = 250   # Number of stocks
    bign: int = 1000  # Number of investors
    bigt: int = 1350  # Number of days
    bigk: int = 250   # Number of funds
    # Market parameters, :11-13
    marketstartval: float = 100
    drift: float = 0.000307888
    marketvol: float = 0.01
    # Stock parameters, :16-20
    impactrange: object = field(default_factory=lambda: StepRange(0.00075, 0.0001, 0.0075))
    betamean: float = 1
    betastd: float = 0.3
    stockstartval: float = 100
    stockvolrange: object = field(default_factory=lambda: StepRange(0.001, 0.001, 0.01))
    # Investor parameters, :23-27
    reviewprobability: float = 1 / 63
    perfwindow: UnitRange = field(default_factory=lambda: UnitRange(1, 100))
    invcaprange: Tuple[int, int] = (50, 150)
    thresholdmean: float = 0
    thresholdstd: float = 0.05
    # Fund parameters, :30-31
    portfsizerange: UnitRange = field(default_factory=lambda: UnitRange(10, 100))
    plotpath: str = "./plots/"

    def __call__(self, **overrides) -> "ParamSet":
        """``testparams(perfwindow=3:3)``-style override, as ``@with_kw`` constructors allow."""
        return replace(self, **_coerce(overrides))

    def validate(self) -> None:
        """The sanity inequalities of /root/reference/test/params_test.jl:3-24."""
        if not (self.bign >= self.bigk >= 1 and self.bigm >= 1):
            raise ValueError("need bign >= bigk >= 1 and bigm >= 1")
        pw = self.perfwindow
        if not (0 < pw[0] <= pw[-1] < self.bigt):
            raise ValueError("need 0 < perfwindow[1] <= perfwindow[end] < bigt")
        ps = self.portfsizerange
        if not (1 <= ps[0] <= ps[-1]):
            raise ValueError("need 1 <= portfsizerange[1] <= portfsizerange[end]")
        if not (0 <= self.reviewprobability <= 1):
            raise ValueError("reviewprobability must be a probability")
        if self.invcaprange[0] > self.invcaprange[1] or self.invcaprange[0] < 0:
            raise ValueError("invcaprange must be a non-negative (lo, hi)")
        if len(self.impactrange) < 1 or len(self.stockvolrange) < 1:
            raise ValueError("impactrange / stockvolrange must be non-empty")


def _coerce(kw):
    out = dict(kw)
    for name in ("perfwindow", "portfsizerange"):
        v = out.get(name)
        if isinstance(v, (tuple, list)) and len(v) == 2:
            out[name] = UnitRange(int(v[0]), int(v[1]))
        elif isinstance(v, range):
            out[name] = UnitRange(v.start, v.stop - 1)
    for name in ("impactrange", "stockvolrange"):
        v = out.get(name)
        if isinstance(v, (tuple, list)) and len(v) == 3 and not isinstance(v, ValueTable):
            out[name] = StepRange(float(v[0]), float(v[1]), float(v[2]))
    known = {f.name for f in fields(ParamSet)}
    bad = set(out) - known
    if bad:
        raise TypeError(f"unknown parameter(s): {sorted(bad)}")
    return out


def default(**overrides) -> ParamSet:
    """``Params.default(; overrides...)``, /root/reference/src/params.jl:4."""
    return ParamSet(**_coerce(overrides))
