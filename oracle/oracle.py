"""ctypes binding of the CPU oracle (oracle/fsb_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package.
Indices in this binding are 0-based; time t / price columns are 1-based as in the reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SEL = {"best": 0, "goodenough": 1, "probabilistic": 2, "random": 3}
FAITHFUL_HISTORY = 1
NO_VOLUME_QUIRK = 2

A_MARKET, A_STOCK_VALUE, A_BETA, A_VOL, A_IMPACT, A_VOLUME, A_INV_ASSETS, A_THRESHOLD, \
    A_HOLDINGS, A_STAKES, A_FUND_VALUE, A_TRACE_NAV_OPEN, A_TRACE_P_SELL, A_TRACE_P_RESP, \
    A_NAV_CLOSE = range(15)

EV_REVIEWER, EV_DIVEST, EV_CASHOUT, EV_FLOOR, EV_SPAWN, EV_SPAWN_N, EV_CHOICE, EV_PICK = range(1, 9)
TAPE_ANCHOR, TAPE_RPSIZE, TAPE_RPPICK, TAPE_CHOICE_U, TAPE_CHOICE_IDX = range(1, 6)

LOG_COLUMNS = ("liquidated", "replacedby", "initialsize", "liquidity", "failtime", "ninvestors", "nstocks")

TRACE_DTYPE = np.dtype([("t", "<i4"), ("kind", "<i4"), ("a", "<i4"), ("b", "<i4"), ("v", "<f8")])
TAPE_DTYPE = np.dtype([("key", "<u8"), ("value", "<f8")])


class OrcParams(C.Structure):
    _fields_ = [
        ("bigm", C.c_int64), ("bign", C.c_int64), ("bigt", C.c_int64), ("bigk", C.c_int64),
        ("marketstartval", C.c_double), ("drift", C.c_double), ("marketvol", C.c_double),
        ("impact_values", C.POINTER(C.c_double)), ("n_impact", C.c_int64),
        ("betamean", C.c_double), ("betastd", C.c_double), ("stockstartval", C.c_double),
        ("vol_values", C.POINTER(C.c_double)), ("n_vol", C.c_int64),
        ("reviewprobability", C.c_double),
        ("perf_lo", C.c_int64), ("perf_hi", C.c_int64),
        ("cap_lo", C.c_int64), ("cap_hi", C.c_int64),
        ("thresholdmean", C.c_double), ("thresholdstd", C.c_double),
        ("psize_lo", C.c_int64), ("psize_hi", C.c_int64),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libfsb_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("fsb_oracle.c", "fsb_oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        P = C.POINTER
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [P(OrcParams)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_ptr.restype = P(C.c_double)
        L.orc_ptr.argtypes = [C.c_void_p, C.c_int]
        L.orc_horizon.restype = P(C.c_int64)
        L.orc_horizon.argtypes = [C.c_void_p]
        L.orc_set_rng.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.orc_set_dense_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_event_tape.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_fill_step_tape.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_tape_key.restype = C.c_uint64
        L.orc_tape_key.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64]
        L.orc_enable_trace.argtypes = [C.c_void_p, C.c_int]
        L.orc_trace_count.restype = C.c_int64
        L.orc_trace_count.argtypes = [C.c_void_p]
        L.orc_trace.restype = C.c_void_p
        L.orc_trace.argtypes = [C.c_void_p]
        L.orc_initialise.argtypes = [C.c_void_p]
        L.orc_modelrun.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int]
        L.orc_finalise_fund_value.argtypes = [C.c_void_p, C.c_int64]
        L.orc_log_rows.restype = C.c_int64
        L.orc_log_rows.argtypes = [C.c_void_p]
        L.orc_log_col.restype = P(C.c_double)
        L.orc_log_col.argtypes = [C.c_void_p, C.c_int]
        L.orc_floor_hits.restype = C.c_int64
        L.orc_floor_hits.argtypes = [C.c_void_p]
        for name in ("orc_fundcapitalinit", "orc_fundstakeinit"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_fundvalinit.argtypes = [C.c_void_p, C.c_int64]
        L.orc_fundrevalue.argtypes = [C.c_void_p, C.c_int64]
        L.orc_perfreview.restype = C.c_int64
        L.orc_perfreview.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_liquidate.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_marketmake.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.orc_executeorder_sell.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_executeorder_buy.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_disburse_cash.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_disburse_shares.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_respawn.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, P(C.c_int64)]
        L.orc_reinvest.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, P(C.c_int64)]
        L.orc_philox4x32_10.argtypes = [P(C.c_uint32), P(C.c_uint32), P(C.c_uint32)]
        L.orc_det_log.restype = C.c_double
        L.orc_det_log.argtypes = [C.c_double]
        L.orc_det_norminv.restype = C.c_double
        L.orc_det_norminv.argtypes = [C.c_double]
        for name in ("orc_draw_normal", "orc_draw_uniform"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_draw_discrete.restype = C.c_uint32
        L.orc_draw_discrete.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_stylisedfacts.argtypes = [C.c_void_p] * 3 + [C.c_int64] * 4 + [C.c_double] + [C.c_void_p] * 7
        L.orc_dotw.restype = C.c_double
        L.orc_dotw.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64]
        _LIB = L
    return _LIB


def _range_values(r):
    return np.asarray(r.values() if hasattr(r, "values") else list(r), dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


class World:
    """One replication's dense state (types.jl:15-50) plus the oracle's functions."""

    def __init__(self, params, seed: int = 0, rep: int = 0):
        L = lib()
        self.params = params
        self._imp = _range_values(params.impactrange)
        self._vol = _range_values(params.stockvolrange)
        p = OrcParams(
            bigm=params.bigm, bign=params.bign, bigt=params.bigt, bigk=params.bigk,
            marketstartval=params.marketstartval, drift=params.drift, marketvol=params.marketvol,
            impact_values=self._imp.ctypes.data_as(C.POINTER(C.c_double)), n_impact=len(self._imp),
            betamean=params.betamean, betastd=params.betastd, stockstartval=params.stockstartval,
            vol_values=self._vol.ctypes.data_as(C.POINTER(C.c_double)), n_vol=len(self._vol),
            reviewprobability=params.reviewprobability,
            perf_lo=params.perfwindow[0], perf_hi=params.perfwindow[-1],
            cap_lo=params.invcaprange[0], cap_hi=params.invcaprange[1],
            thresholdmean=params.thresholdmean, thresholdstd=params.thresholdstd,
            psize_lo=params.portfsizerange[0], psize_hi=params.portfsizerange[-1])
        self._h = C.c_void_p(L.orc_create(C.byref(p)))
        self.K, self.M, self.N, self.T = params.bigk, params.bigm, params.bign, params.bigt
        self._keep = []
        L.orc_set_rng(self._h, seed, rep)
        K, M, N, T = self.K, self.M, self.N, self.T
        self.market = self._view(A_MARKET, (T,))
        self.stock_value = self._view(A_STOCK_VALUE, (M, T))
        self.beta = self._view(A_BETA, (M,))
        self.vol = self._view(A_VOL, (M,))
        self.impact = self._view(A_IMPACT, (M,))
        self.volume = self._view(A_VOLUME, (M, T))
        self.inv_assets = self._view(A_INV_ASSETS, (N, K + 1))
        self.threshold = self._view(A_THRESHOLD, (N,))
        self.holdings = self._view(A_HOLDINGS, (K, M))
        self.stakes = self._view(A_STAKES, (K, N))
        self.fund_value = self._view(A_FUND_VALUE, (K, T))
        self.nav_close = self._view(A_NAV_CLOSE, (K,))
        self.horizon = np.ctypeslib.as_array(lib().orc_horizon(self._h), shape=(N,))

    def _view(self, which, shape):
        ptr = lib().orc_ptr(self._h, which)
        n = int(np.prod(shape))
        flat = np.ctypeslib.as_array(ptr, shape=(n,))
        return flat.reshape(shape, order="F")

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state(self, st):
        """Overwrite the dense state (after-initialise arrays of a recording / of another world)."""
        for name in ("market", "stock_value", "beta", "vol", "impact", "volume", "inv_assets", "threshold", "holdings",
                     "stakes", "fund_value", "horizon"):
            getattr(self, name)[...] = st[name]

    # draws ---------------------------------------------------------------------------------
    def set_rng(self, seed, rep):
        lib().orc_set_rng(self._h, seed, rep)

    def set_dense_tape(self, z_mkt=None, z_stock=None, u_review=None):
        arrs = []
        for a in (z_mkt, z_stock, u_review):
            if a is None:
                arrs.append(None)
            else:
                a = np.asfortranarray(np.asarray(a, dtype=np.float64))
                self._keep.append(a)
                arrs.append(a.ctypes.data)
        lib().orc_set_dense_tape(self._h, *arrs)

    def set_event_tape(self, recs):
        """recs: iterable of (site, t, a, b, value); sorted here by key."""
        L = lib()
        arr = np.zeros(len(recs), dtype=TAPE_DTYPE)
        for i, (site, t, a, b, v) in enumerate(recs):
            arr[i] = (L.orc_tape_key(site, t, a, b), v)
        arr.sort(order="key")
        self._keep.append(arr)
        L.orc_set_event_tape(self._h, arr.ctypes.data, len(arr))
        return arr

    # trace ---------------------------------------------------------------------------------
    def enable_trace(self, on=True):
        lib().orc_enable_trace(self._h, int(on))
        if on:
            self.trace_nav_open = self._view(A_TRACE_NAV_OPEN, (self.K, self.T))
            self.trace_p_sell = self._view(A_TRACE_P_SELL, (self.M, self.T))
            self.trace_p_resp = self._view(A_TRACE_P_RESP, (self.M, self.T))

    def trace(self):
        n = lib().orc_trace_count(self._h)
        if n == 0:
            return np.zeros(0, dtype=TRACE_DTYPE)
        buf = (C.c_char * (n * TRACE_DTYPE.itemsize)).from_address(lib().orc_trace(self._h))
        return np.frombuffer(buf, dtype=TRACE_DTYPE, count=n).copy()

    # model ---------------------------------------------------------------------------------
    def initialise(self):
        return lib().orc_initialise(self._h)

    def modelrun(self, t_first=None, t_last=None, selector="probabilistic", flags=0):
        if t_first is None:
            t_first = self.params.perfwindow[-1] + 2
        if t_last is None:
            t_last = self.T
        if selector not in SEL:
            raise ValueError("Please specify a valid fund selection method")
        return lib().orc_modelrun(self._h, t_first, t_last, SEL[selector], flags)

    def finalise_fund_value(self, t_now):
        lib().orc_finalise_fund_value(self._h, t_now)

    def log(self):
        n = lib().orc_log_rows(self._h)
        out = {}
        for c, name in enumerate(LOG_COLUMNS):
            col = np.ctypeslib.as_array(lib().orc_log_col(self._h, c), shape=(n,)).copy() if n else np.zeros(0)
            out[name] = col if name in ("initialsize", "liquidity") else col.astype(np.int64)
        return out

    def floor_hits(self):
        return lib().orc_floor_hits(self._h)

    # per-function entry points ---------------------------------------------------------------
    def fundcapitalinit(self):
        lib().orc_fundcapitalinit(self._h)

    def fundstakeinit(self):
        lib().orc_fundstakeinit(self._h)

    def fundvalinit(self, horizon):
        lib().orc_fundvalinit(self._h, horizon)

    def fundrevalue(self, targets):
        for k in targets:
            lib().orc_fundrevalue(self._h, int(k))

    def perfreview(self, t, reviewers):
        rv = _i64(reviewers)
        di = np.zeros(max(1, len(rv) * self.K), dtype=np.int64)
        df = np.zeros_like(di)
        n = lib().orc_perfreview(self._h, t, rv.ctypes.data, len(rv), di.ctypes.data, df.ctypes.data)
        return np.stack([di[:n], df[:n]], axis=1)

    def liquidate(self, col, divestments):
        d = _i64(divestments).reshape(-1, 2)
        di, df = _i64(d[:, 0]), _i64(d[:, 1])
        sale = np.zeros((len(d), self.M))
        lib().orc_liquidate(self._h, col, di.ctypes.data, df.ctypes.data, len(d), sale.ctypes.data)
        return sale

    def marketmake(self, t, ordervals):
        ov = np.ascontiguousarray(np.asarray(ordervals, dtype=np.float64).reshape(-1, self.M))
        lib().orc_marketmake(self._h, t, ov.ctypes.data, len(ov))

    def executeorder_sell(self, t, ordervals):
        ov = np.ascontiguousarray(np.asarray(ordervals, dtype=np.float64).reshape(-1, self.M))
        cash = np.zeros(len(ov))
        lib().orc_executeorder_sell(self._h, t, ov.ctypes.data, len(ov), cash.ctypes.data)
        return cash

    def executeorder_buy(self, t, ordervals):
        ov = np.ascontiguousarray(np.asarray(ordervals, dtype=np.float64).reshape(-1, self.M))
        shares = np.zeros_like(ov)
        lib().orc_executeorder_buy(self._h, t, ov.ctypes.data, len(ov), shares.ctypes.data)
        return shares

    def disburse_cash(self, divestments, cash):
        d = _i64(divestments).reshape(-1, 2)
        di, df = _i64(d[:, 0]), _i64(d[:, 1])
        c = np.ascontiguousarray(np.asarray(cash, dtype=np.float64))
        lib().orc_disburse_cash(self._h, di.ctypes.data, df.ctypes.data, len(d), c.ctypes.data)

    def disburse_shares(self, funds, shares):
        f = _i64(funds)
        s = np.ascontiguousarray(np.asarray(shares, dtype=np.float64).reshape(-1, self.M))
        lib().orc_disburse_shares(self._h, f.ctypes.data, len(f), s.ctypes.data)

    def respawn(self, t):
        buy = np.zeros((self.K, self.M))
        funds = np.zeros(self.K, dtype=np.int64)
        n = C.c_int64(0)
        rc = lib().orc_respawn(self._h, t, buy.ctypes.data, funds.ctypes.data, C.byref(n))
        return rc, buy[: n.value].copy(), funds[: n.value].copy()

    def reinvest(self, t, selector="probabilistic"):
        buy = np.zeros((self.N, self.M))
        funds = np.zeros(self.N, dtype=np.int64)
        n = C.c_int64(0)
        rc = lib().orc_reinvest(self._h, t, SEL[selector], buy.ctypes.data, funds.ctypes.data, C.byref(n))
        return rc, buy[: n.value].copy(), funds[: n.value].copy()

    def build_log(self):
        """Builds the initial liquidation log (functions.jl:448-456) without stepping."""
        t = self.params.perfwindow[-1] + 2
        lib().orc_modelrun(self._h, t, t - 1, 0, 0)


def fill_step_tape(seed, rep0, t, z_mkt, z_stock, u_review, threads=1):
    """One day's exogenous draws of replications rep0.. in the engine's host-tape layout (z_mkt[R], z_stock[R, M],
    u_review[R, N], C-contiguous float64, filled in place); `threads` host threads share the replications."""
    import threading
    R, M = z_stock.shape
    N = u_review.shape[1]
    f = lib().orc_fill_step_tape

    def work(lo, hi):
        if hi > lo:
            f(seed, rep0 + lo, hi - lo, t, M, N, z_mkt[lo:].ctypes.data, z_stock[lo:].ctypes.data, u_review[lo:].ctypes.data)

    if threads <= 1:
        work(0, R)
        return
    ths = [threading.Thread(target=work, args=((R * i) // threads, (R * (i + 1)) // threads)) for i in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()


# RNG building blocks -------------------------------------------------------------------------
def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return tuple(o)


def det_log(x):
    return lib().orc_det_log(x)


def det_norminv(u):
    return lib().orc_det_norminv(u)


def draw_normal(seed, rep, site, a, b, idx):
    return lib().orc_draw_normal(seed, rep, site, a, b, idx)


def draw_uniform(seed, rep, site, a, b, idx):
    return lib().orc_draw_uniform(seed, rep, site, a, b, idx)


def draw_discrete(seed, rep, site, a, b, idx, n):
    return lib().orc_draw_discrete(seed, rep, site, a, b, idx, n)


def dotw(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    return lib().orc_dotw(x.ctypes.data, 1, y.ctypes.data, 1, len(x))


def stylisedfacts(marketval, stocksval, stocksvolume, perfwindow_end, maxlag=10, returnspctile=5):
    """Func.calc_stylisedfacts (functions.jl:604-648), array outputs only; dict keyed by the reference's names."""
    mk = np.ascontiguousarray(marketval, dtype=np.float64)
    sv = np.asfortranarray(stocksval, dtype=np.float64)
    vo = np.asfortranarray(stocksvolume, dtype=np.float64)
    M, T = sv.shape
    out = dict(stockpacf=np.zeros((M, maxlag), order="F"), marketpacf=np.zeros(maxlag),
               stockvolacluster=np.zeros((M, maxlag), order="F"), marketvolacluster=np.zeros(maxlag),
               stockkurtoses=np.zeros(M), lossgainratios=np.zeros(M), corrs=np.zeros(M))
    rc = lib().orc_stylisedfacts(mk.ctypes.data, sv.ctypes.data, vo.ctypes.data, M, T, perfwindow_end, maxlag,
                                 float(returnspctile), *(out[k].ctypes.data for k in (
                                     "stockpacf", "marketpacf", "stockvolacluster", "marketvolacluster",
                                     "stockkurtoses", "lossgainratios", "corrs")))
    if rc != 0:
        raise ValueError(f"orc_stylisedfacts: {rc}")
    return out
