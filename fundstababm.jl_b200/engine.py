"""Engine -- thin object over the opaque ``fsb_engine*`` handle of include/fsb.h.

All arrays crossing this class are numpy, column-major (Fortran order) with the shapes of the
reference's structs (src/types.jl:15-50).  Ids exposed here are 0-based unless a method says
otherwise; the C ABI speaks Julia's 1-based ids where it fills Julia arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

LOG_COLUMNS = ("liquidated", "replacedby", "initialsize", "liquidity", "failtime", "ninvestors", "nstocks")


def _f(a, shape=None):
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


class Engine:
    def __init__(self, points, n_reps=1, *, rep_offset=0, reps_per_point=0, device=0, keep_history=False,
                 event_capacity=0, log_capacity=0, trace_capacity=0, block_threads=0,
                 blocks_per_sm=0):
        if not isinstance(points, (list, tuple)):
            points = [points]
        for p in points:
            p.validate()
        self.points = list(points)
        self.params = self.points[0]
        self._keep = []
        arr = (L.FsbParams * len(points))(*[L.make_params(p, self._keep) for p in points])
        cfg = L.FsbConfig(n_reps=n_reps, rep_offset=rep_offset, reps_per_point=reps_per_point, device=device,
                          keep_history=int(keep_history), event_capacity=event_capacity, log_capacity=log_capacity,
                          trace_capacity=trace_capacity,
                          block_threads=block_threads, blocks_per_sm=blocks_per_sm)
        self._h = C.c_void_p()
        L.check(L.lib().fsb_create(arr, len(points), C.byref(cfg), C.byref(self._h)))
        p = self.params
        self.K, self.M, self.N, self.T, self.W = p.bigk, p.bigm, p.bign, p.bigt, p.perfwindow[-1]
        self.n_reps, self.keep_history, self.device = n_reps, bool(keep_history), device
        self.t_first_default = self.W + 2

    # lifecycle ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            L.lib().fsb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    def next_day(self):
        """The day the next run must start with (days are simulated once and in order)."""
        return int(L.lib().fsb_next_day(self._h))

    def redzones(self):
        """(buffers, damaged) of an engine created under FSB_REDZONE=1; buffers == -1 without red zones."""
        n, bad = C.c_int64(), C.c_int64()
        L.check(L.lib().fsb_debug_redzones(self._h, C.byref(n), C.byref(bad)))
        return n.value, bad.value

    def device_bytes(self):
        return L.lib().fsb_device_bytes(self._h)

    # state in ----------------------------------------------------------------------------------
    def upload_state(self, rep, market, stocks, investors, funds):
        K, M, N, T = self.K, self.M, self.N, self.T
        a = dict(
            market=_f(market.value, (T,)), sv=_f(stocks.value, (M, T)), beta=_f(stocks.beta, (M,)),
            vol=_f(stocks.vol, (M,)), impact=_f(stocks.impact, (M,)), volume=_f(stocks.volume, (M, T)),
            assets=_f(investors.assets, (N, K + 1)),
            horizon=np.ascontiguousarray(investors.horizon, dtype=np.int64),
            thr=_f(investors.threshold, (N,)), hold=_f(funds.holdings, (K, M)), stakes=_f(funds.stakes, (K, N)),
            fv=_f(funds.value, (K, T)))
        L.check(L.lib().fsb_upload_state(
            self._h, rep, a["market"].ctypes.data, a["sv"].ctypes.data, a["beta"].ctypes.data, a["vol"].ctypes.data,
            a["impact"].ctypes.data, a["volume"].ctypes.data, a["assets"].ctypes.data, a["horizon"].ctypes.data,
            a["thr"].ctypes.data, a["hold"].ctypes.data, a["stakes"].ctypes.data, a["fv"].ctypes.data))

    def init_device(self, seed):
        L.check(L.lib().fsb_init_device(self._h, int(seed)))

    def set_seed(self, seed):
        L.check(L.lib().fsb_set_seed(self._h, int(seed)))

    # tapes / trace -----------------------------------------------------------------------------
    def set_dense_tape(self, rep, z_mkt=None, z_stock=None, u_review=None):
        ptrs = []
        for a, shape in ((z_mkt, (self.T,)), (z_stock, (self.M, self.T)), (u_review, (self.N, self.T))):
            if a is None:
                ptrs.append(None)
            else:
                a = _f(a, shape)
                self._keep.append(a)
                ptrs.append(a.ctypes.data)
        L.check(L.lib().fsb_set_dense_tape(self._h, rep, *ptrs))

    def set_event_tape(self, rep, recs):
        """recs: iterable of (site, t, a, b, value) with 0-based ids; sorted here."""
        lib = L.lib()
        arr = np.zeros(len(recs), dtype=L.TAPE_DTYPE)
        for i, (site, t, a, b, v) in enumerate(recs):
            arr[i] = (lib.fsb_tape_key(site, t, a, b), v)
        arr.sort(order="key")
        L.check(lib.fsb_set_event_tape(self._h, rep, arr.ctypes.data if len(arr) else None, len(arr)))

    def set_trace(self, rep, on=True):
        L.check(L.lib().fsb_set_trace(self._h, rep, int(on)))

    # hot path ----------------------------------------------------------------------------------
    def run(self, t_first=None, t_last=None, selector="probabilistic", flags=0):
        if selector not in L.SEL:
            raise ValueError("Please specify a valid fund selection method in reinvest!")
        t_first = self.next_day() if t_first is None else t_first
        t_last = self.T if t_last is None else t_last
        L.check(L.lib().fsb_run(self._h, t_first, t_last, L.SEL[selector], flags))

    def run_async(self, t_first=None, t_last=None, selector="probabilistic", flags=0):
        t_first = self.next_day() if t_first is None else t_first
        t_last = self.T if t_last is None else t_last
        L.check(L.lib().fsb_run_async(self._h, t_first, t_last, L.SEL[selector], flags))

    def run_step_hosttape(self, t, z_mkt, z_stock, u_review, nav_open_out=None, selector="probabilistic", flags=0):
        """One step for the whole batch with this step's shocks taken from host arrays
        z_mkt[R], z_stock[R, M], u_review[R, N] (C-contiguous; pinned memory recommended)."""
        R = self.n_reps
        for a, shape in ((z_mkt, (R,)), (z_stock, (R, self.M)), (u_review, (R, self.N))):
            if a.dtype != np.float64 or a.shape != shape or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"host tape must be C-contiguous float64 of shape {shape}")
        out = None
        if nav_open_out is not None:
            if nav_open_out.shape != (R, self.K) or not nav_open_out.flags["C_CONTIGUOUS"]:
                raise ValueError("nav_open_out must be C-contiguous (R, K)")
            out = nav_open_out.ctypes.data
        L.check(L.lib().fsb_run_step_hosttape(self._h, t, L.SEL[selector], flags, z_mkt.ctypes.data,
                                              z_stock.ctypes.data, u_review.ctypes.data, out))

    def sync(self):
        L.check(L.lib().fsb_sync(self._h))

    def last_run_timing(self):
        tot, stp, n = C.c_double(), C.c_double(), C.c_int64()
        L.check(L.lib().fsb_last_run_timing(self._h, C.byref(tot), C.byref(stp), C.byref(n)))
        return tot.value, stp.value, n.value

    def last_pass_launches(self):
        return int(L.lib().fsb_last_pass_launches(self._h))

    def rep_status(self):
        out = np.zeros(self.n_reps, dtype=np.int32)
        L.check(L.lib().fsb_rep_status(self._h, out.ctypes.data))
        return out

    # state out ---------------------------------------------------------------------------------
    def download_state(self, rep):
        K, M, N, T = self.K, self.M, self.N, self.T
        o = dict(
            market=np.zeros(T), stock_value=np.zeros((M, T), order="F"), beta=np.zeros(M), vol=np.zeros(M),
            impact=np.zeros(M), volume=np.zeros((M, T), order="F"), inv_assets=np.zeros((N, K + 1), order="F"),
            horizon=np.zeros(N, dtype=np.int64), threshold=np.zeros(N), holdings=np.zeros((K, M), order="F"),
            stakes=np.zeros((K, N), order="F"), fund_value=np.zeros((K, T), order="F"))
        L.check(L.lib().fsb_download_state(self._h, rep, *[o[k].ctypes.data for k in (
            "market", "stock_value", "beta", "vol", "impact", "volume", "inv_assets", "horizon", "threshold",
            "holdings", "stakes", "fund_value")]))
        return o

    def download_compact(self, rep):
        K, M, N = self.K, self.M, self.N
        o = dict(holdings=np.zeros((K, M), order="F"), price_now=np.zeros(M), inv_fund=np.zeros(N, dtype=np.int64),
                 inv_amount=np.zeros(N), inv_stake=np.zeros(N), nav_open=np.zeros(K), nav_close=np.zeros(K),
                 market_now=np.zeros(1))
        L.check(L.lib().fsb_download_compact(self._h, rep, *[o[k].ctypes.data for k in (
            "holdings", "price_now", "inv_fund", "inv_amount", "inv_stake", "nav_open", "nav_close", "market_now")]))
        o["inv_fund"] -= 1  # 0-based; K == cash
        return o

    def download_log(self, rep):
        n = C.c_int64()
        L.check(L.lib().fsb_download_log(self._h, rep, C.byref(n), *([None] * 7)))
        n = n.value
        cols = {name: np.zeros(n, dtype=np.float64 if name in ("initialsize", "liquidity") else np.int64)
                for name in LOG_COLUMNS}
        nn = C.c_int64()
        L.check(L.lib().fsb_download_log(self._h, rep, C.byref(nn), *[cols[k].ctypes.data for k in LOG_COLUMNS]))
        return cols

    def download_trace(self, rep):
        K, M, T = self.K, self.M, self.T
        n = C.c_int64()
        nav = np.zeros((K, T), order="F")
        ps = np.zeros((M, T), order="F")
        pr = np.zeros((M, T), order="F")
        L.check(L.lib().fsb_download_trace(self._h, rep, nav.ctypes.data, ps.ctypes.data, pr.ctypes.data, None, C.byref(n)))
        recs = np.zeros(max(n.value, 1), dtype=L.TRACE_DTYPE)
        L.check(L.lib().fsb_download_trace(self._h, rep, None, None, None, recs.ctypes.data, C.byref(n)))
        return dict(nav_open=nav, p_sell=ps, p_resp=pr, events=recs[: n.value])

    def stylised_facts(self, rep_first=0, n_reps=None, maxlag=10, returnspctile=5, out=None):
        """Func.calc_stylisedfacts (functions.jl:604-648) on the device; arrays keyed by the reference's names,
        leading axis = replication.  ``out``: the dict a previous call with the same shape returned -- its arrays
        are filled again (a sweep loop saves the page faults of 46 KB of fresh result memory per replication)."""
        n = self.n_reps - rep_first if n_reps is None else n_reps
        M = self.M
        if out is None:
            o = dict(stockpacf=np.zeros((n, maxlag, M)), marketpacf=np.zeros((n, maxlag)),
                     stockvolacluster=np.zeros((n, maxlag, M)), marketvolacluster=np.zeros((n, maxlag)),
                     stockkurtoses=np.zeros((n, M)), lossgainratios=np.zeros((n, M)), corrs=np.zeros((n, M)))
        else:
            o = dict(out)
            for k in ("stockpacf", "stockvolacluster"):
                o[k] = o[k].transpose(0, 2, 1)  # back to the ABI's [rep][lag][stock]
            if o["stockpacf"].shape != (n, maxlag, M) or not all(v.flags["C_CONTIGUOUS"] for v in o.values()):
                raise ValueError("out does not come from a stylised_facts call of the same shape")
        L.check(L.lib().fsb_stylised_facts(self._h, rep_first, n, maxlag, float(returnspctile), *[o[k].ctypes.data for k in (
            "stockpacf", "marketpacf", "stockvolacluster", "marketvolacluster", "stockkurtoses", "lossgainratios",
            "corrs")]))
        for k in ("stockpacf", "stockvolacluster"):
            o[k] = o[k].transpose(0, 2, 1)  # [rep][stock][lag]: the reference's bigm x nlags matrix per replication
        return o

    def last_facts_ms(self):
        ms = C.c_double()
        L.check(L.lib().fsb_last_facts_timing(self._h, C.byref(ms)))
        return ms.value

    # statistics --------------------------------------------------------------------------------
    def comm_init_rank(self, n_ranks, rank, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        L.check(L.lib().fsb_comm_init_rank(self._h, n_ranks, rank, buf))

    def stats_raw(self, local=False):
        """Integer ensemble statistics per sweep point: summed over the ranks of the communicator (one ncclAllReduce), or
        this engine's replications only (``local=True``, no collective)."""
        wpp = L.lib().fsb_stats_words_per_point(self._h)
        out = np.zeros((len(self.points), wpp), dtype=np.uint64)
        L.check((L.lib().fsb_stats_local if local else L.lib().fsb_stats)(self._h, out.ctypes.data))
        return out

    def stats(self):
        return [decode_stats(row, self.T) for row in self.stats_raw()]

    # debug -------------------------------------------------------------------------------------
    def debug_draws(self, seed, rep, site, a, b, idx0, n, kind, n_discrete=0):
        out = np.zeros(n)
        L.check(L.lib().fsb_debug_draws(self._h, seed, rep, site, a, b, idx0, n, kind, n_discrete, out.ctypes.data))
        return out

    def debug_dot(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        out = C.c_double()
        L.check(L.lib().fsb_debug_dot(self._h, x.ctypes.data, y.ctypes.data, len(x), C.byref(out)))
        return out.value


def comm_init_all(engines):
    """ONE process driving several devices (what a Julia task calling the library does): one NCCL communicator over the
    engines' devices (ncclCommInitAll)."""
    arr = (C.c_void_p * len(engines))(*[e.handle for e in engines])
    L.check(L.lib().fsb_comm_init_all(arr, len(engines)))


def stats_all(engines):
    """The per-point ensemble statistics summed over the engines of one process: a grouped ncclAllReduce, one call per
    device (fsb_stats_all).  Returns the decoded statistics, one dict per sweep point."""
    e0 = engines[0]
    wpp = L.lib().fsb_stats_words_per_point(e0.handle)
    out = np.zeros((len(e0.points), wpp), dtype=np.uint64)
    arr = (C.c_void_p * len(engines))(*[e.handle for e in engines])
    L.check(L.lib().fsb_stats_all(arr, len(engines), out.ctypes.data))
    return out


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    L.check(L.lib().fsb_comm_unique_id(buf))
    return bytes(buf)


def decode_stats(row, T):
    h = L.STATS_HEADER
    o = dict(
        replications=int(row[L.ST_REPS]), liquidations=int(row[L.ST_LIQUIDATIONS]),
        funds_ever=int(row[L.ST_FUNDS_EVER]), floor_hits=int(row[L.ST_FLOOR_HITS]),
        divestments=int(row[L.ST_DIVESTMENTS]), reviews=int(row[L.ST_REVIEWS]),
        reinvestments=int(row[L.ST_REINVESTMENTS]), error_replications=int(row[L.ST_ERR_REPS]),
        replication_steps=int(row[L.ST_STEPS]), nan_replications=int(row[L.ST_NAN_REPS]))
    o["failtime_hist"] = row[h: h + T + 1].astype(np.int64)
    h += T + 1
    o["liquidations_per_replication_hist"] = row[h: h + L.STATS_LIQ_BINS].astype(np.int64)
    h += L.STATS_LIQ_BINS
    o["market_drawdown_hist"] = row[h: h + L.STATS_DD_BINS].astype(np.int64)
    h += L.STATS_DD_BINS
    o["flow_hist"] = row[h: h + L.STATS_FLOW_BINS].astype(np.int64)
    return o
