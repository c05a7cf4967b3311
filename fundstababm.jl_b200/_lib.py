"""ctypes binding of csrc/libfsb_b200.so (include/fsb.h).  No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None

# status codes (include/fsb.h)
OK = 0
ERR_BAD_PARAMS, ERR_BAD_SELECTOR, ERR_CUDA, ERR_NCCL, ERR_BAD_STATE, ERR_REPLICATION, ERR_NO_HISTORY, ERR_RANGE = \
    -1, -2, -3, -4, -5, -6, -7, -8
REP_Q6_NO_ANCHOR, REP_TAPE_MISSING, REP_TAPE_INVALID, REP_EVENT_OVERFLOW, REP_LOG_OVERFLOW, REP_TRACE_OVERFLOW = \
    1, 2, 4, 8, 16, 32
SEL = {"best": 0, "goodenough": 1, "probabilistic": 2, "random": 3}
RUN_NO_VOLUME_QUIRK = 1
EV_REVIEWER, EV_DIVEST, EV_CASHOUT, EV_FLOOR, EV_SPAWN, EV_SPAWN_N, EV_CHOICE, EV_PICK = range(1, 9)
TAPE_ANCHOR, TAPE_RPSIZE, TAPE_RPPICK, TAPE_CHOICE_U, TAPE_CHOICE_IDX = range(1, 6)
STATS_HEADER, STATS_LIQ_BINS, STATS_DD_BINS, STATS_FLOW_BINS = 16, 256, 64, 64
ST_REPS, ST_LIQUIDATIONS, ST_FUNDS_EVER, ST_FLOOR_HITS, ST_DIVESTMENTS, ST_REVIEWS, ST_REINVESTMENTS, \
    ST_ERR_REPS, ST_STEPS, ST_NAN_REPS = range(10)

TRACE_DTYPE = np.dtype([("t", "<i4"), ("kind", "<i4"), ("a", "<i4"), ("b", "<i4"), ("v", "<f8")])
TAPE_DTYPE = np.dtype([("key", "<u8"), ("value", "<f8")])

EXPORTS = (
    "fsb_version", "fsb_last_error", "fsb_create", "fsb_destroy", "fsb_upload_state", "fsb_init_device", "fsb_set_seed",
    "fsb_set_dense_tape", "fsb_set_event_tape", "fsb_tape_key", "fsb_set_trace", "fsb_run", "fsb_run_async", "fsb_run_step_hosttape",
    "fsb_sync", "fsb_download_state", "fsb_download_compact", "fsb_download_log", "fsb_download_trace",
    "fsb_rep_status", "fsb_stylised_facts", "fsb_comm_unique_id", "fsb_comm_init_rank", "fsb_comm_init_all", "fsb_stats", "fsb_stats_local", "fsb_stats_all",
    "fsb_stats_words_per_point", "fsb_last_run_timing", "fsb_last_pass_launches", "fsb_last_facts_timing", "fsb_device_bytes", "fsb_debug_redzones", "fsb_next_day", "fsb_debug_draws", "fsb_debug_dot",
    "fsb_debug_phase_cycles", "fsb_debug_pass_chunks",
)


class FsbParams(C.Structure):
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


class FsbConfig(C.Structure):
    _fields_ = [
        ("n_reps", C.c_int64), ("rep_offset", C.c_int64), ("reps_per_point", C.c_int64),
        ("device", C.c_int32), ("keep_history", C.c_int32), ("event_capacity", C.c_int32),
        ("log_capacity", C.c_int32), ("trace_capacity", C.c_int32),
        ("block_threads", C.c_int32), ("blocks_per_sm", C.c_int32),
    ]


class FsbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libfsb_b200 error {code}: {message}")
        self.code = code


def so_path() -> str:
    # FSB_LIB: an alternative build of the same library (kernel experiments under scripts/), never a fallback
    return os.environ.get("FSB_LIB") or _build.SO


def lib():
    """Loads the CUDA extension; fails loudly if it is missing (there is no CPU fallback)."""
    global _LIB
    if _LIB is None:
        path = so_path()
        if not os.path.exists(path):
            raise FsbError(ERR_CUDA, f"{path} is not built; run `python __graft_entry__.py build` "
                                     "(fundstababm.jl_b200 has no CPU fallback)")
        L = C.CDLL(path)
        vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
        P = C.POINTER
        L.fsb_version.restype = i32
        L.fsb_last_error.restype = C.c_char_p
        L.fsb_create.argtypes = [P(FsbParams), i32, P(FsbConfig), P(vp)]
        L.fsb_destroy.argtypes = [vp]
        L.fsb_upload_state.argtypes = [vp, i64] + [vp] * 12
        L.fsb_init_device.argtypes = [vp, u64]
        L.fsb_set_seed.argtypes = [vp, u64]
        L.fsb_set_dense_tape.argtypes = [vp, i64, vp, vp, vp]
        L.fsb_set_event_tape.argtypes = [vp, i64, vp, i64]
        L.fsb_tape_key.restype = u64
        L.fsb_tape_key.argtypes = [i32, i64, i64, i64]
        L.fsb_set_trace.argtypes = [vp, i64, i32]
        L.fsb_run.argtypes = [vp, i64, i64, i32, i32]
        L.fsb_run_async.argtypes = [vp, i64, i64, i32, i32]
        L.fsb_sync.argtypes = [vp]
        L.fsb_run_step_hosttape.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp]
        L.fsb_download_state.argtypes = [vp, i64] + [vp] * 12
        L.fsb_download_compact.argtypes = [vp, i64] + [vp] * 8
        L.fsb_download_log.argtypes = [vp, i64, P(i64)] + [vp] * 7
        L.fsb_download_trace.argtypes = [vp, i64, vp, vp, vp, vp, P(i64)]
        L.fsb_rep_status.argtypes = [vp, vp]
        L.fsb_stylised_facts.argtypes = [vp, i64, i64, i32, dbl] + [vp] * 7
        L.fsb_comm_unique_id.argtypes = [vp]
        L.fsb_comm_init_rank.argtypes = [vp, i32, i32, vp]
        L.fsb_comm_init_all.argtypes = [P(vp), i32]
        L.fsb_stats.argtypes = [vp, vp]
        L.fsb_stats_local.argtypes = [vp, vp]
        L.fsb_stats_all.argtypes = [P(vp), i32, vp]
        L.fsb_stats_words_per_point.restype = i64
        L.fsb_stats_words_per_point.argtypes = [vp]
        L.fsb_last_run_timing.argtypes = [vp, P(dbl), P(dbl), P(i64)]
        L.fsb_last_facts_timing.argtypes = [vp, P(dbl)]
        L.fsb_last_pass_launches.restype = i64
        L.fsb_last_pass_launches.argtypes = [vp]
        L.fsb_device_bytes.restype = i64
        L.fsb_device_bytes.argtypes = [vp]
        L.fsb_debug_redzones.argtypes = [vp, P(i64), P(i64)]
        L.fsb_next_day.restype = i64
        L.fsb_next_day.argtypes = [vp]
        L.fsb_debug_draws.argtypes = [vp, u64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, i32, i32,
                                      C.c_uint32, vp]
        L.fsb_debug_dot.argtypes = [vp, vp, vp, i64, vp]
        L.fsb_debug_phase_cycles.argtypes = [vp, vp, i32]
        L.fsb_debug_pass_chunks.argtypes = [i32, vp, vp, i32]
        _LIB = L
    return _LIB


def check(code):
    if code != OK:
        raise FsbError(code, lib().fsb_last_error().decode("utf-8", "replace"))


def range_values(r):
    return np.ascontiguousarray(r.values() if hasattr(r, "values") else list(r), dtype=np.float64)


def make_params(p, keep):
    """ParamSet -> fsb_params; ``keep`` collects the arrays that must outlive the call."""
    imp, vol = range_values(p.impactrange), range_values(p.stockvolrange)
    keep += [imp, vol]
    return FsbParams(
        bigm=p.bigm, bign=p.bign, bigt=p.bigt, bigk=p.bigk,
        marketstartval=p.marketstartval, drift=p.drift, marketvol=p.marketvol,
        impact_values=imp.ctypes.data_as(C.POINTER(C.c_double)), n_impact=len(imp),
        betamean=p.betamean, betastd=p.betastd, stockstartval=p.stockstartval,
        vol_values=vol.ctypes.data_as(C.POINTER(C.c_double)), n_vol=len(vol),
        reviewprobability=p.reviewprobability,
        perf_lo=p.perfwindow[0], perf_hi=p.perfwindow[-1],
        cap_lo=p.invcaprange[0], cap_hi=p.invcaprange[1],
        thresholdmean=p.thresholdmean, thresholdstd=p.thresholdstd,
        psize_lo=p.portfsizerange[0], psize_hi=p.portfsizerange[-1])
