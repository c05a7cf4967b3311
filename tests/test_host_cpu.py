"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/fsb.h declares,
fails loudly without a CUDA device, and the host-side mirror of the reference's Params / Types API
behaves like the reference's own tests expect (test/params_test.jl, test/types_test.jl)."""
import ctypes as C
import importlib
import os
import re

import numpy as np
import pytest

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "fsb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fsb_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_what_the_binding_lists(fsb):
    assert set(_declared_functions()) == set(fsb._lib.EXPORTS)


def test_library_exports_every_declared_symbol(fsb):
    lib = fsb._lib.lib()
    missing = [n for n in _declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.fsb_version() == 200
    assert os.path.dirname(fsb._lib.so_path()).endswith(os.path.join("fundstababm.jl_b200", "csrc"))


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/fsb.h compiles as C99 (no C++ or CUDA types in the signatures) and a C
    translation unit can name every declared entry point."""
    import subprocess
    src = tmp_path / "abi.c"
    calls = "\n".join(f"    p[{i}] = (void *){n};" for i, n in enumerate(_declared_functions()))
    src.write_text('#include "fsb.h"\nint main(void) {\n    void *p[64];\n' + calls + "\n    return p[0] == 0;\n}\n")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Wno-pedantic", "-fsyntax-only",
                          "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_tape_key_layout(fsb):
    lib = fsb._lib.lib()
    assert lib.fsb_tape_key(3, 1234, 17, 5) == (3 << 56) | (1234 << 32) | (17 << 16) | 5


def test_big_shape_column_chunks_cover_every_horizon_once(fsb):
    """The big-shape pass (config 4) splits a day's distinct horizons into chunks that different CTAs take at the same time
    (DESIGN.md 4 'Big shapes'); the split is computed by the functions the kernel itself calls (fsb_debug_pass_chunks): every
    horizon exactly once and in order, chunk 0 at most 14 of them (next to the two price columns), later chunks at most 15
    (next to the day's price column), sizes within one of each other, as few chunks as the capacities allow."""
    import ctypes as C
    lib = fsb._lib.lib()
    first, count = (C.c_int32 * 16)(), (C.c_int32 * 16)()
    for ncols in range(0, 180):
        nch = lib.fsb_debug_pass_chunks(ncols, first, count, 16)
        assert nch >= 1
        assert 14 + 15 * (nch - 1) >= ncols and (nch == 1 or 14 + 15 * (nch - 2) < ncols)
        nxt = 0
        for ch in range(nch):
            assert first[ch] == nxt and 0 <= count[ch] <= (14 if ch == 0 else 15)
            nxt += count[ch]
        assert nxt == ncols
        assert max(count[:nch]) - min(count[:nch]) <= 1
    assert lib.fsb_debug_pass_chunks(500, first, count, 16) == -1


def test_no_cpu_fallback_without_a_device(fsb):
    """On a box without a GPU every compute entry point must fail, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(fsb._lib.FsbError) as ei:
        fsb.Engine(fsb.Params.default(), 1)
    assert ei.value.code == fsb._lib.ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fundstababm.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "fsb_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


# ---- Params: /root/reference/test/params_test.jl:3-24 ------------------------------------------
def test_params_default_sanity(fsb):
    p = fsb.Params.default()
    assert p.bign >= p.bigk >= 1 and p.bigm >= 1 and p.bigt >= 1
    assert p.marketstartval > 0 and p.marketvol >= 0
    assert 0 < p.perfwindow[0] <= p.perfwindow[-1] < p.bigt
    assert p.portfsizerange[-1] <= p.bigm
    assert p.impactrange[0] >= 0 and p.stockvolrange[0] >= 0
    assert (p.bigm, p.bign, p.bigt, p.bigk) == (250, 1000, 1350, 250)
    p.validate()


def test_params_ranges_match_julia(fsb):
    p = fsb.Params.default()
    imp = p.impactrange.values()
    assert len(imp) == 68 and imp[0] == 0.00075 and imp[1] == 0.00085 and imp[-1] == 0.00745
    vol = p.stockvolrange.values()
    assert vol == [0.001, 0.002, 0.003, 0.004, 0.005, 0.006, 0.007, 0.008, 0.009, 0.01]
    assert list(p.perfwindow) == list(range(1, 101)) and len(p.portfsizerange) == 91


def test_params_overrides_and_errors(fsb):
    p = fsb.Params.default(bigk=3, bign=4, bigm=5, bigt=6, perfwindow=(1, 3), portfsizerange=(1, 5))
    assert p(perfwindow=(3, 3)).perfwindow[0] == 3 and p.perfwindow[0] == 1
    with pytest.raises(TypeError):
        fsb.Params.default(nosuchfield=1)
    with pytest.raises(ValueError):
        fsb.Params.default(bigk=10, bign=5).validate()
    with pytest.raises(ValueError):
        fsb.Params.default(perfwindow=(1, 2000)).validate()


def test_params_marshal_to_the_pod(fsb):
    keep = []
    q = fsb._lib.make_params(fsb.Params.default(), keep)
    assert (q.bigm, q.bign, q.bigt, q.bigk) == (250, 1000, 1350, 250)
    assert q.n_impact == 68 and q.n_vol == 10 and q.impact_values[67] == 0.00745
    assert (q.perf_lo, q.perf_hi, q.cap_lo, q.cap_hi, q.psize_lo, q.psize_hi) == (1, 100, 50, 150, 10, 100)
    assert C.sizeof(fsb._lib.FsbParams) == 23 * 8 and C.sizeof(fsb._lib.FsbConfig) == 3 * 8 + 8 * 4


# ---- Types: /root/reference/test/types_test.jl:9-128 -------------------------------------------
def test_types_shapes_and_dtypes(fsb):
    p = fsb.Params.default(bigk=3, bign=4, bigm=5, bigt=6, perfwindow=(1, 3), portfsizerange=(1, 5))
    market, stocks, investors, funds = fsb.Types.allocate(p)
    K, M, N, T = 3, 5, 4, 6
    assert market.value.shape == (T,)
    assert stocks.value.shape == (M, T) and stocks.volume.shape == (M, T)
    assert stocks.beta.shape == stocks.vol.shape == stocks.impact.shape == (M,)
    assert investors.assets.shape == (N, K + 1) and investors.threshold.shape == (N,)
    assert investors.horizon.dtype == np.int64 and investors.horizon.shape == (N,)
    assert funds.holdings.shape == (K, M) and funds.stakes.shape == (K, N) and funds.value.shape == (K, T)
    for a in (stocks.value, stocks.volume, investors.assets, funds.holdings, funds.stakes, funds.value):
        assert a.dtype == np.float64 and a.flags["F_CONTIGUOUS"] and (a >= 0).all()
    # horizon converts to Int64 (test/types_test.jl:89); matrices must be 2-d
    inv = fsb.Types.RetailInvestor(np.zeros((N, K + 1)), np.array([1.0, 2.0, 3.0, 4.0]), np.zeros(N))
    assert inv.horizon.dtype == np.int64
    with pytest.raises(TypeError):
        fsb.Types.EquityFund(np.zeros(3), np.zeros((K, N)), np.zeros((K, T)))
    sell = fsb.Types.SellMarketOrder(np.zeros((2, M)), np.array([1, 2]))
    buy = fsb.Types.BuyMarketOrder(np.zeros((2, M)), np.array([1, 2]))
    assert sell.investors.dtype == np.int64 and buy.funds.dtype == np.int64


def test_selector_validation_q9(fsb):
    """Func.modelrun's own default "bestperformer" is not a valid selector (Q9, functions.jl:446 vs :344-356)."""
    p = fsb.Params.default(bigk=3, bign=4, bigm=5, bigt=6, perfwindow=(1, 3), portfsizerange=(1, 5))
    structs = fsb.Types.allocate(p)
    with pytest.raises(ValueError, match="valid fund selection method"):
        fsb.Func.modelrun(*structs, p)
    assert fsb.Func.VALID_SELECTORS == ("best", "goodenough", "probabilistic", "random")


def test_boundstest_reports_violations(fsb):
    p = fsb.Params.default(bigk=3, bign=4, bigm=5, bigt=6, perfwindow=(1, 3), portfsizerange=(1, 5))
    market, stocks, investors, funds = fsb.Types.allocate(p)
    assert fsb.Func.boundstest(market, stocks, investors, funds) == []
    funds.holdings[1, 2] = -1.0
    investors.threshold[0] = 2.0
    assert fsb.Func.boundstest(market, stocks, investors, funds) == ["funds.holdings >= 0", "1 >= investors.threshold >= -1"]


def test_calibration_grid_and_shards(fsb):
    pts = fsb.grid(thresholdstd=(0.01, 0.05), reviewprobability=(1 / 63, 1 / 21), impactscale=(1.0, 2.0))
    assert len(pts) == 8
    assert pts[1].impactrange.values()[0] == 2 * 0.00075 and pts[0].impactrange.values()[0] == 0.00075
    from fundstababm_jl_b200.calibrate import shard
    for total, world in ((48000, 8), (10, 3), (7, 7)):
        blocks = [shard(total, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == total
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1


def test_stats_decoding(fsb):
    L = fsb._lib
    T = 20
    row = np.zeros(L.STATS_HEADER + (T + 1) + L.STATS_LIQ_BINS + L.STATS_DD_BINS + L.STATS_FLOW_BINS, dtype=np.uint64)
    row[L.ST_REPS], row[L.ST_LIQUIDATIONS], row[L.ST_STEPS] = 5, 7, 90
    row[L.STATS_HEADER + 12] = 3                       # three failures on day 12
    row[L.STATS_HEADER + T + 1 + 2] = 4                # four replications with two liquidations
    st = fsb.engine.decode_stats(row, T)
    assert st["replications"] == 5 and st["liquidations"] == 7 and st["replication_steps"] == 90
    assert st["failtime_hist"][12] == 3 and st["liquidations_per_replication_hist"][2] == 4
    assert len(st["market_drawdown_hist"]) == L.STATS_DD_BINS and len(st["flow_hist"]) == L.STATS_FLOW_BINS


def test_recorder_file_round_trip(fsb, orc, tmp_path):
    """The file format of julia/record_tapes.jl (the recorder around the real reference, which this image cannot
    run): a recording produced from the oracle is written, read back and replayed through the oracle."""
    tf = importlib.import_module(fsb.__name__ + ".tapefile")
    rec, w = H.oracle_recording(fsb, orc, H.small_params(fsb), seed=21, rep=2, selector="probabilistic")
    path = tmp_path / "rec.fsbtape"
    tf.write_tape_file(path, **rec)
    back = tf.read_tape_file(path)
    assert back["selector"] == "probabilistic" and back["W"] == rec["W"] and back["reviewprobability"] == rec["reviewprobability"]
    for k in tf.STATE_FIELDS:
        assert H.eq(back["init"][k], rec["init"][k]) and H.eq(back["final"][k], rec["final"][k]), k
    assert back["events"] == [tuple(e) for e in rec["events"]]
    for k in ("z_mkt", "z_stock", "u_review"):
        assert H.eq(back["dense"][k], rec["dense"][k])
    # replay through the oracle: a fresh world from the recorded initial state under the recorded tapes
    p = H.small_params(fsb)
    w2 = orc.World(p, seed=999, rep=0)
    w2.load_state(back["init"])
    w2.set_dense_tape(back["dense"]["z_mkt"], back["dense"]["z_stock"], back["dense"]["u_review"])
    w2.set_event_tape([e for e in back["events"] if e[0] != fsb._lib.TAPE_CHOICE_U])
    assert w2.modelrun(selector="random") == 0  # (the taped choice indices override the selector)
    w2.finalise_fund_value(p.bigt)
    assert H.eq(w2.holdings, back["final"]["holdings"]) and H.eq(w2.stock_value, back["final"]["stock_value"])
    assert H.eq(w2.inv_assets, back["final"]["inv_assets"]) and H.eq(w2.stakes, back["final"]["stakes"])
    w.close(); w2.close()
