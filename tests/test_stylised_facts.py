"""Func.calc_stylisedfacts (src/functions.jl:604-648).

CPU part: the oracle's restatement against independent numpy / scipy formulations of the published
algorithms (the reference has no tests for these functions -- TODOs at functions.jl:547,554,560,572 --
and Julia is absent, so this is what pins the oracle; "parity unpinned" against Julia itself).
GPU part: the device kernel, through the C ABI, against the oracle.  Tolerances: the kernel derives the
regression sums of every lag from 11 full-range lagged sums, so PACF coefficients agree to rounding
(1e-9 absolute on values of order 0.1), kurtosis 1e-10 relative, correlations 1e-11 absolute; the
loss/gain ratio is a quotient of counts on bit-identical returns and must be equal.
"""
import numpy as np
import pytest

import helpers as H

SEED = 0xFAC75


def synthetic_history(rng, M, T):
    mk = 100 * np.cumprod(1 + 0.01 * rng.standard_normal(T))
    sv = np.asfortranarray(100 * np.cumprod(1 + 0.01 * rng.standard_t(4, (M, T)), axis=1))
    vo = np.asfortranarray(-np.abs(rng.standard_normal((M, T))) * (rng.random((M, T)) < 0.3))
    return mk, sv, vo


def np_pacf(x, l):
    n = len(x)
    X = np.column_stack([np.ones(n - l)] + [x[l - a:n - a] for a in range(1, l + 1)])
    return np.linalg.lstsq(X, x[l:], rcond=None)[0][-1]


@pytest.mark.parametrize("M,T,W,pct", [(7, 200, 20, 5), (3, 64, 5, 50), (5, 90, 1, 0), (4, 120, 10, 100)])
def test_oracle_against_numpy(orc, M, T, W, pct):
    from scipy.stats import kurtosis
    mk, sv, vo = synthetic_history(np.random.default_rng(M * 1000 + T), M, T)
    o = orc.stylisedfacts(mk, sv, vo, W, maxlag=10, returnspctile=pct)

    def rets(v):
        r = (v[W + 1:] - v[W:-1]) / v[W:-1]
        return r - r.mean()
    for s in range(M):
        x = rets(sv[s])
        for l in range(1, 11):
            assert abs(np_pacf(x, l) - o["stockpacf"][s, l - 1]) < 1e-9
            assert abs(np_pacf(x * x, l) - o["stockvolacluster"][s, l - 1]) < 1e-9
        assert abs(kurtosis(x) - o["stockkurtoses"][s]) < 1e-10 * max(1.0, abs(kurtosis(x)))
        z = x[W:]
        q = np.percentile(abs(z), pct)
        with np.errstate(divide="ignore", invalid="ignore"):
            lg = np.float64((z < -q).sum()) / np.float64((z > q).sum())
        assert H.eq(lg, o["lossgainratios"][s])
        assert abs(np.corrcoef(vo[s, W + 1:], abs(x))[0, 1] - o["corrs"][s]) < 1e-12
    x = rets(mk)
    for l in range(1, 11):
        assert abs(np_pacf(x, l) - o["marketpacf"][l - 1]) < 1e-9
        assert abs(np_pacf(x * x, l) - o["marketvolacluster"][l - 1]) < 1e-9


def test_oracle_rejects_short_series(orc):
    mk, sv, vo = synthetic_history(np.random.default_rng(0), 2, 30)
    with pytest.raises(ValueError):
        orc.stylisedfacts(mk, sv, vo, 10, maxlag=10)   # 19 returns for 10 lags


def fake_facts(rng, R, M, L, nan_reps=()):
    f = dict(stockpacf=rng.standard_normal((R, M, L)), stockvolacluster=rng.standard_normal((R, M, L)),
             marketpacf=rng.standard_normal((R, L)), marketvolacluster=rng.standard_normal((R, L)),
             stockkurtoses=rng.standard_normal((R, M)) + 3, lossgainratios=rng.random((R, M)) + 0.5,
             corrs=rng.random((R, M)) - 0.5)
    for r in nan_reps:
        for v in f.values():
            v[r] = np.nan
    f["corrs"][0, 1] = np.nan          # a stock that never traded
    f["lossgainratios"][1, 0] = np.inf   # no up-move beyond the cutoff
    return f


def test_calibrate_result_averaging(fsb):
    """Per-point means over the finite entries; sharding the replications over ranks only reorders a sum."""
    cal = fsb._calibrate
    R, M, L, reps = 12, 5, 4, 4
    f = fake_facts(np.random.default_rng(2), R, M, L, nan_reps=(5,))
    pts = np.arange(R) // reps
    whole = cal.facts_moments(f, pts, 3)
    parts = sum(cal.facts_moments({k: v[a:b] for k, v in f.items()}, pts[a:b], 3) for a, b in ((0, 5), (5, 7), (7, 12)))
    np.testing.assert_allclose(parts, whole, rtol=1e-13)
    means = cal.facts_means(whole)
    assert len(means) == 3
    k = f["stockkurtoses"]
    assert abs(means[0]["stockkurtoses"] - k[:4].mean()) < 1e-13 and means[0]["stockkurtoses_n"] == 4 * M
    assert means[1]["stockkurtoses_n"] == 3 * M                       # replication 5 is a NaN trajectory
    assert abs(means[1]["stockkurtoses"] - np.nanmean(k[4:8])) < 1e-13
    assert means[0]["corrs_n"] == 4 * M - 1 and means[0]["lossgainratios_n"] == 4 * M - 1
    np.testing.assert_allclose(means[2]["marketpacf"], f["marketpacf"][8:].mean(axis=0), rtol=1e-13)
    np.testing.assert_allclose(means[1]["stockvolacluster"], np.nanmean(f["stockvolacluster"][4:8], axis=(0, 1)), rtol=1e-13)
    with pytest.raises(ValueError):
        fsb.calibrate([fsb.Params.default()], 1, evaluation="kurtosis")


def assert_facts_close(got, ref):
    for k in ("stockpacf", "stockvolacluster", "marketpacf", "marketvolacluster"):
        g = got[k][0] if isinstance(got[k], tuple) else got[k]
        np.testing.assert_allclose(g, ref[k], rtol=0, atol=1e-9, err_msg=k)
    np.testing.assert_allclose(got["stockkurtoses"], ref["stockkurtoses"], rtol=1e-10, err_msg="kurtoses")
    assert H.eq(got["lossgainratios"], ref["lossgainratios"]), (got["lossgainratios"], ref["lossgainratios"])
    np.testing.assert_allclose(got["corrs"], ref["corrs"], rtol=0, atol=1e-11, err_msg="corrs")


@pytest.mark.gpu
@pytest.mark.parametrize("M,T,W,pct,maxlag", [(7, 200, 20, 5, 10), (300, 160, 5, 50, 3), (5, 90, 1, 0, 10),
                                                (4, 120, 10, 100, 7), (1100, 70, 3, 30, 2)])
def test_device_facts_of_uploaded_histories(fsb, orc, M, T, W, pct, maxlag):
    """Func.calc_stylisedfacts on arrays: heavy-tailed synthetic histories, every stock trades."""
    mk, sv, vo = synthetic_history(np.random.default_rng(M + T), M, T)
    params = H.small_params(fsb, bigm=M, bigt=T, perfwindow=(1, W), portfsizerange=(1, min(M, 3)))
    got = fsb.Func.calc_stylisedfacts(mk, sv, vo, params, maxlag=maxlag, returnspctile=pct)
    ref = orc.stylisedfacts(mk, sv, vo, W, maxlag=maxlag, returnspctile=pct)
    assert_facts_close(got, ref)
    assert got["stockpacf"][0].shape == (M, maxlag)
    assert got["stockpacf"][1] == 1.96 / np.sqrt(M) and got["marketpacf"][1] == 1.96 / np.sqrt(T - W - 1)
    for name, mu0, key in (("zScoreKurtoses", 0.0, "stockkurtoses"), ("zScoreLossGain", 1.0, "lossgainratios"),
                           ("zScoreVolumeVola", 0.0, "corrs")):
        v = ref[key]
        with np.errstate(invalid="ignore"):
            z = (v.mean() - mu0) / (v.std(ddof=1) / np.sqrt(len(v)))
        np.testing.assert_allclose(got[name], z, rtol=1e-8, equal_nan=True)


@pytest.mark.gpu
def test_device_facts_after_a_run(fsb, orc):
    """The natural flow (src/runmodel.jl:36,47): simulate, then the facts from the histories still in HBM --
    several replications in one launch, against the oracle run of each."""
    params = H.small_params(fsb, bigt=120, bigm=12, portfsizerange=(2, 8))
    R = 5
    with fsb.Engine(params, R, keep_history=True) as eng:
        eng.init_device(SEED)
        eng.run(selector="probabilistic")
        got = eng.stylised_facts(maxlag=10, returnspctile=5)
        sub = eng.stylised_facts(3, 2, maxlag=4, returnspctile=20)
        again = eng.stylised_facts(maxlag=10, returnspctile=5, out={k: np.full_like(v, 7.0) for k, v in got.items()})
        assert all(H.eq(again[k], got[k]) for k in got)            # result arrays of a previous call can be reused
        with pytest.raises(ValueError):
            eng.stylised_facts(maxlag=9, out=got)
        for r in range(R):
            w = H.oracle_world(orc, params, SEED, r, trace=False)
            w.modelrun(selector="probabilistic")
            st = eng.download_state(r)
            assert H.eq(st["stock_value"], w.stock_value) and H.eq(st["volume"], w.volume)
            ref = orc.stylisedfacts(w.market, w.stock_value, w.volume, params.perfwindow[-1])
            with np.errstate(invalid="ignore"):
                assert_facts_close({k: v[r] for k, v in got.items()}, ref)
            if r >= 3:
                ref = orc.stylisedfacts(w.market, w.stock_value, w.volume, params.perfwindow[-1], maxlag=4, returnspctile=20)
                assert_facts_close({k: v[r - 3] for k, v in sub.items()}, ref)


@pytest.mark.gpu
def test_device_facts_default_shape_runmodel(fsb, orc):
    """runmodel() at the default Params shape returns the stylised facts of its own run (runmodel.jl:47)."""
    res = fsb.runmodel(fsb.Params.default(), seed=SEED, verbose=False)
    W = res.stocks.value.shape[1] - 1249 - 1
    ref = orc.stylisedfacts(res.market.value, res.stocks.value, res.stocks.volume, W)
    assert_facts_close(res.stylefacts, ref)
    assert np.isfinite(res.stylefacts["stockkurtoses"]).all()


@pytest.mark.gpu
def test_device_facts_errors(fsb):
    params = H.small_params(fsb)
    with fsb.Engine(params, 1, keep_history=False) as eng:
        eng.init_device(1)
        with pytest.raises(fsb._lib.FsbError) as ei:
            eng.stylised_facts()
        assert ei.value.code == fsb._lib.ERR_NO_HISTORY
    with fsb.Engine(params, 1, keep_history=True) as eng:
        eng.init_device(1)
        for kw in (dict(maxlag=0), dict(maxlag=11), dict(returnspctile=101), dict(rep_first=1)):
            with pytest.raises(fsb._lib.FsbError):
                eng.stylised_facts(**kw)
    short = H.small_params(fsb, bigt=25)   # 17 returns: too few for 10 lags
    with fsb.Engine(short, 1, keep_history=True) as eng:
        eng.init_device(1)
        with pytest.raises(fsb._lib.FsbError):
            eng.stylised_facts()
        eng.stylised_facts(maxlag=3)


@pytest.mark.gpu
def test_calibrate_evaluates_stylised_facts(fsb):
    """configs[4] in miniature with the "Evaluation" stage of src/calibrate.jl:14-16 on the device."""
    base = H.small_params(fsb, bigt=100)
    pts = fsb.grid(base, thresholdstd=(0.02, 0.1))
    reps = 4
    stats = fsb.calibrate(pts, reps, seed=SEED, evaluation="stylisedfacts", maxlag=5)
    plain = fsb.calibrate(pts, reps, seed=SEED)
    for i in range(len(pts)):
        for k, v in plain[i].items():
            assert H.eq(stats[i][k], v), k                  # keeping the histories changes no outcome
        with fsb.Engine(pts, reps, rep_offset=i * reps, reps_per_point=reps, keep_history=True) as eng:
            eng.init_device(SEED)
            eng.run()
            f = eng.stylised_facts(maxlag=5)
        sf = stats[i]["stylisedfacts"]
        np.testing.assert_allclose(sf["stockkurtoses"], f["stockkurtoses"][np.isfinite(f["stockkurtoses"])].mean(), rtol=1e-12)
        np.testing.assert_allclose(sf["marketpacf"], f["marketpacf"].mean(axis=0), rtol=1e-12)
        np.testing.assert_allclose(sf["stockpacf"], np.nanmean(f["stockpacf"], axis=(0, 1)), rtol=1e-12)
        assert sf["marketpacf"].shape == (5,)


@pytest.mark.gpu
def test_calibrate_history_batches_are_invisible(fsb):
    """Stylised facts for ensembles that only fit as ring engines: calibrate evaluates the rank's
    replications in history-keeping batches; the batch size changes neither the integer statistics nor (beyond the order
    of a float sum) the facts, and both equal one history-keeping engine over all replications."""
    base = fsb.Params.default(bigt=260)
    pts = fsb.grid(base, thresholdstd=(0.03, 0.08))
    reps = 10
    one = fsb.calibrate(pts, reps, seed=SEED, evaluation="stylisedfacts", history_batch=2 * reps)   # a single batch
    for hb in (3, 7):
        got = fsb.calibrate(pts, reps, seed=SEED, evaluation="stylisedfacts", history_batch=hb)
        for a, b in zip(got, one):
            for k in a:
                if k == "stylisedfacts":
                    for kk in a[k]:
                        np.testing.assert_allclose(a[k][kk], b[k][kk], rtol=1e-12, atol=1e-15, equal_nan=True, err_msg=kk)
                else:
                    assert H.eq(a[k], b[k]), k
    plain = fsb.calibrate(pts, reps, seed=SEED)   # ring engine, the library's own allreduce path
    for a, b in zip(one, plain):
        for k in b:
            assert H.eq(a[k], b[k]), k
