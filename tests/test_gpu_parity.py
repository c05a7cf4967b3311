"""Parity tests proper: the CUDA engine, called through the C ABI (ctypes), against the CPU
oracle on the same seeded inputs.  Bar (BASELINE.json north_star): discrete outcomes bit-exact;
FP64 NAVs, prices and flows within 1e-12 relative -- the engine and the oracle share one
canonical operation order, so these tests demand bit-equality and fall back to 1e-12 nowhere.
"""
import importlib
import os

import numpy as np
import pytest

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

SEED = 0x5EEDFACE12345


@pytest.fixture(scope="module")
def tiny_engine(fsb):
    eng = fsb.Engine(H.small_params(fsb), 1, keep_history=True)
    yield eng
    eng.close()


def test_library_loaded_is_in_tree(fsb):
    import os
    so = fsb._lib.so_path()
    assert os.path.exists(so) and "fundstababm.jl_b200/csrc" in so
    assert fsb._lib.lib().fsb_version() == 200


def test_device_draws_bit_exact(tiny_engine, orc):
    """Philox4x32-10 + deterministic inverse-normal on the device == oracle, bit for bit."""
    for site, a, b in ((17, 102, 0), (4, 55, 0), (21, 1349, 3), (22, 7, 0)):
        n = 4096
        dn = tiny_engine.debug_draws(SEED, 12345, site, a, b, 0, n, 0)
        du = tiny_engine.debug_draws(SEED, 12345, site, a, b, 0, n, 1)
        dd = tiny_engine.debug_draws(SEED, 12345, site, a, b, 0, n, 2, 250)
        on = np.array([orc.draw_normal(SEED, 12345, site, a, b, i) for i in range(n)])
        ou = np.array([orc.draw_uniform(SEED, 12345, site, a, b, i) for i in range(n)])
        od = np.array([orc.draw_discrete(SEED, 12345, site, a, b, i, 250) for i in range(n)], dtype=np.float64)
        assert H.eq(dn, on) and H.eq(du, ou) and H.eq(dd, od)
    # extreme uniforms exercise all three AS241 branches
    assert abs(dn).max() > 3.0


def test_canonical_dot_bit_exact(tiny_engine, orc):
    rng = np.random.default_rng(3)
    for n in (1, 2, 5, 63, 64, 65, 250, 251, 5000):
        x = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, n)
        y = rng.standard_normal(n)
        assert tiny_engine.debug_dot(x, y) == orc.dotw(x, y)


@pytest.mark.parametrize("shape", [dict(), dict(bigm=10, bigk=5, bign=33), dict(bigm=3, bigk=2, bign=4, portfsizerange=(1, 3)),
                                   dict(bigm=70, bigk=9, bign=64, portfsizerange=(1, 70))])
def test_init_device_matches_oracle(fsb, orc, shape):
    p = H.small_params(fsb, **shape)
    with fsb.Engine(p, 3, rep_offset=5, keep_history=True) as eng:
        eng.init_device(SEED)
        for r in range(3):
            w = orc.World(p, seed=SEED, rep=5 + r)
            w.initialise()
            st = eng.download_state(r)
            H.assert_state_equal(st, w, fund_value=False)
            # funds.value right after initialise: column 1 is the capital, burn-in columns are
            # sum(holdings .* price) in the reference and a fused dot here (<= 1e-12 relative)
            assert H.eq(st["fund_value"][:, 0], w.fund_value[:, 0])
            np.testing.assert_allclose(st["fund_value"], w.fund_value, rtol=1e-12, atol=0)
            w.build_log()
            H.assert_log_equal(eng.download_log(r), w.log())
            w.close()


@pytest.mark.parametrize("selector", ["probabilistic", "best", "goodenough", "random"])
@pytest.mark.parametrize("shape", [dict(), dict(bigm=11, bigk=5, bign=50, reviewprobability=0.3),
                                   dict(bigm=64, bigk=3, bign=12, portfsizerange=(1, 2), reviewprobability=0.5)])
def test_trajectory_small_worlds(fsb, orc, selector, shape):
    """Whole runs of small worlds (many respawns, shared funds, odd M): every trace point equal."""
    p = H.small_params(fsb, **shape)
    R = 4
    t0, t1 = p.perfwindow[-1] + 2, p.bigt
    with fsb.Engine(p, R, rep_offset=100, keep_history=True) as eng:
        eng.init_device(SEED)
        for r in range(R):
            eng.set_trace(r)
        eng.run(selector=selector)
        for r in range(R):
            w = H.oracle_world(orc, p, SEED, 100 + r)
            assert w.modelrun(selector=selector) == 0
            w.finalise_fund_value(t1)
            H.assert_state_equal(eng.download_state(r), w)
            H.assert_log_equal(eng.download_log(r), w.log())
            H.assert_trace_equal(eng.download_trace(r), w, (t0, t1))
            c = eng.download_compact(r)
            assert H.eq(c["nav_close"], w.nav_close)
            w.close()


def test_upload_path_equals_device_init_path(fsb, orc):
    """Drop-in mode: upload the reference-layout structs of an initialised world, run, download."""
    p = H.small_params(fsb)
    w = H.oracle_world(orc, p, SEED, 7)
    market, stocks, investors, funds = H.structs_from_world(fsb, w)
    log = fsb.Func.modelrun(market, stocks, investors, funds, p, "probabilistic", seed=SEED, rep=7, as_dataframe=False)
    assert w.modelrun(selector="probabilistic") == 0
    w.finalise_fund_value(p.bigt)
    st = dict(market=market.value, stock_value=stocks.value, beta=stocks.beta, vol=stocks.vol, impact=stocks.impact,
              volume=stocks.volume, inv_assets=investors.assets, horizon=investors.horizon,
              threshold=investors.threshold, holdings=funds.holdings, stakes=funds.stakes, fund_value=funds.value)
    H.assert_state_equal(st, w)
    H.assert_log_equal(log, w.log())
    # the runtime invariant check of src/functions.jl:432-444 gives the same verdict on both
    ost = H.structs_from_world(fsb, w)
    assert fsb.Func.boundstest(market, stocks, investors, funds) == fsb.Func.boundstest(*ost)
    w.close()


def test_stop_and_resume_equals_single_run(fsb, orc):
    p = H.small_params(fsb)
    t0 = p.perfwindow[-1] + 2
    with fsb.Engine(p, 2, keep_history=True) as a, fsb.Engine(p, 2, keep_history=True) as b:
        a.init_device(SEED)
        b.init_device(SEED)
        a.run()
        b.run(t0, 30)
        b.run(31, 31)
        b.run(32, p.bigt)
        for r in range(2):
            sa, sb = a.download_state(r), b.download_state(r)
            for k in sa:
                assert H.eq(sa[k], sb[k]), k


def test_ring_history_equals_full_history(fsb):
    """keep_history=0 (burn-in + ring of W+1 columns) gives the same trajectory as the full history."""
    p = H.small_params(fsb)
    with fsb.Engine(p, 3, keep_history=True) as a, fsb.Engine(p, 3, keep_history=False) as b:
        a.init_device(SEED)
        b.init_device(SEED)
        a.run()
        b.run()
        for r in range(3):
            ca, cb = a.download_compact(r), b.download_compact(r)
            for k in ca:
                assert H.eq(ca[k], cb[k]), k
            la, lb = a.download_log(r), b.download_log(r)
            for k in la:
                assert H.eq(la[k], lb[k]), k


def test_sharding_invariance(fsb):
    """Philox counters use GLOBAL replication ids: any sharding gives identical replications and
    the integer ensemble statistics add up exactly (the N>1 contract of SURVEY.md 8e)."""
    p = H.small_params(fsb)
    with fsb.Engine(p, 8) as full, fsb.Engine(p, 3, rep_offset=0) as s0, fsb.Engine(p, 5, rep_offset=3) as s1:
        for e in (full, s0, s1):
            e.init_device(SEED)
            e.run()
        for r in range(8):
            cf = full.download_compact(r)
            cs = s0.download_compact(r) if r < 3 else s1.download_compact(r - 3)
            for k in cf:
                assert H.eq(cf[k], cs[k]), k
        assert H.eq(full.stats_raw(), s0.stats_raw() + s1.stats_raw())


def test_dense_tape_equivalence_mode(fsb, orc):
    """Equivalence mode: the exogenous shocks come from a tape (here produced by numpy, standing in
    for a Julia recorder) and are fed to both implementations."""
    p = H.small_params(fsb)
    rng = np.random.default_rng(11)
    z_mkt = rng.standard_normal(p.bigt)
    z_stock = rng.standard_normal((p.bigm, p.bigt))
    u_review = rng.random((p.bign, p.bigt))
    w = H.oracle_world(orc, p, SEED, 0)
    market, stocks, investors, funds = H.structs_from_world(fsb, w)
    w.set_dense_tape(z_mkt, z_stock, u_review)
    assert w.modelrun(selector="probabilistic") == 0
    w.finalise_fund_value(p.bigt)
    fsb.Func.modelrun(market, stocks, investors, funds, p, "probabilistic", seed=SEED, rep=0,
                      dense_tape=dict(z_mkt=z_mkt, z_stock=z_stock, u_review=u_review))
    assert H.eq(stocks.value, w.stock_value) and H.eq(funds.holdings, w.holdings)
    assert H.eq(funds.value, w.fund_value) and H.eq(investors.assets, w.inv_assets)
    w.close()


def test_event_tape_forces_choices_and_anchors(fsb, orc):
    """Event tape: replay the oracle's own anchors / portfolio picks / fund choices as a tape."""
    p = H.small_params(fsb)
    w = H.oracle_world(orc, p, SEED, 3)
    market, stocks, investors, funds = H.structs_from_world(fsb, w)
    assert w.modelrun(selector="probabilistic") == 0
    w.finalise_fund_value(p.bigt)
    recs, ordinal, pick_j = [], {}, {}
    for e in w.trace():
        t = int(e["t"])
        if e["kind"] == orc.EV_SPAWN:
            i = ordinal.get(t, 0)
            ordinal[t] = i + 1
            recs.append((fsb._lib.TAPE_ANCHOR, t, i, 0, float(e["b"])))
            pick_j[t] = (i, 0)
        elif e["kind"] == orc.EV_SPAWN_N:
            recs.append((fsb._lib.TAPE_RPSIZE, t, pick_j[t][0], 0, float(e["b"])))
        elif e["kind"] == orc.EV_PICK:
            i, j = pick_j[t]
            recs.append((fsb._lib.TAPE_RPPICK, t, i, j, float(e["b"])))
            pick_j[t] = (i, j + 1)
        elif e["kind"] == orc.EV_CHOICE:
            recs.append((fsb._lib.TAPE_CHOICE_IDX, t, int(e["a"]), 0, float(e["b"])))
    assert len(recs) > 10
    # a different seed: only the shocks/reviews come from Philox, so feed those as a dense tape too
    T, M, N = p.bigt, p.bigm, p.bign
    z_mkt = np.array([0.0] + [orc.draw_normal(SEED, 3, 16, t, 0, 0) for t in range(1, T + 1)])[1:]
    z_stock = np.array([[orc.draw_normal(SEED, 3, 17, t, 0, m) for t in range(1, T + 1)] for m in range(M)])
    u_review = np.array([[orc.draw_uniform(SEED, 3, 18, t, 0, n) for t in range(1, T + 1)] for n in range(N)])
    fsb.Func.modelrun(market, stocks, investors, funds, p, "random", seed=999, rep=77,
                      dense_tape=dict(z_mkt=z_mkt, z_stock=z_stock, u_review=u_review), event_tape=recs)
    assert H.eq(stocks.value, w.stock_value) and H.eq(funds.holdings, w.holdings)
    assert H.eq(investors.assets, w.inv_assets) and H.eq(funds.stakes, w.stakes)
    w.close()


def test_default_shape_trajectories(fsb, orc):
    """BASELINE configs[0]/[1] shape: K=M=250, N=1000, T=1350, full 1249-step horizon."""
    p = fsb.Params.default()
    R = 3
    with fsb.Engine(p, R, rep_offset=4000, keep_history=True) as eng:
        eng.init_device(SEED)
        eng.set_trace(1)
        eng.run()
        for r in range(R):
            w = H.oracle_world(orc, p, SEED, 4000 + r, trace=(r == 1))
            assert w.modelrun() == 0
            w.finalise_fund_value(p.bigt)
            H.assert_state_equal(eng.download_state(r), w)
            H.assert_log_equal(eng.download_log(r), w.log())
            if r == 1:
                H.assert_trace_equal(eng.download_trace(r), w, (p.perfwindow[-1] + 2, p.bigt))
            w.close()


def test_batch_properties_and_sampled_oracle(fsb, orc):
    """A larger batch (ring history): size-independent invariants for every replication plus
    bit-exact comparison of a sample against the oracle."""
    p = fsb.Params.default()
    R, t_last = 512, 160
    with fsb.Engine(p, R, rep_offset=10_000) as eng:
        eng.init_device(SEED)
        eng.run(t_last=t_last)
        assert not eng.rep_status().any()
        st = eng.stats()[0]
        assert st["replications"] == R and st["error_replications"] == 0
        assert st["replication_steps"] == R * (t_last - 101)
        assert st["reinvestments"] <= st["divestments"] <= st["reviews"]
        assert st["funds_ever"] == R * p.bigk + st["liquidations"]
        assert st["flow_hist"].sum() == st["replication_steps"]
        for r in list(range(0, R, 37)) + [R - 1]:
            c = eng.download_compact(r)
            assert ((c["inv_fund"] >= 0) & (c["inv_fund"] <= p.bigk)).all()
            assert not ((c["inv_fund"] == p.bigk) & (c["inv_amount"] > 0)).any()   # no cash left at step end
            sums = np.bincount(c["inv_fund"], weights=c["inv_stake"], minlength=p.bigk + 1)[: p.bigk]
            if not np.isnan(c["price_now"]).any():
                np.testing.assert_allclose(sums[sums > 0], 1.0, rtol=0, atol=1e-12)   # stakes sum to one per fund
            assert not (c["holdings"] < 0).any() and not (c["price_now"] <= 0).any()
        for r in (0, 255, 511):
            w = H.oracle_world(orc, p, SEED, 10_000 + r, trace=False)
            assert w.modelrun(t_last=t_last) == 0
            c = eng.download_compact(r)
            assert H.eq(c["holdings"], w.holdings)
            assert H.eq(c["price_now"], w.stock_value[:, t_last - 1])
            assert H.eq(c["nav_close"], w.nav_close)
            H.assert_log_equal(eng.download_log(r), w.log())
            w.close()


def test_statistics_match_oracle_logs(fsb, orc):
    p = H.small_params(fsb)
    R = 6
    with fsb.Engine(p, R) as eng:
        eng.init_device(SEED)
        eng.run()
        st = eng.stats()[0]
    liq, rows, floors, divs, revs, reinv = 0, 0, 0, 0, 0, 0
    fail_hist = np.zeros(p.bigt + 1, dtype=np.int64)
    for r in range(R):
        w = H.oracle_world(orc, p, SEED, r)
        assert w.modelrun() == 0
        log, tr = w.log(), w.trace()
        liq += int(log["liquidated"].sum())
        rows += len(log["liquidated"])
        floors += w.floor_hits()
        revs += int((tr["kind"] == orc.EV_REVIEWER).sum())
        divs += int((tr["kind"] == orc.EV_DIVEST).sum())
        reinv += int((tr["kind"] == orc.EV_CHOICE).sum())
        np.add.at(fail_hist, log["failtime"][log["liquidated"] == 1], 1)
        w.close()
    assert (st["liquidations"], st["funds_ever"], st["floor_hits"]) == (liq, rows, floors)
    assert (st["reviews"], st["divestments"], st["reinvestments"]) == (revs, divs, reinv)
    assert H.eq(st["failtime_hist"], fail_hist)
    assert st["liquidations_per_replication_hist"].sum() == R and st["market_drawdown_hist"].sum() == R


def test_calibrate_sweep_points_are_independent(fsb):
    """configs[4] in miniature: a sweep engine gives, per point, the statistics a single-point
    engine gives for the same global replication ids."""
    base = H.small_params(fsb)
    pts = fsb.grid(base, thresholdstd=(0.02, 0.1), impactscale=(1.0, 4.0))
    reps = 5
    stats = fsb.calibrate(pts, reps, seed=SEED)
    assert len(stats) == 4
    for i, pt in enumerate(pts):
        with fsb.Engine([base] * i + [pt], reps, rep_offset=i * reps, reps_per_point=reps) as eng:
            eng.init_device(SEED)
            eng.run()
            single = eng.stats()[i]
        for k, v in stats[i].items():
            assert H.eq(v, single[k]), (i, k)
    assert stats[0]["divestments"] != stats[1]["divestments"] or stats[0]["liquidations"] != stats[2]["liquidations"]


def test_error_behaviour(fsb):
    p = H.small_params(fsb)
    L = fsb._lib
    with fsb.Engine(p, 2, keep_history=True) as eng:
        with pytest.raises(L.FsbError) as ei:      # no state yet
            eng.run()
        assert ei.value.code == L.ERR_BAD_STATE
        eng.init_device(SEED)
        assert L.lib().fsb_run(eng.handle, 9, 20, 7, 0) == L.ERR_BAD_SELECTOR      # Q9
        assert L.lib().fsb_run(eng.handle, 3, 20, 2, 0) == L.ERR_RANGE
        # days are simulated once and in order: skipping or re-running a day is refused
        assert eng.next_day() == p.perfwindow[-1] + 2
        assert L.lib().fsb_run(eng.handle, 12, 20, 2, 0) == L.ERR_RANGE
        eng.run(t_last=20)
        assert eng.next_day() == 21
        assert L.lib().fsb_run(eng.handle, 20, 25, 2, 0) == L.ERR_RANGE
        assert L.lib().fsb_run_async(eng.handle, 23, 25, 2, 0) == L.ERR_RANGE
        eng.run(t_last=25)                                                          # (continues with day 21)
        with pytest.raises(ValueError):
            fsb.Func.modelrun(*fsb.Types.allocate(p), p)                            # default "bestperformer" (Q9)
        with pytest.raises(L.FsbError) as ei:
            eng.download_state(5)
        assert ei.value.code == L.ERR_RANGE
    with fsb.Engine(p, 1) as eng:
        eng.init_device(SEED)
        with pytest.raises(L.FsbError) as ei:
            eng.download_state(0)
        assert ei.value.code == L.ERR_NO_HISTORY
    # capacity overflow is an error, never a silent drop
    with fsb.Engine(H.small_params(fsb, reviewprobability=0.9), 2, event_capacity=8) as eng:
        eng.init_device(SEED)
        with pytest.raises(L.FsbError) as ei:
            eng.run()
        assert ei.value.code == L.ERR_REPLICATION
        assert (eng.rep_status() & L.REP_EVENT_OVERFLOW).all()
    # an uploaded state with two positions for one investor is rejected
    market, stocks, investors, funds = fsb.Types.allocate(p)
    investors.assets[0, 0] = 1.0
    investors.assets[0, 1] = 1.0
    investors.horizon[:] = 1
    with fsb.Engine(p, 1, keep_history=True) as eng:
        with pytest.raises(L.FsbError) as ei:
            eng.upload_state(0, market, stocks, investors, funds)
        assert ei.value.code == L.ERR_BAD_STATE


def test_single_process_multi_gpu_statistics(fsb):
    """fsb_comm_init_all + fsb_stats_all: ONE process, one engine per device (the path a Julia caller uses), against the
    per-engine statistics and against a single engine holding all replications."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    cal = importlib.import_module(fsb.__name__ + ".calibrate")
    eng_mod = importlib.import_module(fsb.__name__ + ".engine")
    p = fsb.Params.default(bigt=130)
    pts = cal.grid(p, thresholdstd=(0.02, 0.08))
    reps, devices = 24, [0, 1]
    got = cal.calibrate_single_process(pts, reps, devices, seed=SEED)
    with fsb.Engine(pts, len(pts) * reps, reps_per_point=reps) as one:
        one.init_device(SEED)
        try:
            one.run()
        except fsb._lib.FsbError as ex:
            assert ex.code == fsb._lib.ERR_REPLICATION
        want = one.stats()
    for a, b in zip(got, want):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    # and the collective against the sum of the engines' local statistics
    total = len(pts) * reps
    engines = [fsb.Engine(pts, total // 2, rep_offset=g * (total // 2), reps_per_point=reps, device=d) for g, d in enumerate(devices)]
    try:
        eng_mod.comm_init_all(engines)
        for e in engines:
            e.init_device(SEED)
            e.run_async(t_last=120)
        for e in engines:
            e.sync()
        local = sum(e.stats_raw(local=True).astype(np.int64) for e in engines)
        assert np.array_equal(eng_mod.stats_all(engines).astype(np.int64), local)
    finally:
        for e in engines:
            e.close()


@pytest.mark.parametrize("env", [{}, {"FSB_PASS_PACKED": "1"}, {"FSB_FORCE_BIG": "1"}, {"FSB_BATCHES": "3"}, {"FSB_NO_TMA": "1"}],
                         ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()) or "default")
def test_no_kernel_writes_outside_its_buffers(fsb, env, monkeypatch):
    """compute-sanitizer is closed on this GPU pool (profiles/r2_sanitizer.md), so out-of-bounds WRITES are hunted with the
    library's own guard bands: under FSB_REDZONE=1 every device buffer sits between two 256-byte bands of a fill pattern;
    after init, a run with traces, the host-tape step, statistics, stylised facts and every download, no band may have changed."""
    monkeypatch.setenv("FSB_REDZONE", "1")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for p, R, days in ((fsb.Params.default(), 24, 70), (H.small_params(fsb), 5, 50),
                       (H.small_params(fsb, bigk=11, bign=90, bigm=17, bigt=140, perfwindow=(1, 12), portfsizerange=(1, 9)), 7, 120)):
        with fsb.Engine(p, R, keep_history=True, rep_offset=3) as eng:
            eng.init_device(SEED)
            eng.set_trace(1)
            t0 = p.perfwindow[-1] + 2
            eng.run(t_last=t0 + days - 2)
            rng = np.random.default_rng(5)
            nav = np.zeros((R, p.bigk))
            eng.run_step_hosttape(t0 + days - 1, rng.standard_normal(R), rng.standard_normal((R, p.bigm)), rng.random((R, p.bign)), nav)
            eng.stats()
            eng.stylised_facts(maxlag=3)
            for r in range(R):
                eng.download_state(r)
                eng.download_log(r)
            eng.download_trace(1)
            n, bad = eng.redzones()
            assert n > 30 and bad == 0, (n, bad)
    monkeypatch.delenv("FSB_REDZONE")
    with fsb.Engine(H.small_params(fsb), 2) as eng:   # (without the guard bands the call says so)
        assert eng.redzones()[0] == -1


def test_repeated_runs_are_bit_identical(fsb):
    """A data race (the pass kernel reuses its TMA ring slots, kernels hand lists over through global memory, batches run on
    separate streams) would show up as run-to-run differences: the same 1536 default-shape replications three times, with
    one and with four replication batches, every replication's holdings / NAVs / prices compared through checksums."""
    import hashlib
    import os
    p = fsb.Params.default()
    R, t_last = 1536, 101 + 48

    def run():
        with fsb.Engine(p, R, rep_offset=40_000) as eng:
            eng.init_device(SEED)
            eng.run(t_last=t_last)
            h = hashlib.blake2b(digest_size=16)
            for r in range(0, R, 3):
                c = eng.download_compact(r)
                for k in ("holdings", "nav_close", "nav_open", "price_now", "inv_stake", "inv_amount"):
                    h.update(np.ascontiguousarray(c[k]).tobytes())
            return h.hexdigest(), eng.stats_raw().tolist()

    first = run()
    assert run() == first
    os.environ["FSB_BATCHES"] = "4"
    try:
        assert run() == first
    finally:
        del os.environ["FSB_BATCHES"]


def test_stakeless_fund_with_holdings_keeps_the_volume_quirk(fsb, orc):
    """functions.jl:471-476 is entered whenever a fund's stakes sum to zero, also when respawn! then finds no all-zero
    holdings row (Q5): trackvolume! still subtracts the sell orders once more (:474, Q3).  Reached through an uploaded
    state whose stakes are not normalised: a fund's only staked member, with stake 1/2, divests; the fund keeps half of its holdings and
    stays stakeless for good, so every later day with divestments takes that branch."""
    p = H.small_params(fsb, bigt=45)
    w = H.oracle_world(orc, p, SEED, 5, trace=False)
    for k in range(p.bigk):  # one staked member per fund, holding half of it
        members = np.flatnonzero(w.stakes[k] > 0)
        w.stakes[k, members[1:]] = 0.0
        w.stakes[k, members[0]] = 0.5
    market, stocks, investors, funds = H.structs_from_world(fsb, w)
    with fsb.Engine(p, 1, keep_history=True, rep_offset=5) as eng:
        eng.upload_state(0, market, stocks, investors, funds)
        eng.set_seed(SEED)
        eng.run()
        assert w.modelrun() == 0
        w.finalise_fund_value(p.bigt)
        st = eng.download_state(0)
        H.assert_state_equal(st, w)
        # the branch was really taken: some fund ends stakeless with holdings left
        assert any(w.stakes[k].sum() == 0 and w.holdings[k].sum() != 0 for k in range(p.bigk))
    w.close()


@pytest.mark.parametrize("env", [
    {"FSB_CMAX": "6"},                          # small column chunks: several holdings passes per day
    {"FSB_NO_FIXED": "1"},                      # the generic build of the event kernels instead of the fixed-shape one
    {"FSB_SCAN_CHOICE": "1"},                   # the rank rule sampled by the floating-point scan alone (no integer shortcut)
    {"FSB_PASS_PACKED": "1"},                   # the pass over the packed mirror of the holdings (fsb_passx.cuh, opt-in)
    {"FSB_PASS_PACKED": "1", "FSB_XCOLS": "4"},  # ... with small column chunks: several passes over the packed rows per day
    {"FSB_PASS_AHEAD": "0"},                    # no look-ahead for the pass CTA's next replication (header + columns staged on demand)
    {"FSB_PASS16": "0"},                        # chunks of <= 8 columns in the 2 x 2 lane arrangement too (8-row trips only)
    {"FSB_PASS16": "0", "FSB_CMAX": "6"},
    {"FSB_NO_TMA": "1"},                        # register-ring pass (the path of non-default shapes) on the default shape
    {"FSB_NO_TMA": "1", "FSB_CMAX": "6"},
    {"FSB_BATCHES": "3"},                       # replication batches on separate streams
    {"FSB_EVENT_ROWS": "1"},                    # order rows spill to the global scratch
    {"FSB_FORCE_BIG": "1"},                     # the big-shape build of the step kernels (per-asset arrays in global memory)
], ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_schedule_variants_are_invisible(fsb, orc, env, monkeypatch):
    """Launch-schedule knobs (column chunking, TMA-fed or register-fed pass, replication batches, order-row
    spilling) must not change a single bit: default shape, a few replications against the default
    schedule and against the oracle."""
    p = fsb.Params.default()
    R, t_last = 12, 180
    with fsb.Engine(p, R, rep_offset=77) as ref:
        ref.init_device(SEED)
        ref.run(t_last=t_last)
        want = [ref.download_compact(r) for r in range(R)]
        want_stats = ref.stats_raw().tolist()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    with fsb.Engine(p, R, rep_offset=77) as eng:
        eng.init_device(SEED)
        eng.run(t_last=t_last)
        assert not eng.rep_status().any()
        for r in range(R):
            got = eng.download_compact(r)
            for k in got:
                assert H.eq(got[k], want[r][k]), (k, r)
        assert eng.stats_raw().tolist() == want_stats
    w = H.oracle_world(orc, p, SEED, 77, trace=False)
    assert w.modelrun(t_last=t_last) == 0
    assert H.eq(want[0]["holdings"], w.holdings) and H.eq(want[0]["nav_close"], w.nav_close)
    w.close()


@pytest.mark.parametrize("psize", [(1, 250), (120, 250), (1, 3)], ids=["mixed", "dense", "tiny"])
def test_packed_mirror_dense_rows_and_repacking(fsb, orc, psize, monkeypatch):
    """The opt-in pass over the packed mirror of the holdings (fsb_passx.cuh) at portfolio sizes the default Params never
    produce: rows whose longest chain exceeds the packed capacity are read densely, respawned and reinvested rows are
    re-packed.  Default K / M / N (the shape the packed pass supports), every array against the oracle."""
    monkeypatch.setenv("FSB_PASS_PACKED", "1")
    p = fsb.Params.default(bigt=150, portfsizerange=psize)
    R = 3
    with fsb.Engine(p, R, rep_offset=11, keep_history=True) as eng:
        eng.init_device(SEED)
        eng.run()
        for r in range(R):
            w = H.oracle_world(orc, p, SEED, 11 + r, trace=False)
            assert w.modelrun() == 0
            w.finalise_fund_value(p.bigt)
            H.assert_state_equal(eng.download_state(r), w)
            H.assert_log_equal(eng.download_log(r), w.log())
            w.close()


@pytest.mark.parametrize("shape", [
    dict(bigk=300, bign=700, bigm=40, bigt=70, perfwindow=(1, 9), portfsizerange=(3, 20)),    # K > 256: rank by counting
    dict(bigk=20, bign=90, bigm=300, bigt=60, perfwindow=(1, 8), portfsizerange=(5, 120)),    # M > 256: column pitch 384
    dict(bigk=260, bign=600, bigm=270, bigt=55, perfwindow=(1, 6), portfsizerange=(10, 200)),  # both (config-4 like, scaled down)
], ids=["K300", "M300", "K260_M270"])
def test_large_shapes_general_path(fsb, orc, shape):
    """Shapes beyond the default one (BASELINE.json configs[3] scaled down): the general, untuned code paths
    (register-fed pass, rank by counting, wider columns) against the oracle, all selectors' hot one included."""
    p = fsb.Params.default(reviewprobability=0.12, invcaprange=(10, 1000), **shape)
    R = 3
    with fsb.Engine(p, R, keep_history=True) as eng:
        eng.init_device(SEED)
        eng.run(selector="probabilistic")
        assert not eng.rep_status().any()
        for r in range(R):
            w = H.oracle_world(orc, p, SEED, r, trace=False)
            assert w.modelrun(selector="probabilistic") == 0
            w.finalise_fund_value(p.bigt)
            H.assert_state_equal(eng.download_state(r), w)
            H.assert_log_equal(eng.download_log(r), w.log())
            w.close()


@pytest.mark.parametrize("shape,env", [
    (dict(bigk=600, bign=2000, bigm=40, bigt=50), {}),                          # generic build: 1024 = 8 x 128 threads, register network
    (dict(bigk=600, bign=2000, bigm=40, bigt=50), {"FSB_SMEM_SORT": "1"}),      # ... the shared-memory network
    (dict(bigk=600, bign=2000, bigm=40, bigt=50), {"FSB_SCAN_CHOICE": "1"}),    # ... sampled by the reference's scan only
    (dict(bigk=1100, bign=3300, bigm=48, bigt=28), {"FSB_FORCE_BIG": "1"}),     # big build: 2048 = 8 x 256 threads, register network
    (dict(bigk=1100, bign=3300, bigm=48, bigt=28), {"FSB_FORCE_BIG": "1", "FSB_SMEM_SORT": "1", "FSB_SCAN_CHOICE": "1"}),
], ids=lambda v: ",".join(f"{k}={x}" for k, x in v.items()) or "default")
def test_cta_ranked_columns(fsb, orc, shape, env, monkeypatch):
    """256 < K <= 2048: the CTA ranks a horizon column with a bitonic network (pairs in registers when the padded length is
    8 per thread, else in shared memory) and samples the rank rule through exact integer prefix sums; both networks and the
    reference's floating-point scan must give the oracle's choices bit for bit."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    p = fsb.Params.default(reviewprobability=0.12, invcaprange=(10, 1000), perfwindow=(1, 8), portfsizerange=(3, 20), **shape)
    R = 2 if p.bigk < 1000 else 1
    with fsb.Engine(p, R, keep_history=True) as eng:
        eng.init_device(SEED)
        eng.run(selector="probabilistic")
        assert not eng.rep_status().any()
        for r in range(R):
            w = H.oracle_world(orc, p, SEED, r, trace=False)
            assert w.modelrun(selector="probabilistic") == 0
            w.finalise_fund_value(p.bigt)
            H.assert_state_equal(eng.download_state(r), w)
            H.assert_log_equal(eng.download_log(r), w.log())
            w.close()


@pytest.mark.parametrize("shape,R", [
    (dict(bigk=64, bign=256, bigm=5000, bigt=125, portfsizerange=(2500, 5000), reviewprobability=0.1), 3),
    (dict(bigk=2000, bign=8000, bigm=5000, bigt=110, portfsizerange=(2500, 5000)), 1),
], ids=["K64_M5000", "config4_K2000_M5000_N8000"])
def test_big_shapes_per_asset_arrays_in_global_memory(fsb, orc, shape, R):
    """BASELINE.json configs[3] (2000 funds x 5000 assets, dense overlapping portfolios, N = 4K investors, W = 100)
    for its first simulated days, and a 5000-asset universe with fewer funds for longer: a price column (40 KB)
    no longer fits in shared memory next to the event lists, so the engine launches the big-shape build of the
    three step kernels (fsb_step_big.cu).  Every array against the oracle, bit for bit."""
    p = fsb.Params.default(**shape)
    with fsb.Engine(p, R, keep_history=True) as eng:
        eng.init_device(SEED)
        eng.run(selector="probabilistic")
        assert not eng.rep_status().any()
        for r in range(R):
            w = H.oracle_world(orc, p, SEED, r, trace=False)
            assert w.modelrun(selector="probabilistic") == 0
            w.finalise_fund_value(p.bigt)
            H.assert_state_equal(eng.download_state(r), w)
            H.assert_log_equal(eng.download_log(r), w.log())
            w.close()


def test_hosttape_steps_match_the_oracle_on_the_same_draws(fsb, orc):
    """fsb_run_step_hosttape (the e2e path of bench.py): every step's shocks come from host buffers, replication
    batches run on separate streams with their transfers; the oracle gets the same draws as dense tapes."""
    p = H.small_params(fsb, bigk=7, bign=60, bigm=12)
    R, T0, NS = 9, p.perfwindow[-1] + 2, 25
    rng = np.random.default_rng(99)
    zm = rng.standard_normal((NS, R))
    zs = rng.standard_normal((NS, R, p.bigm))
    ur = rng.random((NS, R, p.bign))
    with fsb.Engine(p, R) as eng:
        eng.init_device(SEED)
        nav = np.zeros((R, p.bigk))
        navs = []
        for i in range(NS):
            eng.run_step_hosttape(T0 + i, np.ascontiguousarray(zm[i]), np.ascontiguousarray(zs[i]), np.ascontiguousarray(ur[i]), nav)
            navs.append(nav.copy())
        assert not eng.rep_status().any()
        got = [eng.download_compact(r) for r in range(R)]
    for r in (0, 4, 8):
        w = H.oracle_world(orc, p, SEED, r, trace=True)
        z_mkt = np.zeros(p.bigt); z_stock = np.zeros((p.bigm, p.bigt), order="F"); u_rev = np.zeros((p.bign, p.bigt), order="F")
        for i in range(NS):
            z_mkt[T0 + i - 1] = zm[i, r]; z_stock[:, T0 + i - 1] = zs[i, r]; u_rev[:, T0 + i - 1] = ur[i, r]
        w.set_dense_tape(z_mkt, z_stock, u_rev)
        assert w.modelrun(T0, T0 + NS - 1, "probabilistic", 0) == 0
        assert H.eq(got[r]["holdings"], w.holdings) and H.eq(got[r]["nav_close"], w.nav_close)
        assert H.eq(got[r]["price_now"], w.stock_value[:, T0 + NS - 2])
        for i in range(NS):
            assert H.eq(navs[i][r], w.trace_nav_open[:, T0 + i - 1]), i
        w.close()


def test_replay_of_a_recording_file(fsb, orc, tmp_path):
    """scripts/replay_reference_tapes.py on a recording in the format julia/record_tapes.jl writes (the oracle
    stands in for the reference, which cannot run in this image): upload the recorded state, feed the recorded
    draws, compare with the recorded outcome.  Against the oracle the match is exact."""
    import importlib.util
    tf = importlib.import_module(fsb.__name__ + ".tapefile")
    spec = importlib.util.spec_from_file_location("replay_reference_tapes", os.path.join(ROOT, "scripts", "replay_reference_tapes.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    p = H.small_params(fsb, bigk=9, bigm=9, bign=70, bigt=80)
    rec, w = H.oracle_recording(fsb, orc, p, seed=33, rep=1, selector="probabilistic")
    w.close()
    assert len(rec["events"]) > 20
    path = tmp_path / "rec.fsbtape"
    tf.write_tape_file(path, **rec)
    worst = mod.replay(fsb, tf.read_tape_file(path))
    assert max(worst.values()) == 0.0, worst
