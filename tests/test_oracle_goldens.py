"""Pins the CPU oracle against the reference's own RNG-free golden vectors
(/root/reference/test/functions_test.jl, transcribed in tests/golden/reference_goldens.py).
Comparisons use the reference's own operator: ``==`` where its test says ``==``, isapprox /
atol where it says so (a few-ulp allowance where the only difference is summation order)."""
import numpy as np
import pytest

from golden import reference_goldens as G


def toy_world(fsb, orc):
    p = fsb.Params.default(**G.TESTPARAMS)
    w = orc.World(p)
    w.impact[:] = G.IMPACT
    w.horizon[:] = G.HORIZON
    w.threshold[:] = G.THRESHOLD
    w.inv_assets[:, :] = 0
    w.inv_assets[:, :3] = G.INV_ASSETS[:, :3]
    w.stock_value[:, :] = G.PRICES
    return w


def build_holdings(w):
    # fundholdinit!, functions.jl:119-133 with the picks implied by the golden holdings
    for fund, picks in G.PICKS.items():
        for s in picks:
            w.holdings[fund - 1, s - 1] += (w.fund_value[fund - 1, 0] / len(picks)) / w.stock_value[s - 1, 0]


def isapprox(a, b, rtol=np.sqrt(np.finfo(float).eps)):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


@pytest.fixture()
def world(fsb, orc):
    w = toy_world(fsb, orc)
    w.fundcapitalinit()
    w.fundstakeinit()
    build_holdings(w)
    w.fundvalinit(3)
    yield w
    w.close()


def test_fundcapitalinit_and_stakes_exact(world):
    assert np.array_equal(world.fund_value[:, 0], G.FUND_CAPITAL)            # :121-122 (==)
    assert np.array_equal(world.stakes, G.STAKES)                            # :125-128 (==)
    assert np.array_equal(world.stakes.sum(axis=1), np.ones(3))              # :131


def test_fundholdinit_and_fundvalinit(world):
    assert isapprox(world.holdings, G.HOLDINGS_APPROX)                       # :135-139 (isapprox)
    assert isapprox(world.fund_value, G.FUND_VALUE_HISTORY)                  # :145-151 (isapprox)
    # tighter than the reference: <= 2 ulp element-wise
    nz = G.FUND_VALUE_HISTORY != 0
    rel = np.abs(world.fund_value[nz] - G.FUND_VALUE_HISTORY[nz]) / G.FUND_VALUE_HISTORY[nz]
    assert rel.max() < 4.5e-16


def test_perfreview_q1_denominator(world):
    # :214  perfreview(4, [3], ...) == [3 3]; valchange uses column `horizon` as denominator (Q1)
    d = world.perfreview(4, [3 - 1])
    assert np.array_equal(d + 1, G.PERFREVIEW_T4_REVIEWER3)
    # the other reviewers of the toy world at t = 4, re-derived by hand from the pinned state (the reference has no
    # literal for them): investor 1 (fund 1, h = 1): (V[1,4] - V[1,3]) / V[1,1] with V = 3.18 * price[4,:] -> compare with
    # the same formula evaluated here in plain numpy
    for inv in (0, 1, 3):
        f = int(np.argmax(world.inv_assets[inv, :-1] > 0))
        h = int(world.horizon[inv])
        v = world.holdings[f, :] @ world.stock_value
        expect = (v[3] - v[3 - h]) / v[h - 1] < world.threshold[inv]
        got = world.perfreview(4, [inv])
        assert (len(got) == 1 and got[0, 0] == inv and got[0, 1] == f) if expect else len(got) == 0


def test_liquidate(world):
    sale = world.liquidate(3, G.DIVESTMENTS - 1)                             # :217-234
    assert isapprox(sale, G.SALE_VALUES_APPROX)
    assert np.all(np.signbit(sale))                                          # "-0.0" entries
    assert np.array_equal(world.holdings, G.HOLDINGS_AFTER_LIQUIDATION)
    assert np.array_equal(world.stakes, G.STAKES_AFTER_LIQUIDATION)


def test_executeorder_sell_and_disburse(world):
    cash = world.executeorder_sell(3, G.SALE_VALUES_APPROX)                  # :256-271
    assert isapprox(world.stock_value[:, 2], G.PRICE_COL3_AFTER_SELL)
    assert np.array_equal(world.stock_value[:, [0, 1, 3]], G.PRICES[:, [0, 1, 3]])
    # reference compares with ==; its 5-term sum is sequential, ours is the canonical 4-way
    # tree: identical for the 2-term row, within 1 ulp for the 3-term row
    assert cash[0] == G.CASHOUT[0]
    assert abs(cash[1] - G.CASHOUT[1]) <= np.spacing(G.CASHOUT[1])
    world.disburse_cash(G.DIVESTMENTS - 1, G.CASHOUT_HARDCODED)              # :274-280 (==)
    assert np.array_equal(world.inv_assets, G.INV_ASSETS_AFTER_CASH)


def test_marketmake_exact(fsb, orc):
    p = fsb.Params.default(**dict(G.TESTPARAMS, bigm=2, bigt=3, bigk=1, bign=1, perfwindow=(1, 1)))
    for order, result in ((G.MM_BUY_ORDER, G.MM_BUY_RESULT), (G.MM_SELL_ORDER, G.MM_SELL_RESULT)):
        w = orc.World(p)
        w.stock_value[:, :] = G.MM_PRICES
        w.impact[:] = G.MM_IMPACT
        w.marketmake(3, order)                                               # :418-429 (==)
        assert np.array_equal(w.stock_value[:, 2], result)
        w.close()


def test_marketmake_floor_q10(fsb, orc):
    p = fsb.Params.default(**dict(G.TESTPARAMS, bigm=2, bigt=3, bigk=1, bign=1, perfwindow=(1, 1)))
    w = orc.World(p)
    w.stock_value[:, :] = G.MM_PRICES
    w.impact[:] = G.MM_IMPACT
    w.marketmake(3, np.array([[-200.0, 0.0]]))   # 1 + (-200*0.01) = -1 -> negative -> floor
    assert w.stock_value[0, 2] == 0.000000001 and w.stock_value[1, 2] == 98.0
    assert w.floor_hits() == 1
    w.close()


def test_fundrevalue(fsb, orc):
    p = fsb.Params.default(**G.TESTPARAMS)
    w = orc.World(p)
    w.holdings[:, :] = G.FR_HOLDINGS
    w.stock_value[:, :] = G.FR_PRICES
    w.fundrevalue(range(3))                                                  # :452-455 atol=1
    assert np.allclose(w.fund_value, G.FR_RESULT, rtol=0, atol=1)
    w.close()


def test_respawn_buy_disburse_reinvest_chain(world, orc):
    """The reference's chain :283-383: liquidate -> cash -> respawn -> buy -> shares -> reinvest."""
    w = world
    w.liquidate(3, G.DIVESTMENTS - 1)
    w.executeorder_sell(3, G.SALE_VALUES_APPROX)
    w.disburse_cash(G.DIVESTMENTS - 1, G.CASHOUT_HARDCODED)
    # invariant :288-293: dead funds <= cash holders
    dead = int((w.holdings.sum(axis=1) == 0).sum())
    assert dead <= int((w.inv_assets[:, -1] > 0).sum()) and dead == 1
    w.build_log()
    recs = [(orc.TAPE_ANCHOR, 3, 0, 0, G.RESPAWN_ANCHOR - 1), (orc.TAPE_RPSIZE, 3, 0, 0, len(G.RESPAWN_PICKS))]
    recs += [(orc.TAPE_RPPICK, 3, 0, j, s - 1) for j, s in enumerate(G.RESPAWN_PICKS)]
    w.set_event_tape(recs)
    rc, buy, funds = w.respawn(3)
    assert rc == 0
    assert np.allclose(buy, G.RESPAWN_BUY_VALUES, rtol=0, atol=1e-4)         # :298-299
    assert list(funds + 1) == [3]                                            # :302-303
    assert np.allclose(w.inv_assets, G.INV_ASSETS_AFTER_RESPAWN, rtol=0, atol=0.01)   # :306-310
    assert np.array_equal(w.stakes, G.STAKES_AFTER_RESPAWN)                  # :313-316 (==)
    log = w.log()
    assert list(log["liquidated"]) == [0, 0, 1, 0] and list(log["replacedby"]) == [0, 0, 4, 0]
    assert list(log["failtime"]) == [0, 0, 3, 0] and log["nstocks"][3] == 3 and log["ninvestors"][3] == 1
    shares = w.executeorder_buy(3, buy)
    assert np.allclose(w.stock_value[:, 2], G.PRICE_COL3_AFTER_BUY, rtol=0, atol=1e-4)  # :321-328
    assert np.allclose(shares[0], G.SHARES_OUT, rtol=0, atol=1e-6)           # :331-332
    w.disburse_shares(funds, shares)
    assert np.allclose(w.holdings, G.HOLDINGS_AFTER_SHARES, rtol=0, atol=1e-5)          # :337-340
    assert np.allclose(w.fund_value[2, :4], G.FUND_VALUE_ROW3_AFTER_RESPAWN, rtol=0, atol=1e-3)  # :346-349
    # reinvest!(investors, funds, stocks.value, 3, "best") :364-383; the reference evaluates
    # funds.value as left by the test sequence: fund 2 was never revalued ("FIXME" :344-345),
    # so do not revalue it here either
    rc, order, ofunds = w.reinvest(3, "best")
    assert rc == 0
    assert isapprox(w.inv_assets, G.INV_ASSETS_AFTER_REINVEST)               # :365-369
    assert np.allclose(w.stakes, G.STAKES_AFTER_REINVEST, rtol=0, atol=1e-3)  # :374-377
    assert np.array_equal(order, G.REINVEST_ORDER_VALUES)                    # :380 (==)
    assert list(ofunds + 1) == G.REINVEST_ORDER_FUNDS                        # :383
