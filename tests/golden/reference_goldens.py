"""RNG-free golden vectors transcribed from the reference's own unit tests,
/root/reference/test/functions_test.jl (line numbers cited per item).

Only literals that do not depend on Julia's MersenneTwister streams are used as truth; seeded
draws (betas, vols, prices, ...) appear here as fixed INPUT states exactly as the test prints them.
Indices are 1-based as in the Julia source; the tests convert.
"""
import numpy as np

# testparams, test/functions_test.jl:7-25
TESTPARAMS = dict(
    bigk=3, bign=4, bigm=5, bigt=6, marketstartval=100, drift=0.05, marketvol=0.1,
    perfwindow=(1, 3), betamean=1, betastd=0.3, stockstartval=100,
    stockvolrange=(0.01, 0.01, 0.1), invcaprange=(10, 1000), thresholdmean=0, thresholdstd=0.05,
    portfsizerange=(1, 5), impactrange=(0.00001, 0.00001, 0.0001), reviewprobability=1 / 63)

IMPACT = np.array([0.00008, 0.00004, 0.00007, 0.00001, 0.00004])            # :69-70
HORIZON = np.array([1, 3, 1, 1])                                             # :82-83
THRESHOLD = np.array([0.022728468593375205, -0.03299475538810838,
                      0.12695393970324698, -0.02006345690377245])            # :87-91
# investor assets: diag [318,111,692] (:95-96), total 2013 (:102) => investor 4 holds 892;
# fund capital [318,1003,692] (:121-122) puts it in fund 2
INV_ASSETS = np.array([[318.0, 0, 0, 0], [0, 111.0, 0, 0], [0, 0, 692.0, 0], [0, 892.0, 0, 0]])
FUND_CAPITAL = np.array([318.0, 1003.0, 692.0])                              # :121-122
STAKES = np.array([[1.0, 0.0, 0.0, 0.0],
                   [0.0, 0.1106679960119641, 0.0, 0.8893320039880359],
                   [0.0, 0.0, 1.0, 0.0]])                                     # :125-128
# holdings (:136-139): portfolio picks that generate them (fund: list of 1-based stocks)
PICKS = {1: [4], 2: [1, 3, 3, 3, 4], 3: [1, 4]}
HOLDINGS_APPROX = np.array([[0, 0, 0, 318 / 100, 0],
                            [1003 / 500, 0, 3009 / 500, 1003 / 500, 0],
                            [692 / 200, 0, 0, 692 / 200, 0]])
# prices, columns 1-4 (:248-254)
PRICES = np.array([
    [100.0, 107.903010980603, 110.54316357113026, 119.08511622272532, 0, 0],
    [100.0, 80.4519644607866, 85.81763860422608, 86.49247760086409, 0, 0],
    [100.0, 118.68295171596901, 123.6135620896176, 126.53705519388893, 0, 0],
    [100.0, 103.10552983223091, 110.40592725161265, 116.03243809686083, 0, 0],
    [100.0, 103.28424280970195, 111.56926365234682, 112.85619511929654, 0, 0]])
# fundvalinit! with perfwindow=3:3 (:145-151), compared with isapprox in the reference
FUND_VALUE_HISTORY = np.array([
    [318, 327.8755848664943, 351.09084866012824, 368.98315314801744, 0, 0],
    [1003, 1137.5171362972462, 1187.1302928457408, 1233.1458121219132, 0, 0],
    [692, 730.0895512124052, 764.4838542466905, 813.506737945768, 0, 0]])
PERFREVIEW_T4_REVIEWER3 = np.array([[3, 3]])                                  # :214
DIVESTMENTS = np.array([[3, 3], [4, 2]])                                      # :217
SALE_VALUES_APPROX = np.array([                                               # :220-222 (isapprox)
    [-382.47934595588924, -0.0, -0.0, -382.0045082904202, -0.0],
    [-197.20899618910363, -0.0, -661.5797843033668, -196.96417421712303, -0.0]])
HOLDINGS_AFTER_LIQUIDATION = np.array([                                       # :228-230 (==)
    [0.0, 0.0, 0.0, 3.18, 0.0],
    [0.22199999999999995, 0.0, 0.6659999999999997, 0.22199999999999995, 0.0],
    [0.0, 0.0, 0.0, 0.0, 0.0]])
STAKES_AFTER_LIQUIDATION = np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 0, 0]])   # :233-234 (==)
PRICE_COL3_AFTER_SELL = np.array([105.41671691304936, 85.81763860422608, 117.8889457275222,
                                  109.76671150919375, 111.56926365234682])    # :264-265 (isapprox)
CASHOUT = np.array([744.5346623405912, 1014.8288665706393])                   # :270-271 (==)
CASHOUT_HARDCODED = np.array([746.434849083664, 1021.6613450347874])          # :274
INV_ASSETS_AFTER_CASH = np.array([[318.0, 0, 0, 0], [0, 111.0, 0, 0],
                                  [0, 0, 0, 746.434849083664], [0, 0, 0, 1021.6613450347874]])  # :277-280
# respawn! at t=3 (stale 5-arg call; used as an approximate pin only): one egg (fund 3),
# anchored by investor 3; the printed buy order implies 5 picks {1,1,3,4,4}
RESPAWN_BUY_VALUES = np.array([[298.574, 0.0, 149.287, 298.574, 0.0]])        # :298-299 atol 1e-4
RESPAWN_PICKS = [1, 1, 3, 4, 4]
RESPAWN_ANCHOR = 3
INV_ASSETS_AFTER_RESPAWN = np.array([[318.0, 0, 0, 0], [0, 111.0, 0, 0],
                                     [0, 0, 746.435, 0], [0, 0, 0, 1021.66]])  # :306-310 atol 0.01
STAKES_AFTER_RESPAWN = np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0]])      # :313-316
PRICE_COL3_AFTER_BUY = np.array([107.93469217989708, 85.81763860422608, 119.12089582037991,
                                 110.09444637041521, 111.56926365234682])     # :325-326 atol 1e-4
SHARES_OUT = np.array([2.7662468291692, 0.0, 1.25323939995470, 2.7119805752548, 0.0])  # :332 atol 1e-6
HOLDINGS_AFTER_SHARES = np.array([[0.0, 0.0, 0.0, 3.18, 0.0], [0.222, 0.0, 0.666, 0.222, 0.0],
                                  [2.7662468291692, 0.0, 1.25323939995470, 2.7119805752548, 0.0]])  # :338-340 atol 1e-5
FUND_VALUE_ROW3_AFTER_RESPAWN = np.array([673.1466804, 726.844707282, 746.434998099, 802.677606280])  # :349 atol 1e-3
# reinvest!(..., 3, "best") (:364-383)
INV_ASSETS_AFTER_REINVEST = np.array([[318.0, 0, 0, 0], [0, 111.0, 0, 0],
                                      [0, 0, 746.434849083664, 0], [1021.6613450347874, 0, 0, 0]])
STAKES_AFTER_REINVEST = np.array([[0.25575690228192466, 0, 0, 0.7442430977180754],
                                  [0, 1.0, 0, 0], [0, 0, 1.0, 0]])             # atol 1e-3
REINVEST_ORDER_VALUES = np.array([[0.0, 0.0, 0.0, 1021.6613450347874, 0.0]])  # :380 (==)
REINVEST_ORDER_FUNDS = [1]                                                    # :383
# marketmake! (:418-429), exact
MM_PRICES = np.array([[100.0, 101, 102], [100.0, 99, 98]])
MM_IMPACT = np.array([0.01, 0.05])
MM_BUY_ORDER, MM_BUY_RESULT = np.array([[20.0, 0]]), np.array([122.39999999999999, 98.0])
MM_SELL_ORDER, MM_SELL_RESULT = np.array([[-5.0, -5]]), np.array([96.89999999999999, 73.5])
# fundrevalue! (:434-455), atol 1
FR_HOLDINGS = np.array([[0.0, 0.0, 0.0, 3.18, 0.0], [0.222, 0.0, 0.666, 0.222, 0.0], [0.0] * 5])
FR_PRICES = np.array([[100.0, 107.903, 107.935, 119.085, 0.0, 0.0],
                      [100.0, 80.452, 85.8176, 86.4925, 0.0, 0.0],
                      [100.0, 118.683, 119.121, 126.537, 0.0, 0.0],
                      [100.0, 103.106, 110.094, 116.032, 0.0, 0.0],
                      [100.0, 103.284, 111.569, 112.856, 0.0, 0.0]])
FR_RESULT = np.array([[318.0, 327.876, 351.091, 368.983, 0.0, 0.0],
                      [110.999999999999, 125.886741903284, 127.736985080927, 136.469775818078, 0.0, 0.0],
                      [0.0] * 6])
