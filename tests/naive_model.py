"""A second, deliberately NAIVE restatement of the reference's simulation step, independent of oracle/fsb_oracle.c.

TEST INFRASTRUCTURE.  Written from /root/reference/src/functions.jl alone (line numbers cited per function), with the
reference's own data structures: dense numpy matrices in Julia's shapes (`stocks.value` M x T, `investors.assets`
N x (K+1), `funds.holdings` K x M, `funds.stakes` K x N, `funds.value` K x T), order tables grown row by row,
`fundrevalue!` over the WHOLE history with a plain matrix product (numpy/BLAS summation order, not the oracle's canonical
chains), 1-based ids inside order tables and the log, exactly as the Julia code has them.  Nothing is shared with the
oracle but the draws: exogenous shocks come from dense tapes, respawn anchors / portfolio draws from an event tape, and
the selectors' uniforms from a callback (the Philox words of the oracle's draw site) -- the transliteration samples its
own fund choices from them.

Its purpose ("Next round" 2c): the reference has no integration test of `modelrun`
(test/functions_test.jl:458-465), so the oracle's *orchestration* -- call order, which array is read when, the quirks
Q1-Q15 of SURVEY.md 3.4 -- is pinned by nothing.  Two independent restatements that agree to 1e-12 on shared draws are
the best available defence.
"""
from __future__ import annotations

import math

import numpy as np


class Q6Error(Exception):
    """functions.jl:282 -- `shuffle(potentialinvs)[1:neggs]` throws a BoundsError when dead funds outnumber cash holders."""


def _isless_key(x):
    """Julia's isless on Float64 as a sort key: -0.0 < 0.0, NaN after everything, NaNs equal."""
    if x != x:
        return (1, 0.0, 0)
    return (0, x, 0 if math.copysign(1.0, x) < 0 else 1)


def sortperm(v):
    """Base.sortperm: stable, ascending under isless (functions.jl:404)."""
    return sorted(range(len(v)), key=lambda i: _isless_key(v[i]))


def findmax_index(v):
    """Base.findmax: the first NaN if any, else the first maximum under isless (functions.jl:382, 390)."""
    best = 0
    for i in range(1, len(v)):
        a, b = v[i], v[best]
        if b != b:
            break
        if a != a or _isless_key(b) < _isless_key(a):
            best = i
            if a != a:
                break
    return best


class NaiveWorld:
    def __init__(self, state, *, W, drift, marketvol, reviewprobability, dense, events, choice_words=None,
                 forced_choices=None):
        """state: dict of the reference's arrays after `initialise` (copied); dense: z_mkt[T], z_stock[M,T],
        u_review[N,T]; events: {(kind, t, a, b): value} with kind in 'anchor', 'rpsize', 'rppick' (0-based ids);
        choice_words(t, inv) -> the four 32-bit words of the selector's draw; forced_choices: {(t, inv): fund}."""
        self.market = np.array(state["market"], dtype=np.float64)
        self.value = np.array(state["stock_value"], dtype=np.float64, order="F")
        self.beta = np.array(state["beta"], dtype=np.float64)
        self.vol = np.array(state["vol"], dtype=np.float64)
        self.impact = np.array(state["impact"], dtype=np.float64)
        self.volume = np.array(state["volume"], dtype=np.float64, order="F")
        self.assets = np.array(state["inv_assets"], dtype=np.float64, order="F")
        self.horizon = np.array(state["horizon"], dtype=np.int64)
        self.threshold = np.array(state["threshold"], dtype=np.float64)
        self.holdings = np.array(state["holdings"], dtype=np.float64, order="F")
        self.stakes = np.array(state["stakes"], dtype=np.float64, order="F")
        self.fundvalue = np.array(state["fund_value"], dtype=np.float64, order="F")
        self.K, self.M = self.holdings.shape
        self.N = self.assets.shape[0]
        self.T = self.value.shape[1]
        self.W, self.drift, self.marketvol, self.reviewprobability = W, drift, marketvol, reviewprobability
        self.dense, self.events = dense, events
        self.choice_words, self.forced_choices = choice_words, forced_choices
        self.log = None
        self.trace = []  # (t, kind, a, b) discrete outcomes, 0-based ids
        self.audit = []  # decision margins: ('review', t, rev, fund, valchange, threshold), ('choice', t, inv, returns, u)
        self.floor_hits = 0

    # ---- price functions ---------------------------------------------------------------------------------
    def marketmove(self, t):  # functions.jl:13-17 (Q13: the source, not the stale test)
        z = self.dense["z_mkt"][t - 1]
        self.market[t - 1] = self.market[t - 2] * (1 + self.drift + self.marketvol * z)

    def stockmove(self, t):  # :29-34
        mktreturn = (self.market[t - 1] - self.market[t - 2]) / self.market[t - 2]
        z = self.dense["z_stock"][:, t - 1]
        prev = self.value[:, t - 2]
        self.value[:, t - 1] = ((1 + mktreturn * self.beta) * prev) + self.vol * z * prev

    def fundrevalue(self, targets):  # :266-271 -- the whole history row, current holdings (Q2)
        for k in targets:
            self.fundvalue[k, :] = self.holdings[k, :] @ self.value

    def marketmake(self, t, ordervals):  # :205-211 (Q10)
        netimpact = ordervals.sum(axis=0) * self.impact
        self.value[:, t - 1] = (1 + netimpact) * self.value[:, t - 1]
        nearzero = np.nonzero(self.value[:, t - 1] < 0.000000001)[0]
        self.value[nearzero, t - 1] = 0.000000001
        self.floor_hits += len(nearzero)

    # ---- agent behaviours ----------------------------------------------------------------------------------
    def drawreviewers(self, t):  # :147-151, ids 1-based
        return [n + 1 for n in np.nonzero(self.dense["u_review"][:, t - 1] < self.reviewprobability)[0]]

    def perfreview(self, t, reviewers):  # :157-176 (Q1: denominator column `horizon`)
        divestments = np.zeros((1, 2), dtype=np.int64)
        for rev in reviewers:
            horizon = int(self.horizon[rev - 1])
            fnds = [f + 1 for f in np.nonzero(self.assets[rev - 1, :-1] > 0)[0]]
            for fnd in fnds:
                fv = self.fundvalue
                valchange = (fv[fnd - 1, t - 1] - fv[fnd - 1, t - horizon - 1]) / fv[fnd - 1, horizon - 1]
                self.audit.append(("review", t, rev - 1, fnd - 1, float(valchange), float(self.threshold[rev - 1])))
                if valchange < self.threshold[rev - 1]:
                    pair = np.array([[rev, fnd]], dtype=np.int64)
                    divestments = pair if divestments[0, :].sum() == 0 else np.vstack([divestments, pair])
        return divestments

    def liquidate(self, t, divestments):  # :183-201
        values = np.zeros((0, self.M))
        investors = []
        stockvals = self.value[:, t - 1]
        for row in range(divestments.shape[0]):
            investor, fund = int(divestments[row, 0]), int(divestments[row, 1])
            stake = self.stakes[fund - 1, investor - 1]
            salevals = -stake * (self.holdings[fund - 1, :] * stockvals)
            values = np.vstack([values, salevals[None, :]])
            investors.append(investor)
            remainder = 1 - stake
            self.holdings[fund - 1, :] = self.holdings[fund - 1, :] * remainder
            self.stakes[fund - 1, investor - 1] = 0
            tot = self.stakes[fund - 1, :].sum()
            if tot > 0:
                self.stakes[fund - 1, :] = self.stakes[fund - 1, :] / tot
        return values, investors

    def trackvolume_sell(self, t, sellvalues):  # :555-558
        self.volume[:, t - 1] -= sellvalues.sum(axis=0)

    def executeorder_sell(self, t, values, investors):  # :213-224
        old = self.value.copy()
        self.marketmake(t, values)
        priceratio = self.value[:, t - 1] / old[:, t - 1]
        return [(investors[o], -np.sum(values[o, :] * priceratio)) for o in range(values.shape[0])]

    def executeorder_buy(self, t, values, funds):  # :229-239
        self.marketmake(t, values)
        return [(funds[o], values[o, :] / self.value[:, t - 1]) for o in range(values.shape[0])]

    def disburse_cash(self, divestments, cashout):  # :243-252 (Q8)
        for row in range(divestments.shape[0]):
            inv, fund = int(divestments[row, 0]), int(divestments[row, 1])
            self.assets[inv - 1, fund - 1] = 0
            self.assets[inv - 1, -1] = cashout[row][1]

    def disburse_shares(self, sharesout):  # :256-263
        for fund, shares in sharesout:
            self.holdings[fund - 1, :] += shares
            self.fundrevalue([fund - 1])

    def calc_liquidity(self, k, col):  # :324-330
        weights = self.holdings[k, :] * self.value[:, col - 1]
        weights = weights / weights.sum()
        return float((weights * self.impact).sum())

    def respawn(self, t):  # :275-321 (Q5, Q6, Q15)
        values, funds = np.zeros((0, self.M)), []
        spawn = [k + 1 for k in np.nonzero(self.holdings.sum(axis=1) == 0)[0]]
        neggs = len(spawn)
        potentialinvs = [n + 1 for n in np.nonzero(self.assets[:, -1] > 0)[0]]
        if neggs > len(potentialinvs):
            raise Q6Error(f"t={t}: {neggs} dead funds, {len(potentialinvs)} cash holders")
        # Random.shuffle(potentialinvs)[1:neggs]: the tape holds the first neggs entries of the shuffled list
        anchorinvs = [int(self.events[("anchor", t, i, 0)]) + 1 for i in range(neggs)]
        assert all(a in potentialinvs for a in anchorinvs) and len(set(anchorinvs)) == neggs
        for idx in range(neggs):
            inv, egg = anchorinvs[idx], spawn[idx]
            capital = self.assets[inv - 1, -1]
            self.assets[inv - 1, egg - 1] = capital
            self.assets[inv - 1, -1] = 0
            # fundholdinit! (:119-133) on a COPY of the egg's row: nofm = rand(portfsizerange, 1), picks with replacement
            row = self.holdings[egg - 1, :].copy()
            nofm = int(self.events[("rpsize", t, idx, 0)])
            selection = [int(self.events[("rppick", t, idx, jj)]) + 1 for jj in range(nofm)]
            stockvals = self.value[:, t - 1]
            for stock in selection:
                row[stock - 1] += (capital / len(selection)) / stockvals[stock - 1]
            buyorder = row * stockvals
            values = np.vstack([values, buyorder[None, :]])
            funds.append(egg)
            self.stakes[egg - 1, inv - 1] = 1
            self.trace.append((t, "spawn", egg - 1, inv - 1))
            # log chain (:298-309)
            new, old, chkval = egg, None, 1
            while chkval == 1:
                old = new
                chkval = self.log["liquidated"][old - 1]
                new = self.log["replacedby"][old - 1]
            self.log["liquidated"][old - 1] = 1
            self.log["replacedby"][old - 1] = len(self.log["liquidated"]) + 1
            self.log["failtime"][old - 1] = t
            fundsize = buyorder.sum()
            nstocks = int((buyorder > 0).sum())
            weights = buyorder / fundsize
            fundliquidity = float((weights * self.impact).sum())
            for col, v in zip(("liquidated", "replacedby", "initialsize", "liquidity", "failtime", "ninvestors", "nstocks"),
                              (0, 0, fundsize, fundliquidity, 0, 1, nstocks)):
                self.log[col].append(v)
        return values, funds

    # selectors, :381-415 -------------------------------------------------------------------------------------
    def select(self, selection, horizonreturns, t, reinv):
        if self.forced_choices is not None:
            return int(self.forced_choices[(t, reinv - 1)]) + 1
        K = len(horizonreturns)
        if selection == "best":
            return findmax_index(horizonreturns) + 1
        w = self.choice_words(t, reinv - 1)
        if selection == "random":  # rand(1:K): the draw site maps its third word to 0..K-1 by multiply-high
            return ((w[2] * K) >> 32) + 1
        if selection == "goodenough":
            candidates = [i + 1 for i in range(K) if horizonreturns[i] > self.threshold[reinv - 1]]
            if not candidates:
                return findmax_index(horizonreturns) + 1
            return candidates[(w[2] * len(candidates)) >> 32]
        if selection == "probabilistic":  # Q14
            normaliser = sum(range(1, K + 1))
            order = sortperm(list(horizonreturns))
            probability = [0.0] * K
            for rank, idx in enumerate(order):
                probability[idx] = (rank + 1) / normaliser
            # rand(Categorical(p)), Distributions v0.20: cp = 0; i = 0; while cp < u && i < n: cp += p[i += 1]
            u = ((((w[0] << 32) | w[1]) >> 12) + 0.5) * 2.0 ** -52
            cp, i = 0.0, 0
            while cp < u and i < K:
                i += 1
                cp += probability[i - 1]
            return max(i, 1)
        raise ValueError("Please specify a valid fund selection method in reinvest!")

    def reinvest(self, t, selection):  # :335-373 (Q7, Q12)
        values, funds = np.zeros((0, self.M)), []
        reinvestors = [n + 1 for n in np.nonzero(self.assets[:, -1] > 0)[0]]
        for reinv in reinvestors:
            injection = self.assets[reinv - 1, -1]
            h = int(self.horizon[reinv - 1])
            fv = self.fundvalue
            with np.errstate(all="ignore"):
                horizonreturns = (fv[:, t - 1] - fv[:, t - h - 1]) / fv[:, t - h - 1]  # :375-379
            choice = self.select(selection, horizonreturns, t, reinv)
            self.audit.append(("choice", t, reinv - 1, [float(x) for x in horizonreturns]))
            self.trace.append((t, "choice", reinv - 1, choice - 1))
            self.assets[reinv - 1, choice - 1] = injection
            self.assets[reinv - 1, -1] = 0
            fundval = self.holdings[choice - 1, :] @ self.value[:, t - 1]
            with np.errstate(all="ignore"):
                stakevals = self.stakes[choice - 1, :] * fundval
                stakevals[reinv - 1] = injection
                newcapital = fundval + injection
                self.stakes[choice - 1, :] = stakevals / newcapital
                portfoliovalues = self.holdings[choice - 1, :] * self.value[:, t - 1]
                buyorder = (portfoliovalues / fundval) * injection
            values = np.vstack([values, buyorder[None, :]])
            funds.append(choice)
        return values, funds

    # ---- modelrun, :446-491 ---------------------------------------------------------------------------------
    def build_log(self):  # :448-456 (Q4: sized with the number of FUNDS)
        K = self.K
        with np.errstate(all="ignore"):
            liq = [self.calc_liquidity(k, 1) for k in range(K)]
        self.log = dict(liquidated=[0] * K, replacedby=[0] * K, initialsize=list(self.fundvalue[:, 0]), liquidity=liq,
                        failtime=[0] * K, ninvestors=list((self.stakes > 0).sum(axis=1)),
                        nstocks=list((self.holdings > 0).sum(axis=1)))

    def modelrun(self, t_first, t_last, selection="probabilistic"):
        if self.log is None:
            self.build_log()
        K = self.K
        for t in range(t_first, t_last + 1):
            self.marketmove(t)
            self.stockmove(t)
            self.fundrevalue(range(K))
            reviewers = self.drawreviewers(t)
            divestments = self.perfreview(t, reviewers)
            if divestments[0, :].sum() > 0:
                for row in range(divestments.shape[0]):
                    self.trace.append((t, "divest", int(divestments[row, 0]) - 1, int(divestments[row, 1]) - 1))
                sellvalues, sellinvestors = self.liquidate(t, divestments)
                self.trackvolume_sell(t, sellvalues)
                cashout = self.executeorder_sell(t, sellvalues, sellinvestors)
                self.disburse_cash(divestments, cashout)
                if np.any(self.stakes.sum(axis=1) == 0):
                    buyvalues, buyfunds = self.respawn(t)
                    self.trackvolume_sell(t, sellvalues)  # (sic, Q3)
                    sharesout = self.executeorder_buy(t, buyvalues, buyfunds)
                    self.disburse_shares(sharesout)
                self.fundrevalue(range(K))
                if np.any(self.assets[:, -1] > 0):
                    buyvalues, buyfunds = self.reinvest(t, selection)
                    self.trackvolume_sell(t, sellvalues)  # (sic, Q3)
                    sharesout = self.executeorder_buy(t, buyvalues, buyfunds)
                    self.disburse_shares(sharesout)
        return self.log
