"""The oracle's two evaluation modes agree bit for bit (VERDICT r1 "weak" 3).

ORC_FAITHFUL_HISTORY is the reference's literal semantics: fundrevalue! rewrites the WHOLE K x T value history at
functions.jl:460, :479 and after every buy order (:260), and perfreview / horizonreturncalc read funds.value.  The default
("lazy") mode computes the same back-cast NAVs on demand from the price history (Q2).  Every GPU parity test compares with
the lazy oracle, so lazy == faithful is what ties those tests to the reference's own data flow."""
import numpy as np
import pytest

import helpers as H


def both(orc, p, seed, rep, selector, t_last=None):
    out = []
    for flags in (0, orc.FAITHFUL_HISTORY):
        w = H.oracle_world(orc, p, seed, rep)
        rc = w.modelrun(t_last=t_last, selector=selector, flags=flags)
        w.finalise_fund_value(t_last or p.bigt)
        out.append((rc, H.world_state(w), w.log(), w.trace(), np.array(w.nav_close), w.floor_hits()))
        w.close()
    return out


def assert_same(a, b):
    assert a[0] == b[0]
    for k in a[1]:
        assert H.eq(a[1][k], b[1][k]), k
    H.assert_log_equal(a[2], b[2])
    assert len(a[3]) == len(b[3])
    for f in ("t", "kind", "a", "b"):
        assert np.array_equal(a[3][f], b[3][f]), f
    assert H.eq(a[3]["v"], b[3]["v"]) and H.eq(a[4], b[4]) and a[5] == b[5]


@pytest.mark.parametrize("selector", ["probabilistic", "best", "goodenough", "random"])
@pytest.mark.parametrize("shape", [dict(), dict(bigk=11, bign=90, bigm=17, bigt=140, perfwindow=(1, 12), portfsizerange=(1, 9))],
                         ids=["k6", "k11"])
def test_lazy_equals_faithful_small_worlds(fsb, orc, selector, shape):
    p = H.small_params(fsb, **shape)
    for rep in range(3):
        lazy, faithful = both(orc, p, 99, rep, selector)
        assert_same(lazy, faithful)


def test_lazy_equals_faithful_default_shape(fsb, orc):
    """One default-shape replication (K = M = 250, N = 1000), 60 simulated days with the full K x T rewrite per call."""
    p = fsb.Params.default()
    lazy, faithful = both(orc, p, 20261020, 1, "probabilistic", t_last=p.perfwindow[-1] + 1 + 60)
    assert_same(lazy, faithful)
