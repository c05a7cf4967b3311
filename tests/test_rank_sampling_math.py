"""The arithmetic behind the integer shortcut of the rank-proportional fund choice (csrc/fsb_step.cuh:
select_fund_prob / select_fund_ranked).  The reference (src/functions.jl:400-410, Q14) draws from Categorical(p) with
p_k = rank_k / S, S = K (K + 1) / 2, by the sequential scan `cp += p[i]` until cp >= u.  The kernels instead bisect the exact
integer prefix sums C_i of the ranks for the smallest i with C_i >= u S, unless u S lies within a guard of an integer.  This
test pins, on the CPU and in plain numpy, what that relies on: the scan's partial sums stay within (K + 1) 2^-53 S of C_i in
units of C, the guard the kernels use is larger than that at both shapes, and outside the guard both rules pick the same fund."""
import numpy as np
import pytest


def guard(K):
    S = 0.5 * K * (K + 1)
    return max(1e-6, 4.0 * (K + 1) * 2.0 ** -53 * S)


@pytest.mark.parametrize("K", [250, 2000])
def test_scan_partial_sums_stay_within_the_bound(K):
    rng = np.random.default_rng(K)
    S = 0.5 * K * (K + 1)
    worst = 0.0
    for _ in range(20):
        ranks = rng.permutation(K) + 1
        p = ranks.astype(np.float64) / S            # ranktab[rank - 1]: one rounding per term
        cp = np.cumsum(p)                           # numpy's cumsum adds sequentially: the scan's partial sums
        C = np.cumsum(ranks.astype(np.int64))
        worst = max(worst, float(np.max(np.abs(cp * S - C))))
    bound = (K + 1) * 2.0 ** -53 * S
    assert worst <= bound
    assert guard(K) >= 2.0 * bound + 1e-9           # (room for the rounding of u * S itself, < 1e-9 at these sizes)


@pytest.mark.parametrize("K", [250, 2000])
def test_bisection_equals_the_scan_outside_the_guard(K):
    rng = np.random.default_rng(7 * K)
    S = 0.5 * K * (K + 1)
    g = guard(K)
    checked = 0
    for _ in range(8):
        ranks = rng.permutation(K) + 1
        p = ranks.astype(np.float64) / S
        cp = np.cumsum(p)
        C = np.cumsum(ranks.astype(np.int64))
        us = np.concatenate([rng.random(4000), (C[rng.integers(0, K, 200)] + rng.choice([-3e-6, 3e-6], 200)) / S])
        for u in us:
            x = u * S
            if abs(x - np.rint(x)) <= g or not (0.0 < u < 1.0):
                continue
            # the reference's scan: first i with cp_i >= u (the last fund if none, as the kernel's loop ends at K)
            hit = np.nonzero(cp >= u)[0]
            scan = int(hit[0]) if len(hit) else K - 1
            lo, hi = 0, K - 1
            while lo < hi:                          # select_fund_ranked's bisection over the integer prefix sums
                mid = (lo + hi) >> 1
                if float(C[mid]) >= x:
                    hi = mid
                else:
                    lo = mid + 1
            assert lo == scan, (K, u)
            checked += 1
    assert checked > 30000
