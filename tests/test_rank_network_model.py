"""A numpy model of the CTA-wide bitonic network of block_rank_sort_reg (csrc/fsb_step.cuh): thread tid owns the positions
8 tid .. 8 tid + 7; per merge size k2 the strides >= 8 exchange slot r with thread tid ^ (stride / 8) (shuffles inside a warp,
shared memory across warps: the same partner formula), the strides 4, 2, 1 stay inside a thread.  The model follows the
kernel's index and direction expressions literally and must sort any input into the stable ascending order of (key, fund
index) -- ties, padding and already sorted / reversed inputs included.  (The GPU tests compare the kernel itself with the
oracle; this pins the network's index logic where no GPU is needed.)"""
import numpy as np
import pytest


def network(keys, K, NT):
    P = 8 * NT
    kk = np.full((NT, 8), np.iinfo(np.uint64).max, dtype=np.uint64)
    ii = np.zeros((NT, 8), dtype=np.int64)
    for tid in range(NT):
        for r in range(8):
            e = tid * 8 + r
            if e < K:
                kk[tid, r] = keys[e]
            ii[tid, r] = e
    tids = np.arange(NT)

    def less(ka, ia, kb, ib):
        return (ka < kb) | ((ka == kb) & (ia < ib))

    k2 = 2
    while k2 <= P:
        up_t = ((tids * 8) & k2) == 0
        tm = k2 >> 4
        while tm > 0:
            keepmin = ((tids & tm) == 0) == up_t
            partner = tids ^ tm
            ok, oi = kk[partner].copy(), ii[partner].copy()
            take = less(ok, oi, kk, ii) == keepmin[:, None]
            kk, ii = np.where(take, ok, kk), np.where(take, oi, ii)
            tm >>= 1
        for j2 in (4, 2, 1):
            if j2 < k2:
                for r in range(8):
                    if (r & j2) == 0:
                        up = up_t if k2 >= 8 else np.full(NT, (r & k2) == 0)
                        sw = less(kk[:, r | j2], ii[:, r | j2], kk[:, r], ii[:, r]) == up
                        xk, yk, xi, yi = kk[:, r].copy(), kk[:, r | j2].copy(), ii[:, r].copy(), ii[:, r | j2].copy()
                        kk[:, r], kk[:, r | j2] = np.where(sw, yk, xk), np.where(sw, xk, yk)
                        ii[:, r], ii[:, r | j2] = np.where(sw, yi, xi), np.where(sw, xi, yi)
        k2 <<= 1
    rank = np.zeros(K, dtype=np.int64)
    for tid in range(NT):
        for r in range(8):
            if ii[tid, r] < K:
                rank[ii[tid, r]] = tid * 8 + r + 1
    return rank


@pytest.mark.parametrize("NT,K", [(128, 600), (128, 1024), (256, 1100), (256, 2000), (256, 2048), (32, 200)])
def test_register_network_gives_the_stable_ranks(NT, K):
    rng = np.random.default_rng(NT + K)
    cases = [rng.integers(0, 2 ** 63, K, dtype=np.uint64),
             rng.integers(0, 7, K, dtype=np.uint64),                       # many ties: the fund index decides
             np.arange(K, dtype=np.uint64), np.arange(K, dtype=np.uint64)[::-1].copy(),
             np.full(K, np.iinfo(np.uint64).max, dtype=np.uint64)]         # real keys equal to the padding value
    for keys in cases:
        want = np.empty(K, dtype=np.int64)
        want[np.argsort(keys, kind="stable")] = np.arange(1, K + 1)
        assert np.array_equal(network(keys, K, NT), want)
