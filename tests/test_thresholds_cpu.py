"""CPU check of the arithmetic behind the selection kernel's grouped thresholds (csrc/select_tc.cu, refresh_grouped):
a numpy restatement of pair_low3 / quad_bound / the share-out of the 33 keys over the groups, checked against sorting.

The bound must (1) never exceed the true 33rd best score of the keys seen so far -- no top-k (k <= 32) member is ever
filtered out -- and (2) never fall below the plain minimum of the published 2nd-best scores it replaced."""
import numpy as np
import pytest


def pair_low3(u1, u2, v1, v2):
    """the three smallest of a pair's four published values, ascending (select_tc.cu pair_low3)"""
    u1, v1 = np.maximum(u1, u2), np.maximum(v1, v2)
    s, S, p = np.minimum(u2, v2), np.maximum(u2, v2), np.minimum(u1, v1)
    return s, np.minimum(S, p), np.maximum(S, p)


def quad_bound(x, y, want):
    """want-th largest of a group's 8 values = (9 - want)-th smallest, want in {6, 7, 8} (select_tc.cu quad_bound)"""
    if want >= 8:
        return np.minimum(x[0], y[0])
    if want == 7:
        return np.minimum(np.minimum(x[1], y[1]), np.maximum(x[0], y[0]))
    return np.minimum(np.minimum(x[2], y[2]), np.minimum(np.maximum(x[0], y[1]), np.maximum(x[1], y[0])))


def share_out(vsplits):
    """keys each group vouches for (select_tc.cu refresh_grouped): quads 6 (+ up to 2), the trailing pair 3 (+ 1)"""
    nq, np_ = vsplits >> 2, (vsplits >> 1) & 1
    deficit = max(0, 33 - (6 * nq + 3 * np_))
    pair_want = 3 + (1 if np_ and deficit > 0 else 0)
    deficit -= pair_want - 3
    q_add, q_rem = (deficit // nq, deficit % nq) if nq else (0, 0)
    return [6 + q_add + (1 if qi < q_rem else 0) for qi in range(nq)], (pair_want if np_ else 0)


def grouped_bound(b1, b2):
    """b1, b2: [virtual splits] published best / 2nd best (any of them may be -inf: not published yet)"""
    v = len(b1)
    wants, pair_want = share_out(v)
    bound = np.inf
    for qi, want in enumerate(wants):
        i = 4 * qi
        x = pair_low3(b1[i], b2[i], b1[i + 1], b2[i + 1])
        y = pair_low3(b1[i + 2], b2[i + 2], b1[i + 3], b2[i + 3])
        bound = min(bound, quad_bound(x, y, want))
    if pair_want:
        x = pair_low3(b1[v - 2], b2[v - 2], b1[v - 1], b2[v - 1])
        bound = min(bound, x[0] if pair_want >= 4 else x[1])
    return bound


@pytest.mark.parametrize('vsplits', list(range(18, 33, 2)))
def test_share_out_reaches_33_keys(vsplits):
    wants, pair_want = share_out(vsplits)
    assert sum(wants) + pair_want >= 33
    assert all(6 <= w <= 8 for w in wants) and pair_want in (0, 3, 4)
    assert len(wants) * 4 + (2 if pair_want else 0) == vsplits


def test_group_formulas_against_sorting():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        vals = rng.standard_normal((4, 2))
        vals.sort(axis=1)
        b2, b1 = vals[:, 0].copy(), vals[:, 1].copy()
        if rng.random() < 0.2:          # a cut published an exact 2nd best above the tracker's best: the clamp's case
            b1[0] = b2[0] - 1.0
        eff1 = np.maximum(b1, b2)
        x = pair_low3(b1[0], b2[0], b1[1], b2[1])
        y = pair_low3(b1[2], b2[2], b1[3], b2[3])
        assert list(x) == sorted([eff1[0], b2[0], eff1[1], b2[1]])[:3]
        everything = sorted(np.concatenate([eff1, b2]), reverse=True)
        for want in (6, 7, 8):
            assert quad_bound(x, y, want) == everything[want - 1]


@pytest.mark.parametrize('vsplits', [18, 20, 22, 24, 26, 32])
def test_grouped_bound_is_valid_and_not_looser_than_the_minimum(vsplits):
    rng = np.random.default_rng(vsplits)
    keys_per_split = 40
    for trial in range(300):
        scores = rng.standard_normal((vsplits, keys_per_split)) * rng.uniform(0.1, 3.0)
        if trial % 7 == 0:
            scores = np.round(scores, 1)                      # plenty of exact ties
        seen = rng.integers(0, keys_per_split + 1, size=vsplits)        # keys each split has scanned so far
        if trial % 3:
            seen = np.maximum(seen, 2)
        b1 = np.full(vsplits, -np.inf)
        b2 = np.full(vsplits, -np.inf)
        pool = []
        for v in range(vsplits):
            s = np.sort(scores[v, :seen[v]])[::-1]
            pool += list(s)
            if len(s) >= 1:
                b1[v] = s[0]
            if len(s) >= 2:
                b2[v] = s[1]
            if trial % 5 == 0 and len(s) >= 2 and rng.random() < 0.3:
                b1[v] = -np.inf                                 # the best score's store has not landed yet
        bound = grouped_bound(b1, b2)
        assert bound >= b2.min()                               # (2) at least as tight as the plain minimum
        if bound > -np.inf:
            assert len(pool) >= 33
            true33 = np.sort(np.array(pool))[::-1][32]
            assert bound <= true33, (trial, bound, true33)     # (1) 33 keys seen so far score at or above the bound
