"""Item-item CF (BASELINE config 3; absent from the reference -> oracle = sklearn cosine_similarity on the item columns,
parity unpinned): every path of bb_item_item_topk -- shared-memory counting sort and its global-atomic fallback, single-
unit items and hot items split over several units, integer and float accumulators, several accumulator ranges -- at
the catalog size of config 3 and beyond.  Run on a B200: ``pytest -m gpu``."""
import numpy as np
import pytest
import scipy.sparse as sp

import brick_oracle as bo
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bb():
    import brickbrain_rec_engine_b200 as m
    return m


def _check_rows(indptr, indices, data, n_items, k, got_s, got_i, rows, rtol=2e-5):
    """Exact cosine rows of the item-item matrix for a subset of items (sklearn's arithmetic: normalize, then X^T X)."""
    from sklearn.preprocessing import normalize
    x = sp.csr_matrix((data, indices, indptr), shape=(len(indptr) - 1, n_items), dtype=np.float64).tocsc()
    xn = normalize(x, axis=0)
    sim = (xn[:, rows].T @ xn).toarray()                      # [len(rows), n_items]
    for r, i in enumerate(rows):
        s = sim[r].copy()
        s[i] = -1.0
        pos = np.nonzero(s > 0)[0]
        order = pos[np.lexsort((pos, -s[pos]))][:k]
        n_pos = len(order)
        gi, gs = got_i[i], got_s[i]
        np.testing.assert_allclose(gs[:n_pos], s[order], rtol=rtol, atol=1e-7, err_msg=f"item {i}")
        # ids must match wherever the neighbouring scores are not tied
        ref = s[order]
        gap = np.abs(np.diff(ref))
        ok = np.concatenate([[True], gap > 1e-6]) & np.concatenate([gap > 1e-6, [True]]) if n_pos > 1 else np.ones(n_pos, bool)
        if n_pos == k and len(pos) > k:                      # the k-th may tie with the first one cut off
            nxt = np.sort(s[pos])[::-1][k]
            ok[-1] &= abs(ref[-1] - nxt) > 1e-6
        assert np.array_equal(gi[:n_pos][ok], order[ok]), f"item {i}"
        if n_pos < k:                                        # zero-similarity fill: ascending ids, score 0
            fill = [j for j in range(n_items) if j != i and j not in set(order)][: k - n_pos]
            assert gi[n_pos:n_pos + len(fill)].tolist() == fill and np.all(gs[n_pos:n_pos + len(fill)] == 0)


def test_config3_shape_hot_and_cold_items(bb):
    """25,216 items (the reference's catalog), 150k Zipf users: the popular items have far more than one unit's worth
    of raters (partial rows + last-unit select), the tail has a handful (single-unit path)."""
    n_items, n_users, k = 25216, 150_000, 20
    indptr, indices, data = synth.zipf_ratings_csr(n_users, n_items, 20, 3)
    eng = bb.ScoringEngine(1, 1, 0)
    s, i = eng.item_item_topk(indptr, indices, data, n_items, k)
    deg = np.bincount(indices, minlength=n_items)
    assert deg.max() > 3 * 2048                               # several units for the hottest items
    order = np.argsort(-deg)
    rows = np.concatenate([order[:12], order[500:506], order[5000:5006], order[-12:], np.nonzero(deg == 0)[0][:2]]).astype(int)
    _check_rows(indptr, indices, data, n_items, k, s, i, rows)
    eng.close()


def test_float_ratings_use_the_float_accumulator(bb):
    n_items, n_users, k = 3000, 20_000, 16
    indptr, indices, data = synth.zipf_ratings_csr(n_users, n_items, 12, 5)
    data = (data + np.random.default_rng(1).random(data.shape[0]).astype(np.float32) * 0.5).astype(np.float32)
    eng = bb.ScoringEngine(1, 1, 0)
    s, i = eng.item_item_topk(indptr, indices, data, n_items, k)
    rows = np.arange(0, n_items, 97)
    _check_rows(indptr, indices, data, n_items, k, s, i, rows, rtol=2e-4)     # fp32 sums of non-integers: order dependent
    eng.close()


def test_wide_item_sets_take_the_fallback_sort_and_two_ranges(bb):
    """More items than the shared-memory histogram / accumulator hold (> 28k / > 53k): global-atomic CSC build and an
    accumulator that walks the item axis in ranges."""
    n_items, n_users, k = 70_000, 60_000, 10
    indptr, indices, data = synth.zipf_ratings_csr(n_users, n_items, 15, 9)
    eng = bb.ScoringEngine(1, 1, 0)
    s, i = eng.item_item_topk(indptr, indices, data, n_items, k)
    deg = np.bincount(indices, minlength=n_items)
    order = np.argsort(-deg)
    rows = np.concatenate([order[:6], order[3000:3004], order[-4:]]).astype(int)
    _check_rows(indptr, indices, data, n_items, k, s, i, rows)
    eng.close()


def test_users_without_ratings_and_unrated_items(bb):
    """Empty CSR rows between rated ones (the scatter finds a rating's user from row-start bitmaps: empty rows must not
    shift the owners), items nobody rated, fewer positive neighbours than k (zero-similarity fill)."""
    rng = np.random.default_rng(11)
    n_items, n_users, k = 700, 5000, 12
    lens = rng.integers(0, 9, size=n_users)
    lens[rng.random(n_users) < 0.4] = 0                        # 40% of the users rated nothing
    lens[:40] = 0
    lens[-3:] = 0
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.concatenate([np.sort(rng.choice(600, size=l, replace=False)) for l in lens if l] or [np.zeros(0, int)]).astype(np.int32)
    data = rng.integers(1, 6, size=indices.shape[0]).astype(np.float32)      # items 600..699 stay unrated
    eng = bb.ScoringEngine(1, 1, 0)
    s, i = eng.item_item_topk(indptr, indices, data, n_items, k)
    _check_rows(indptr, indices, data, n_items, k, s, i, np.arange(0, n_items, 7))
    eng.close()


def test_massive_ties_overflow_the_survivor_buffer(bb):
    """Every user rated every item with the same rating: all cosines are exactly 1, the selection bound admits the whole
    row (2999 survivors > the 1024-key buffer), so pass C repeats until the row is drained; ties go to the lower id."""
    n_items, n_users, k = 3000, 6, 24
    indptr = np.arange(n_users + 1, dtype=np.int64) * n_items
    indices = np.tile(np.arange(n_items, dtype=np.int32), n_users)
    data = np.full(indices.shape[0], 2.0, dtype=np.float32)
    eng = bb.ScoringEngine(1, 1, 0)
    s, i = eng.item_item_topk(indptr, indices, data, n_items, k)
    np.testing.assert_allclose(s, 1.0, rtol=1e-6)
    for row in (0, 1, 7, 1500, 2999):
        expect = [j for j in range(k + 1) if j != row][:k]
        assert i[row].tolist() == expect, f"item {row}"
    eng.close()


def test_mid_sized_item_sets_one_cta_per_sm_and_large_k(bb):
    """40k items: the accumulator (160 KB) leaves room for ONE CTA of 1024 threads per SM and still covers the catalog
    in one range; k = 128 is the selector's capacity."""
    n_items, n_users, k = 40_000, 80_000, 128
    indptr, indices, data = synth.zipf_ratings_csr(n_users, n_items, 18, 21)
    eng = bb.ScoringEngine(1, 1, 0)
    s, i = eng.item_item_topk(indptr, indices, data, n_items, k)
    deg = np.bincount(indices, minlength=n_items)
    order = np.argsort(-deg)
    rows = np.concatenate([order[:5], order[2000:2004], order[-4:]]).astype(int)
    _check_rows(indptr, indices, data, n_items, k, s, i, rows)
    eng.close()
