"""The oracle restatement vs the golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

import brick_oracle as bo


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _assert_ranked_equal(ids_ref, sc_ref, ids, sc, rtol=1e-6, what=""):
    """ids identical where the reference's adjacent score gap exceeds summation noise."""
    n_ref = int((ids_ref >= 0).sum())
    assert int((np.asarray(ids) >= 0).sum()) == n_ref, what
    s_ref = np.asarray(sc_ref[:n_ref], dtype=np.float64)
    np.testing.assert_allclose(np.asarray(sc[:n_ref], dtype=np.float64), s_ref, rtol=rtol, atol=1e-7, err_msg=what)
    gaps = np.abs(np.diff(s_ref))
    tie_free = np.concatenate([[True], gaps > 1e-6]) & np.concatenate([gaps > 1e-6, [True]]) if n_ref > 1 else np.ones(n_ref, bool)
    assert np.array_equal(np.asarray(ids[:n_ref])[tie_free], ids_ref[:n_ref][tie_free]), what


def test_content_fp32(golden_dir):
    g = _load(golden_dir, "content_fp32_golden.npz")
    feat = g["q8"].astype(np.float32) / g["scale"]
    k = int(g["k"])
    valid = np.zeros(feat.shape[0], dtype=bool)
    valid[g["valid_ids"]] = True
    for j, t in enumerate(g["targets"]):
        ids, sc = bo.content_similar(feat, int(t), k)
        _assert_ranked_equal(g["ids_nofilter"][j], g["scores_nofilter"][j], ids, sc, what=f"target {t}")
        ids, sc = bo.content_similar(feat, int(t), k, valid_mask=valid)
        _assert_ranked_equal(g["ids_filter"][j], g["scores_filter"][j], ids, sc, what=f"target {t} filtered")
        # engine spec (exclude self) == reference (drop rank 0) when the target is the unique max
        ids2, _ = bo.content_similar(feat, int(t), k, drop="self")
        assert np.array_equal(ids2, bo.content_similar(feat, int(t), k)[0])


def test_content_f64(golden_dir):
    g = _load(golden_dir, "content_f64_golden.npz")
    for j, t in enumerate(g["targets"]):
        ids, sc = bo.content_similar(g["feat"], int(t), int(g["k"]))
        _assert_ranked_equal(g["ids"][j], g["scores"][j], ids, sc, rtol=1e-12, what=f"target {t}")


def test_faiss_fixture(golden_dir):
    g = _load(golden_dir, "faiss_fixture_golden.npz")
    emb = g["emb"]
    assert emb.shape == (10, 384)
    np.testing.assert_allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-6)
    for t in range(10):
        ids, sc = bo.content_similar(emb, t, 5)
        _assert_ranked_equal(g["ids"][t], g["scores"][t], ids, sc, what=f"fixture row {t}")
        # FAISS IndexFlatL2 ranking on unit vectors == cosine ranking (2 - 2cos)
        ids_l2, d = bo.vector_search(emb[t], emb, 6, metric="l2")
        assert ids_l2[0] == t
        assert np.array_equal(ids_l2[1:], ids)
        np.testing.assert_allclose(d[1:], 2.0 - 2.0 * sc, atol=2e-6)


def test_cf(golden_dir):
    g = _load(golden_dir, "cf_golden.npz")
    uf, itf = g["user_factors"], g["item_factors"]
    ratings = g["ratings"]
    user_row = {int(u): i for i, u in enumerate(g["user_ids"])}
    valid = np.zeros(itf.shape[0], dtype=bool)
    valid[g["valid_ids"]] = True
    for j, u in enumerate(g["users"]):
        r = user_row[int(u)]
        ids, sc = bo.cf_recommend(uf[r], itf, int(g["k"]), rated_mask=ratings[r] > 0)
        _assert_ranked_equal(g["ids_nofilter"][j], g["scores_nofilter"][j], ids, sc, rtol=1e-12, what=f"user {u}")
        ids, sc = bo.cf_recommend(uf[r], itf, int(g["k"]), rated_mask=ratings[r] > 0, valid_mask=valid)
        _assert_ranked_equal(g["ids_filter"][j], g["scores_filter"][j], ids, sc, rtol=1e-12, what=f"user {u} filtered")


def test_masks(golden_dir):
    g = _load(golden_dir, "mask_golden.npz")
    themes = {int(k): v for k, v in json.loads(str(g["themes_json"])).items()}
    owned = {int(k): np.array(v, dtype=np.int64) for k, v in json.loads(str(g["owned_json"])).items()}
    wish = {int(k): np.array(v, dtype=np.int64) for k, v in json.loads(str(g["wish_json"])).items()}
    specs = json.loads(str(g["scenario_specs"]))
    off = g["valid_off"]
    for j, name in enumerate(g["scenario_names"]):
        pairs = [(t, v) for t, v in specs[j]["pairs"]]
        spec = bo.compile_constraints(pairs, themes, int(g["now_year"]))
        mask = bo.constraint_mask(spec, g["num_parts"], g["year"], g["theme_id"], owned, wish)
        expect = g["valid_flat"][off[j]:off[j + 1]]
        assert np.array_equal(np.nonzero(mask)[0], expect), f"scenario {name}"


def test_hybrid(golden_dir):
    g = _load(golden_dir, "hybrid_golden.npz")
    comb = json.loads(str(g["combine_json"]))
    ids, sc = bo.hybrid_combine(comb["content_ids"], comb["content_scores"], comb["cf_ids"], comb["cf_scores"],
                                0.4, 0.6, int(g["k"]))
    _assert_ranked_equal(np.array(comb["out_ids"]), np.array(comb["out_scores"]), ids, sc, rtol=1e-12)
    w = bo.normalize_weights(0.7, 0.5)
    assert w == (0.7 / 1.2, 0.5 / 1.2)


def test_hybrid_end_to_end(golden_dir):
    """HybridRecommender.get_recommendations restated from oracle parts."""
    g = _load(golden_dir, "hybrid_golden.npz")
    gc = _load(golden_dir, "content_fp32_golden.npz")
    gf = _load(golden_dir, "cf_golden.npz")
    feat = gc["q8"].astype(np.float32) / gc["scale"]
    themes = {int(k): v for k, v in json.loads(str(g["themes_json"])).items()}
    owned = {int(k): np.array(v, dtype=np.int64) for k, v in json.loads(str(g["owned_json"])).items()}
    wish = {int(k): np.array(v, dtype=np.int64) for k, v in json.loads(str(g["wish_json"])).items()}
    user_row = {int(u): i for i, u in enumerate(gf["user_ids"])}
    k = int(g["k"])
    type_of = dict(price_max="price_max", price_min="price_min", pieces_max="pieces_max", pieces_min="pieces_min",
                   age_min="age_min", age_max="age_max", year_min="year_min", year_max="year_max",
                   required_themes="themes_required", excluded_themes="themes_excluded",
                   max_complexity="complexity_max", min_complexity="complexity_min")
    for case in json.loads(str(g["cases_json"])):
        kw = case["kwargs"]
        valid = None
        if kw:
            pairs = [(type_of[a], v) for a, v in kw.items() if a in type_of]
            if kw.get("exclude_owned") and kw.get("user_id"):
                pairs.append(("exclude_owned", kw["user_id"]))
            if kw.get("exclude_wishlisted") and kw.get("user_id"):
                pairs.append(("exclude_wishlisted", kw["user_id"]))
            spec = bo.compile_constraints(pairs, themes, int(g["now_year"]))
            valid = bo.constraint_mask(spec, g["num_parts"], g["year"], g["theme_id"], owned, wish)
            assert int(valid.sum()) == case["n_valid"]
            if not valid.any():
                assert all(i < 0 for i in case["ids"])
                continue
        c_ids = c_sc = f_ids = f_sc = []
        if case["liked"] is not None:
            c_ids, c_sc = bo.content_similar(feat, case["liked"], 2 * k, valid_mask=valid)
        if case["user_id"] is not None:
            r = user_row[case["user_id"]]
            f_ids, f_sc = bo.cf_recommend(gf["user_factors"][r], gf["item_factors"], 2 * k,
                                          rated_mask=gf["ratings"][r] > 0, valid_mask=valid)
        if len(c_ids) and len(f_ids):
            ids, sc = bo.hybrid_combine(c_ids, c_sc, f_ids, f_sc, case["weights"][0], case["weights"][1], k)
        elif len(c_ids):
            ids, sc = c_ids[:k], c_sc[:k]
        else:
            ids, sc = f_ids[:k], f_sc[:k]
        ref_ids = np.array(case["ids"])
        ref_sc = np.array([np.nan if s is None else s for s in case["scores"]])
        _assert_ranked_equal(ref_ids, ref_sc, np.asarray(ids), np.asarray(sc), rtol=2e-6, what=str(case["kwargs"]))


def test_svd_training_oracle_reproduces_reference_factors(golden_dir):
    """truncated_svd_factors vs the factors the reference's train_svd_model produced (make_golden.py, cf_golden.npz)
    and vs scikit-learn's TruncatedSVD run here."""
    g = np.load(os.path.join(golden_dir, "cf_golden.npz"), allow_pickle=False)
    ratings = g["ratings"].astype(np.float64)
    uf, itf, s = bo.truncated_svd_factors(ratings, g["user_factors"].shape[1])
    np.testing.assert_allclose(uf, g["user_factors"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(itf, g["item_factors"], rtol=0, atol=1e-12)
    # more users than items (no transpose inside the randomized SVD): factors trained by the reference itself
    gt = np.load(os.path.join(golden_dir, "svd_tall_golden.npz"), allow_pickle=False)
    uft, itft, _ = bo.truncated_svd_factors(gt["ratings"].astype(np.float64), gt["user_factors"].shape[1])
    np.testing.assert_allclose(uft, gt["user_factors"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(itft, gt["item_factors"], rtol=0, atol=1e-12)
    from sklearn.decomposition import TruncatedSVD
    rng = np.random.default_rng(3)
    tall = (rng.random((300, 40)) < 0.2) * rng.integers(1, 6, (300, 40))       # more users than items: no transpose
    svd = TruncatedSVD(n_components=12, random_state=42)
    ref_u = svd.fit_transform(tall.astype(np.float64))
    uf2, itf2, s2 = bo.truncated_svd_factors(tall, 12)
    np.testing.assert_allclose(uf2, ref_u, rtol=0, atol=1e-10)
    np.testing.assert_allclose(itf2, svd.components_.T, rtol=0, atol=1e-12)
    np.testing.assert_allclose(s2, svd.singular_values_, rtol=1e-12)
