"""Parity of the CUDA path (through the C ABI) against the oracle and the golden
vectors produced by the unmodified reference.  Run on a B200: ``pytest -m gpu``.

Bars: exact-fp32 path -> mask and ids bit-identical on
tie-free inputs, scores within 1e-5 relative; bf16 path -> recall@k >= 0.999 with
re-scoring.
"""
import json
import os

import numpy as np
import pytest

import brick_oracle as bo
import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5          # BASELINE.json north_star: scores agree within 1e-5 relative


@pytest.fixture(scope="module")
def bb():
    import brickbrain_rec_engine_b200 as m
    return m


def _gap_free(scores_sorted_desc, tol):
    """positions whose neighbours are further than tol away (ids must match there)."""
    s = np.asarray(scores_sorted_desc, dtype=np.float64)
    if s.size <= 1:
        return np.ones(s.size, dtype=bool)
    g = np.abs(np.diff(s))
    return np.concatenate([[True], g > tol]) & np.concatenate([g > tol, [True]])


def assert_topk_equal(ref_ids, ref_scores, ids, scores, rtol=RTOL, atol=2e-6, what=""):
    ref_ids, ids = np.asarray(ref_ids), np.asarray(ids)
    assert ref_ids.shape == ids.shape, what
    for r in range(ref_ids.shape[0]):
        n = int((ref_ids[r] >= 0).sum())
        assert int((ids[r] >= 0).sum()) == n, f"{what} row {r}: result count {int((ids[r] >= 0).sum())} != {n}"
        assert (ids[r][n:] == -1).all()
        rs = np.asarray(ref_scores[r][:n], dtype=np.float64)
        np.testing.assert_allclose(np.asarray(scores[r][:n], dtype=np.float64), rs, rtol=rtol, atol=atol,
                                   err_msg=f"{what} row {r}")
        ok = _gap_free(rs, 4e-6 + 4 * rtol * float(np.abs(rs).max(initial=0)))
        assert np.array_equal(ids[r][:n][ok], ref_ids[r][:n][ok]), f"{what} row {r} ids differ"


def _catalog(n, d, f, seed, kind="aniso"):
    themes = synth.themes_table()
    emb = synth.embeddings(n, d, seed, kind)
    itf = synth.cf_factors(n, f, seed + 1) if f else None
    parts, year, theme = synth.metadata(n, seed + 2, themes)
    return themes, emb, itf, parts, year, theme


# ---------------------------------------------------------------------------------------------
# dense scores straight out of each scan pipeline
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,b", [(1000, 70), (257, 3), (25216, 40)])
def test_dense_scores_fp32(bb, n, b):
    _, emb, itf, *_ = _catalog(n, 384, 64, 1)
    q = synth.embeddings(b, 384, 9, "aniso") * 2.5
    u = synth.cf_factors(b, 64, 10)
    eng = bb.ScoringEngine(n, 384, 64)
    eng.load_catalog(emb=emb, item_factors=itf)
    got = eng.debug_scores(q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, precision="fp32")
    want = bo.fused_hybrid_scores(q, u, bo.l2_normalize_rows(emb), itf, np.full(b, 0.4, np.float32),
                                  np.full(b, 0.6, np.float32))
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=2e-6)
    eng.close()


@pytest.mark.parametrize("n,b", [(1000, 70), (257, 3), (25216, 300)])
def test_dense_scores_bf16_tcgen05(bb, n, b):
    import torch
    _, emb, itf, *_ = _catalog(n, 384, 64, 1)
    q = synth.embeddings(b, 384, 9, "aniso") * 2.5
    u = synth.cf_factors(b, 64, 10)
    eng = bb.ScoringEngine(n, 384, 64)
    eng.load_catalog(emb=emb, item_factors=itf)
    got = eng.debug_scores(q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, precision="bf16")
    rnd = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).to(torch.float32).numpy()
    qq = np.concatenate([np.float32(0.4) * bo.l2_normalize_rows(q), np.float32(0.6) * u], axis=1)
    ee = np.concatenate([bo.l2_normalize_rows(emb), itf], axis=1)
    want = rnd(qq) @ rnd(ee).T            # bf16 operands, fp32 accumulate
    # the host restatement normalizes with a different summation order, so a few operand elements land on the
    # neighbouring bf16 value (1 ulp = 2^-8 relative): bound those, and require everything else to be tight
    err = np.abs(got - want)
    assert err.max() < 3e-4
    assert (err > 2e-5).mean() < 1e-3
    eng.close()


# ---------------------------------------------------------------------------------------------
# fused hybrid top-k with constraints: exact path vs oracle
# ---------------------------------------------------------------------------------------------
def _scenario_descs(bb, eng, themes, rated_lists, b):
    """Per-query descriptors cycling through the reference's constraint scenarios."""
    filt = bb.HardConstraintFilter(eng, synth.set_nums(eng.n_items), themes, rated=rated_lists)
    descs, specs = [], []
    scen = synth.SCENARIO_KWARGS
    for j in range(b):
        name, kw = scen[j % len(scen)]
        kw = {k: v for k, v in kw.items() if k not in ("exclude_owned", "exclude_wishlisted", "user_id")}
        cons = filt.create_constraint_set(**kw)
        descs.append(filt.descriptor(cons, rated_user=j % len(rated_lists)))
        pairs = [(c.constraint_type.value, c.value) for c in cons]
        specs.append(bo.compile_constraints(pairs, themes, now_year=__import__("datetime").datetime.now().year))
    return np.stack(descs), specs


@pytest.mark.parametrize("precision", ["fp32", "bf16_rescore"])
def test_fused_hybrid_topk_with_constraints(bb, precision):
    n, d, f, b, k = 25216, 384, 64, 72, 50
    themes, emb, itf, parts, year, theme = _catalog(n, d, f, 21)
    q_items = np.arange(0, n, n // b)[:b].astype(np.int32)
    u = synth.cf_factors(b, f, 30)
    rated = {j: np.sort(np.random.default_rng(100 + j).choice(n, 25, replace=False)).astype(np.int32) for j in range(6)}
    eng = bb.ScoringEngine(n, d, f)
    eng.load_catalog(emb=emb, item_factors=itf, num_parts=parts, year=year, theme_id=theme)
    descs, specs = _scenario_descs(bb, eng, themes, rated, b)
    scores, ids = eng.search(k, q_item=q_items, q_cf=u, w_content=0.4, w_cf=0.6, constraints=descs, self_id=q_items,
                             precision=precision)
    e_unit = bo.l2_normalize_rows(emb)
    ref = bo.fused_hybrid_scores(e_unit[q_items], u, e_unit, itf, np.full(b, 0.4, np.float32), np.full(b, 0.6, np.float32))
    valid = np.stack([bo.constraint_mask(s, parts, year, theme) for s in specs])
    rid, rs = bo.fused_hybrid_topk(ref, k, valid=valid, self_ids=q_items, exclusions=[rated[j % 6] for j in range(b)])
    assert_topk_equal(rid, rs, ids, scores, what=precision)
    assert eng.info()["launches"] > 0
    # same batch through the deduplicated form (distinct constraint sets + per-query index -> class bitmaps)
    distinct, index = eng.dedupe_constraints(descs)
    assert distinct.shape[0] < b
    s2, i2 = eng.search(k, q_item=q_items, q_cf=u, w_content=0.4, w_cf=0.6, constraints=distinct, constraint_index=index,
                        self_id=q_items, precision=precision)
    assert np.array_equal(i2, ids) and np.array_equal(s2, scores)
    # device-resident per-query descriptors are not de-duplicated: this exercises the per-row prefilter of the epilogue
    import torch
    dev_desc = torch.from_numpy(descs.view(np.uint8).reshape(b, -1).copy()).cuda()
    s3, i3 = eng.search(k, q_item=torch.from_numpy(q_items).cuda(), q_cf=torch.from_numpy(u).cuda(), w_content=0.4, w_cf=0.6,
                        constraints=dev_desc, self_id=torch.from_numpy(q_items).cuda(), precision=precision)
    assert np.array_equal(i3.cpu().numpy(), ids) and np.array_equal(s3.cpu().numpy(), scores)
    eng.close()


def test_bf16_recall(bb):
    """bf16 tcgen05 path: recall@k vs the exact ranking, isotropic and anisotropic catalogs."""
    n, d, b, k = 50000, 384, 256, 100
    for kind in ("iso", "aniso"):
        emb = synth.embeddings(n, d, 5, kind)
        q = synth.embeddings(b, d, 6, kind)
        eng = bb.ScoringEngine(n, d, 0)
        eng.load_catalog(emb=emb)
        ref = bo.fused_hybrid_scores(q, None, emb, None, np.ones(b, np.float32), np.zeros(b, np.float32))
        rid, _ = bo.fused_hybrid_topk(ref, k)
        for prec, bar in (("bf16_rescore", 0.999), ("bf16", 0.90)):
            _, ids = eng.search(k, q_emb=q, precision=prec)
            rec = np.mean([len(set(ids[r]) & set(rid[r])) / k for r in range(b)])
            assert rec >= bar, f"{kind} {prec}: recall@{k} = {rec:.4f}"
        eng.close()


# ---------------------------------------------------------------------------------------------
# golden vectors from the unmodified reference
# ---------------------------------------------------------------------------------------------
def _golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _golden_hybrid(bb, golden_dir, precision="fp32"):
    gc = _golden(golden_dir, "content_fp32_golden.npz")
    gf = _golden(golden_dir, "cf_golden.npz")
    gh = _golden(golden_dir, "hybrid_golden.npz")
    feat = gc["q8"].astype(np.float32) / gc["scale"]
    n = feat.shape[0]
    themes = {int(k): v for k, v in json.loads(str(gh["themes_json"])).items()}
    owned = {int(k): np.array(v, dtype=np.int32) for k, v in json.loads(str(gh["owned_json"])).items()}
    wish = {int(k): np.array(v, dtype=np.int32) for k, v in json.loads(str(gh["wish_json"])).items()}
    cat = bb.CatalogTable(synth.set_nums(n), [f"Set {i}" for i in range(n)], gh["year"], gh["num_parts"],
                          gh["theme_id"], [themes.get(int(t), "Unknown") for t in gh["theme_id"]])
    ratings = gf["ratings"]
    user_ids = [int(u) for u in gf["user_ids"]]
    rated = {u: np.nonzero(ratings[i] > 0)[0].astype(np.int32) for i, u in enumerate(user_ids)}
    hyb = bb.HybridRecommender.from_arrays(cat, feat, themes, user_ids=user_ids, user_factors=gf["user_factors"],
                                           item_factors=gf["item_factors"], rated=rated, owned=owned, wishlisted=wish,
                                           precision=precision)
    return hyb, cat, gc, gf, gh


def _recs_arrays(recs, cat, k):
    ids = np.full((1, k), -1, dtype=np.int64)
    sc = np.full((1, k), np.nan)
    for j, r in enumerate(recs):
        ids[0, j] = cat.index_of[r.set_num]
        sc[0, j] = r.score
    return ids, sc


def test_golden_content_and_cf(bb, golden_dir):
    hyb, cat, gc, gf, _ = _golden_hybrid(bb, golden_dir)
    k = int(gc["k"])
    sn = cat.set_nums
    valid_filter = [sn[i] for i in gc["valid_ids"]]
    for j, t in enumerate(gc["targets"]):
        ids, sc = _recs_arrays(hyb.content_recommender.get_similar_sets(sn[int(t)], k), cat, k)
        assert_topk_equal(gc["ids_nofilter"][j][None], gc["scores_nofilter"][j][None], ids, sc, what=f"content {t}")
        ids, sc = _recs_arrays(hyb.content_recommender.get_similar_sets(sn[int(t)], k, valid_set_filter=valid_filter), cat, k)
        assert_topk_equal(gc["ids_filter"][j][None], gc["scores_filter"][j][None], ids, sc, what=f"content {t} filtered")
    assert hyb.content_recommender.get_similar_sets("no-such-set", k) == []
    k = int(gf["k"])
    for j, u in enumerate(gf["users"]):
        ids, sc = _recs_arrays(hyb.collaborative_recommender.get_recommendations(int(u), k), cat, k)
        assert_topk_equal(gf["ids_nofilter"][j][None], gf["scores_nofilter"][j][None], ids, sc, what=f"cf {u}")
        ids, sc = _recs_arrays(hyb.collaborative_recommender.get_recommendations(int(u), k, valid_set_filter=valid_filter), cat, k)
        assert_topk_equal(gf["ids_filter"][j][None], gf["scores_filter"][j][None], ids, sc, what=f"cf {u} filtered")
    hyb.engine.close()


def test_golden_hybrid_end_to_end(bb, golden_dir):
    hyb, cat, gc, gf, gh = _golden_hybrid(bb, golden_dir)
    hyb.constraint_filter.compiler.now_year = int(gh["now_year"])
    k = int(gh["k"])
    for case in json.loads(str(gh["cases_json"])):
        hyb.content_weight, hyb.collaborative_weight = case["weights"]
        cons = hyb.constraint_filter.create_constraint_set(**case["kwargs"]) if case["kwargs"] else None
        liked = cat.set_nums[case["liked"]] if case["liked"] is not None else None
        recs, cres = hyb.get_recommendations(user_id=case["user_id"], liked_set=liked, top_k=k, constraints=cons)
        if case["n_valid"] is not None:
            assert len(cres.valid_set_nums) == case["n_valid"]
        ids, sc = _recs_arrays(recs, cat, k)
        ref_sc = np.array([[np.nan if s is None else s for s in case["scores"]]])
        assert_topk_equal(np.array([case["ids"]]), ref_sc, ids, sc, what=str(case["kwargs"]))
    comb = json.loads(str(gh["combine_json"]))
    s, i = hyb.engine.hybrid_combine([comb["content_ids"]], [comb["content_scores"]], [comb["cf_ids"]],
                                     [comb["cf_scores"]], [0.4, 0.6], k)
    assert list(i[0]) == comb["out_ids"]
    np.testing.assert_allclose(s[0], comb["out_scores"], rtol=1e-7)
    hyb.engine.close()


def test_golden_masks(bb, golden_dir):
    g = _golden(golden_dir, "mask_golden.npz")
    themes = {int(k): v for k, v in json.loads(str(g["themes_json"])).items()}
    owned = {int(k): np.array(v, dtype=np.int32) for k, v in json.loads(str(g["owned_json"])).items()}
    wish = {int(k): np.array(v, dtype=np.int32) for k, v in json.loads(str(g["wish_json"])).items()}
    n = g["num_parts"].shape[0]
    eng = bb.ScoringEngine(n, 8, 0)
    eng.load_catalog(emb=np.ones((n, 8), np.float32), num_parts=g["num_parts"], year=g["year"], theme_id=g["theme_id"])
    filt = bb.HardConstraintFilter(eng, synth.set_nums(n), themes, owned=owned, wishlisted=wish)
    filt.compiler.now_year = int(g["now_year"])
    specs = json.loads(str(g["scenario_specs"]))
    off = g["valid_off"]
    for j, name in enumerate(g["scenario_names"]):
        cons = [bb.HardConstraint(bb.ConstraintType(t), v) for t, v in specs[j]["pairs"]]
        res = filt.apply_constraints(cons)
        got = np.array([filt.index_of[s] for s in res.valid_set_nums], dtype=np.int64)
        assert np.array_equal(got, g["valid_flat"][off[j]:off[j + 1]]), f"scenario {name}"
        assert res.performance_stats["output_sets"] == len(got)
    eng.close()


def test_vector_store_faiss_fixture(bb, golden_dir):
    g = _golden(golden_dir, "faiss_fixture_golden.npz")
    emb = g["emb"]
    docs = [bb.Document(page_content=f"set {i}", metadata={"set_num": f"{i}-1", "name": f"Set {i}", "year": 2020,
                                                          "num_parts": 100 + i, "theme": "City"}) for i in range(10)]
    store = bb.BrickVectorStore(emb, docs)
    for t in range(10):
        hits = store.similarity_search_with_score_by_vector(emb[t], k=6)
        assert hits[0][0].metadata["set_num"] == f"{t}-1"
        ids = np.array([[int(d.metadata["set_num"].split("-")[0]) for d, _ in hits[1:]]])
        sc = np.array([[s for _, s in hits[1:]]])
        assert_topk_equal(g["ids"][t][None], g["scores"][t][None], ids, sc, what=f"fixture row {t}")
    l2 = bb.BrickVectorStore(emb, docs, score_mode="l2")
    d = l2.similarity_search_with_score_by_vector(emb[3], k=4)
    np.testing.assert_allclose([s for _, s in d][1:], 2.0 - 2.0 * g["scores"][3][:3], atol=3e-6)
    qvd = bb.make_query_vector_database(store)
    rows = qvd(emb[0], 5)
    assert len(rows) == 5 and set(rows[0]) >= {"set_num", "name", "year", "num_parts", "theme", "theme_id", "img_url", "relevance_score"}
    retr = store.as_retriever(search_kwargs={"k": 20})
    assert retr.k == 20


# ---------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16_rescore"])
def test_edge_cases(bb, precision):
    n, d = 300, 64
    emb = synth.embeddings(n, d, 3, "iso")
    emb[10] = emb[20] = emb[30]                      # exact duplicates -> ties broken by lower id
    emb[40] = 0.0                                    # zero row -> cosine 0 (sklearn normalize semantics)
    parts = np.full(n, 100, np.int32)
    parts[:290] = 0                                  # only 10 rows satisfy num_parts > 0
    eng = bb.ScoringEngine(n, d, 0)
    eng.load_catalog(emb=emb, num_parts=parts, year=np.full(n, 2020, np.int16), theme_id=np.full(n, -1, np.int16))
    # ties
    s, i = eng.search(5, q_emb=emb[30:31], precision=precision)
    assert list(i[0][:3]) == [10, 20, 30]
    # fewer valid rows than k -> -1 / -inf padding
    c = bb.no_constraint()
    c["parts_min"] = 1
    s, i = eng.search(16, q_emb=emb[:2], constraints=c, precision=precision)
    assert ((i >= 0).sum(axis=1) == 10).all() and (i[:, 10:] == -1).all() and np.isneginf(s[:, 10:]).all()
    assert set(i[0][:10]) == set(range(290, 300))
    # nothing valid at all
    c["flags"] = 1
    s, i = eng.search(4, q_emb=emb[:1], constraints=c, precision=precision)
    assert (i == -1).all()
    # zero query: every cosine is 0 -> lowest ids
    s, i = eng.search(4, q_emb=np.zeros((1, d), np.float32), precision=precision)
    assert list(i[0]) == [0, 1, 2, 3] and np.allclose(s, 0.0)
    # error behaviour
    with pytest.raises(bb.BrickscoreError):
        eng.search(500, q_emb=emb[:1])
    with pytest.raises(bb.BrickscoreError):
        eng.search(5, q_item=np.array([n + 5], np.int32))
    eng.close()


def test_merge_topk_and_sorted_output(bb):
    rng = np.random.default_rng(0)
    g, b, k = 4, 33, 20
    scores = rng.standard_normal((g, b, k)).astype(np.float32)
    scores = -np.sort(-scores, axis=2)
    ids = rng.permutation(g * b * k).astype(np.int32).reshape(g, b, k)
    ids[1, :, 15:] = -1
    eng = bb.ScoringEngine(0, 8, 0)
    s, i = eng.merge_topk(scores, ids, k)
    for r in range(b):
        cand = [(float(scores[l, r, j]), int(ids[l, r, j])) for l in range(g) for j in range(k) if ids[l, r, j] >= 0]
        cand.sort(key=lambda x: (-x[0], x[1]))
        assert [c[1] for c in cand[:k]] == list(i[r])
        np.testing.assert_array_equal(np.array([c[0] for c in cand[:k]], np.float32), s[r])
    eng.close()


def test_item_item_topk(bb):
    n_users, n_items, k = 3000, 400, 20
    indptr, indices, data = synth.zipf_ratings_csr(n_users, n_items, 12, seed=4)
    eng = bb.ScoringEngine(0, 8, 0)
    s, i = eng.item_item_topk(indptr, indices, data, n_items, k)
    rid, rs = bo.item_item_topk(indptr, indices, data, n_items, k)
    assert_topk_equal(rid, rs, i, s, rtol=2e-6, what="item-item")
    eng.close()


def test_large_catalog_properties(bb):
    """BASELINE-scale shapes checked through size-independent properties, with torch on the GPU as the checker:
    sortedness, uniqueness, every returned score equals the exact dot product, and the k-th score bounds
    everything that was not returned."""
    import torch
    n, d, f, b, k = 1_000_000, 384, 64, 512, 100
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(1)
    emb = torch.nn.functional.normalize(torch.randn(n, d, device=dev, generator=g) + 0.7 * torch.randn(1, d, device=dev, generator=g), dim=1)
    itf = torch.randn(n, f, device=dev, generator=g) / 8.0
    parts = torch.randint(1, 5000, (n,), device=dev, dtype=torch.int32, generator=g)
    q = torch.nn.functional.normalize(torch.randn(b, d, device=dev, generator=g), dim=1)
    u = torch.randn(b, f, device=dev, generator=g) / 8.0
    eng = bb.ScoringEngine(n, d, f, max_batch=b)
    eng.load_catalog(emb=emb, item_factors=itf, num_parts=parts, year=torch.full((n,), 2015, device=dev, dtype=torch.int16),
                     theme_id=torch.full((n,), 7, device=dev, dtype=torch.int16))
    c = bb.no_constraint()
    c["parts_min"], c["parts_max"] = 1, 2500
    scores, ids = eng.search(k, q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=c, precision="bf16_rescore")
    torch.cuda.synchronize()
    ref = 0.4 * (q @ emb.T) + 0.6 * (u @ itf.T)
    ref[:, parts > 2500] = -float("inf")
    top = torch.topk(ref, k, dim=1)
    ids64 = ids.long()
    assert (scores[:, :-1] >= scores[:, 1:]).all()
    assert all(len(set(r.tolist())) == k for r in ids64.cpu())
    assert (parts[ids64] <= 2500).all()
    torch.testing.assert_close(scores, torch.gather(ref, 1, ids64), rtol=2e-5, atol=2e-6)
    recall = np.mean([len(set(a.tolist()) & set(t.tolist())) / k for a, t in zip(ids64.cpu(), top.indices.cpu())])
    assert recall >= 0.999, recall
    eng.close()


# ---------------------------------------------------------------------------------------------
# more shapes, the reference's own feature shape, bulk / sharded drivers, re-entrancy
# ---------------------------------------------------------------------------------------------
def test_golden_content_f64_features(bb, golden_dir):
    """The reference's own content matrix is ~10 float64 columns (recommendation_system.py:179-191); the engine
    computes in fp32, so scores agree to 1e-5 relative and ids wherever the reference's gap exceeds that."""
    g = _golden(golden_dir, "content_f64_golden.npz")
    feat = g["feat"]
    n = feat.shape[0]
    themes = synth.themes_table()
    cat = bb.CatalogTable(synth.set_nums(n), [f"Set {i}" for i in range(n)], np.full(n, 2020), np.full(n, 100),
                          np.ones(n), ["x"] * n)
    hyb = bb.HybridRecommender.from_arrays(cat, feat, themes)
    k = int(g["k"])
    for j, t in enumerate(g["targets"]):
        ids, sc = _recs_arrays(hyb.content_recommender.get_similar_sets(cat.set_nums[int(t)], k), cat, k)
        assert_topk_equal(g["ids"][j][None], g["scores"][j][None], ids, sc, what=f"f64 features, target {t}")
    hyb.engine.close()


@pytest.mark.parametrize("n,d,f,b,k", [(5, 3, 0, 2, 4), (300, 10, 5, 257, 128), (70000, 64, 0, 1, 100), (1000, 0, 16, 33, 10)])
def test_odd_shapes_all_precisions(bb, n, d, f, b, k):
    rng = np.random.default_rng(n + d + f)
    emb = rng.standard_normal((n, d)).astype(np.float32) if d else None
    itf = rng.standard_normal((n, f)).astype(np.float32) if f else None
    q = rng.standard_normal((b, d)).astype(np.float32) if d else None
    u = rng.standard_normal((b, f)).astype(np.float32) if f else None
    eng = bb.ScoringEngine(n, d, f)
    eng.load_catalog(emb=emb, item_factors=itf)
    wc, wf = (0.4 if d else 0.0), (0.6 if f else 0.0)
    ref = bo.fused_hybrid_scores(q, u, bo.l2_normalize_rows(emb) if d else None, itf,
                                 np.full(b, wc, np.float32), np.full(b, wf, np.float32))
    kk = min(k, 128)
    rid, rs = bo.fused_hybrid_topk(ref, kk)
    s, i = eng.search(kk, q_emb=q, q_cf=u, w_content=wc, w_cf=wf, precision="fp32")
    assert_topk_equal(rid, rs, i, s, what=f"fp32 n={n} d={d} f={f}")
    # bf16 scan + exact re-score: the candidate set comes from bf16 scores (k' = min(128, 1.25k+8) over-fetch), so on
    # a dense ranking (k close to n) a few boundary items may differ; every returned score is exact fp32
    s, i = eng.search(kk, q_emb=q, q_cf=u, w_content=wc, w_cf=wf, precision="bf16_rescore")
    recall = np.mean([len(set(i[r][i[r] >= 0]) & set(rid[r][rid[r] >= 0])) / max(1, (rid[r] >= 0).sum()) for r in range(b)])
    assert recall >= (0.999 if kk <= 100 and n >= 1000 else 0.97), recall
    for r in range(b):
        m = i[r] >= 0
        np.testing.assert_allclose(s[r][m], ref[r][i[r][m]], rtol=RTOL, atol=2e-6)
        assert (np.diff(s[r][m]) <= 0).all()
    eng.close()


def test_bulk_and_sharded_single_rank(bb):
    import torch
    n, d, f, users, k = 20000, 64, 16, 700, 20
    emb = synth.embeddings(n, d, 1, "aniso")
    itf = synth.cf_factors(n, f, 2)
    uf = synth.cf_factors(users, f, 3)
    liked = np.random.default_rng(4).integers(0, n, users).astype(np.int32)
    eng = bb.ScoringEngine(n, d, f)
    eng.load_catalog(emb=emb, item_factors=itf)
    outs = bb.bulk_recommend(eng, uf, liked, k, batch=256, precision="bf16_rescore")
    ids = np.concatenate([np.asarray(i) for _, i in outs])
    e_unit = bo.l2_normalize_rows(emb)
    ref = bo.fused_hybrid_scores(e_unit[liked], uf, e_unit, itf, np.full(users, 0.4, np.float32), np.full(users, 0.6, np.float32))
    rid, _ = bo.fused_hybrid_topk(ref, k, self_ids=liked)
    assert (ids == rid).mean() > 0.999
    # device-resident inputs give device-resident outputs
    outs_d = bb.bulk_recommend(eng, torch.from_numpy(uf).cuda(), torch.from_numpy(liked).cuda(), k, batch=512)
    assert outs_d[0][1].is_cuda and np.array_equal(torch.cat([i for _, i in outs_d]).cpu().numpy(), ids)
    eng.close()


def test_reentrant_two_streams(bb):
    """bb_search is re-entrant on an immutable catalog: two host threads on two CUDA streams."""
    import threading
    import torch
    n, d, b, k = 30000, 128, 300, 32
    emb = synth.embeddings(n, d, 1, "aniso")
    eng = bb.ScoringEngine(n, d, 0)
    eng.load_catalog(emb=emb)
    qs = [synth.embeddings(b, d, 10 + t, "aniso") for t in range(2)]
    want = [eng.search(k, q_emb=q, precision="bf16_rescore") for q in qs]
    got = [None, None]

    def work(t):
        st = torch.cuda.Stream()
        for _ in range(5):
            got[t] = eng.search(k, q_emb=qs[t], precision="bf16_rescore", stream=st.cuda_stream)

    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    for t in range(2):
        assert np.array_equal(got[t][1], want[t][1]) and np.array_equal(got[t][0], want[t][0])
    eng.close()


def test_sharded_search_two_gpus():
    """Row-sharded catalog on 2 GPUs (NCCL all-gather + merge, thresholds shared through NVLink peer memory) returns
    exactly what one engine over the whole catalog returns (tools/shard_check.py).  Needs a 2-GPU box."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SC_N="200000", SC_B="512")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "shard_check.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert r.returncode == 0 and "SHARD_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


# ---------------------------------------------------------------------------------------------
# CF factor training: CSR x dense fp64 kernel + the randomized truncated SVD around it
# ---------------------------------------------------------------------------------------------
def test_csr_spmm_matches_scipy(bb):
    import scipy.sparse as sp
    import torch
    rng = np.random.default_rng(11)
    eng = bb.ScoringEngine(64, 64, 0)
    for (r, c, dens, l) in ((500, 300, 0.05, 60), (64, 1536, 0.02, 17), (3000, 40, 0.9, 128), (5, 7, 0.0, 33)):
        a = sp.random(r, c, density=dens, format="csr", random_state=7, dtype=np.float32,
                      data_rvs=lambda n: rng.integers(1, 6, n).astype(np.float32))
        if r == 3000:                                  # one very long row (> the long-row kernel's threshold)
            a = sp.vstack([a, sp.csr_matrix(np.tile(rng.integers(1, 6, c).astype(np.float32), (1, 1)))]).tocsr()
            a = sp.hstack([a] * 70).tocsr()            # 2800 columns, dense rows of ~2500 entries
        a.sort_indices()
        x = rng.standard_normal((a.shape[1], l))
        y = eng.csr_spmm(torch.from_numpy(a.indptr.astype(np.int64)).cuda(), torch.from_numpy(a.indices.astype(np.int32)).cuda(),
                         torch.from_numpy(a.data.astype(np.float32)).cuda(), a.shape[1], torch.from_numpy(x).cuda())
        ref = a.astype(np.float64) @ x
        np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-12, atol=1e-12)
    with pytest.raises(Exception):
        eng.csr_spmm(torch.zeros(3, dtype=torch.int64).cuda(), torch.zeros(1, dtype=torch.int32).cuda(),
                     torch.zeros(1, dtype=torch.float32).cuda(), 4, torch.zeros((4, 200), dtype=torch.float64).cuda())   # L > 128
    eng.close()


def test_svd_training_matches_reference_factors(bb, golden_dir):
    """train_svd_model on the GPU vs the factors the unmodified reference trained (cf_golden.npz, TruncatedSVD
    random_state=42) and vs the oracle on a tall matrix; fp64 arithmetic, bar 1e-8 absolute on factors of O(10)."""
    from brickbrain_rec_engine_b200 import cf_training
    g = _golden(golden_dir, "cf_golden.npz")
    ratings = g["ratings"].astype(np.float32)
    eng = bb.ScoringEngine(64, 64, 0)
    uf, itf, s = cf_training.truncated_svd(eng, ratings, g["user_factors"].shape[1])
    np.testing.assert_allclose(uf, g["user_factors"], rtol=0, atol=1e-8)
    np.testing.assert_allclose(itf, g["item_factors"], rtol=0, atol=1e-9)
    gt = _golden(golden_dir, "svd_tall_golden.npz")          # 300 users x 96 items: no transpose inside the SVD
    uft, itft, _ = cf_training.truncated_svd(eng, gt["ratings"].astype(np.float32), gt["user_factors"].shape[1])
    np.testing.assert_allclose(uft @ itft.T, gt["user_factors"] @ gt["item_factors"].T, rtol=0, atol=1e-8)
    np.testing.assert_allclose(itft, gt["item_factors"], rtol=0, atol=1e-7)
    rng = np.random.default_rng(5)
    tall = ((rng.random((4000, 600)) < 0.03) * rng.integers(1, 6, (4000, 600))).astype(np.float32)
    ouf, oitf, os_ = bo.truncated_svd_factors(tall, 50)
    uf2, itf2, s2 = cf_training.truncated_svd(eng, tall, 50)
    np.testing.assert_allclose(s2, os_, rtol=1e-10)
    np.testing.assert_allclose(uf2 @ itf2.T, ouf @ oitf.T, rtol=0, atol=1e-8)       # what scoring consumes
    np.testing.assert_allclose(itf2, oitf, rtol=0, atol=1e-7)
    # recommender seam: same guards as recommendation_system.py:381-396
    cf = bb.CollaborativeFilteringRecommender(engine=eng)
    cf.user_item_matrix = ratings
    cf.train_svd_model(50)
    assert cf.svd_model is not None
    np.testing.assert_allclose(cf.item_factors, g["item_factors"], rtol=0, atol=1e-9)
    cf.user_item_matrix = ratings[:2]
    cf.train_svd_model(50)
    assert cf.svd_model is None
    eng.close()


def test_small_batch_exact_path_matches_tile_scan(bb):
    """B <= 16 exact-fp32 searches take the HBM-streaming path (bb_scan_small.cu); it must return what the 128-row
    FFMA tile scan returns (BB_NO_SMALL=1) and what the oracle returns: filters, self exclusion, constraints,
    exclusion lists, ragged catalog size, k up to 128."""
    n, d, f = 70001, 384, 64
    themes, emb, itf, parts, year, theme = _catalog(n, d, f, 77)
    eng = bb.ScoringEngine(n, d, f)
    eng.load_catalog(emb=emb, item_factors=itf, num_parts=parts, year=year, theme_id=theme)
    rated = {j: np.sort(np.random.default_rng(200 + j).choice(n, 40, replace=False)).astype(np.int32) for j in range(6)}
    e_unit = bo.l2_normalize_rows(emb)
    for b, k in ((1, 10), (3, 50), (16, 128)):
        q_items = (np.arange(b, dtype=np.int32) * 997 + 5) % n
        u = synth.cf_factors(b, f, 31 + b)
        descs, specs = _scenario_descs(bb, eng, themes, rated, b)
        kw = dict(q_item=q_items, q_cf=u, w_content=0.4, w_cf=0.6, constraints=descs, self_id=q_items, precision="fp32")
        s1, i1 = eng.search(k, **kw)
        assert eng.info()["last_units"] == b * min(128, max(1, -(-n // 1024)))         # the small path ran
        os.environ["BB_NO_SMALL"] = "1"
        try:
            s2, i2 = eng.search(k, **kw)
        finally:
            del os.environ["BB_NO_SMALL"]
        ref = bo.fused_hybrid_scores(e_unit[q_items], u, e_unit, itf, np.full(b, 0.4, np.float32), np.full(b, 0.6, np.float32))
        valid = np.stack([bo.constraint_mask(s, parts, year, theme) for s in specs])
        rid, rs = bo.fused_hybrid_topk(ref, k, valid=valid, self_ids=q_items, exclusions=[rated[j % 6] for j in range(b)])
        assert_topk_equal(rid, rs, i1, s1, what=f"small path b={b}")
        assert_topk_equal(rid, rs, i2, s2, what=f"tile path b={b}")
    eng.close()
