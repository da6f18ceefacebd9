"""GPU parity of the round-2 paths: per-call filters and per-query exclusion lists, the pilot pass (threshold
seeding), the collective search of a row-sharded catalog (union thresholds + owner inboxes) with several shards on
ONE GPU, concurrent requests with different filters.  Run on a B200: ``pytest -m gpu``."""
import os
import threading

import numpy as np
import pytest

import brick_oracle as bo
import synth
from test_gpu_engine import assert_topk_equal, _catalog

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bb():
    import brickbrain_rec_engine_b200 as m
    return m


def _mix_constraints(bb, n_sets=4):
    L = bb._lib
    cons = np.zeros(n_sets, dtype=L.CONSTRAINT_DTYPE)
    cons[:] = L.no_constraint()
    cons[1]["parts_min"], cons[1]["parts_max"] = 200, 900
    cons[2]["year_min"] = 2012
    cons[3]["parts_max"], cons[3]["year_max"] = 150, 2005
    return cons


def _mix_valid(cons_idx, parts, year):
    v = [np.ones(parts.shape[0], bool), (parts >= 200) & (parts <= 900), year >= 2012, (parts <= 150) & (year <= 2005)]
    return np.stack([v[c] for c in cons_idx])


def _same_exact_ranking(ref_i, ref_s, got_i, got_s, what):
    """Two exact-fp32 paths (tensor pipe + re-score vs FFMA tile scan) sum in different orders.  Position by position the
    scores must agree to fp32 rounding -- a missed or wrong item would show up as a lower score -- and ids may differ only
    where that happens between items whose scores tie within the rounding (a handful per thousand)."""
    ref_i, got_i = np.asarray(ref_i), np.asarray(got_i)
    assert ref_i.shape == got_i.shape, what
    np.testing.assert_allclose(np.asarray(got_s, np.float64), np.asarray(ref_s, np.float64), rtol=1e-5, atol=2e-6, err_msg=what)
    assert ((ref_i >= 0) == (got_i >= 0)).all(), what
    assert (ref_i != got_i).mean() < 0.01, (what, float((ref_i != got_i).mean()))


@pytest.mark.parametrize("precision", ["fp32", "bf16_rescore"])
def test_per_call_allow_bitmaps_and_query_exclusions(bb, precision):
    """The request's valid_set_filter travels with the call (bb_query_batch.allow_bitmaps) and the users' rated
    lists are per-query attributes (excl_lists), so a batch keeps few distinct constraint sets (class bitmaps)."""
    n, d, f, b, k = 30000, 384, 64, 96, 40
    themes, emb, itf, parts, year, theme = _catalog(n, d, f, 41)
    rng = np.random.default_rng(3)
    q = synth.embeddings(b, d, 42, "aniso") * 1.7
    u = synth.cf_factors(b, f, 43)
    rated = [np.sort(rng.choice(n, 60, replace=False)).astype(np.int32) for _ in range(b)]
    allow_a = rng.random(n) < 0.4
    allow_b = rng.random(n) < 0.05
    eng = bb.ScoringEngine(n, d, f)
    eng.load_catalog(emb=emb, item_factors=itf, num_parts=parts, year=year, theme_id=theme)
    eng.set_exclusion_lists(rated)
    cons = _mix_constraints(bb, 4)
    cons[1]["allow_bitmap"] = 0
    cons[3]["allow_bitmap"] = 1
    cidx = (np.arange(b) % 4).astype(np.int32)
    allow = np.stack([eng.pack_bitmap(allow_a), eng.pack_bitmap(allow_b)])
    excl = np.arange(b, dtype=np.int32).reshape(b, 1)
    s, i = eng.search(k, q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, constraint_index=cidx,
                      allow_bitmaps=allow, excl_lists=excl, precision=precision)
    valid = _mix_valid(cidx, parts, year)
    valid[cidx == 1] &= allow_a
    valid[cidx == 3] &= allow_b
    ref = bo.fused_hybrid_scores(q, u, bo.l2_normalize_rows(emb), itf, np.full(b, 0.4, np.float32), np.full(b, 0.6, np.float32))
    rid, rs = bo.fused_hybrid_topk(ref, k, valid=valid, exclusions=rated)
    assert_topk_equal(rid, rs, i, s, what=precision)
    # a device-resident bitmap (as bb_apply_constraints writes it) is used in place
    import torch
    allow_dev = torch.from_numpy(allow.view(np.int32)).cuda()
    s2, i2 = eng.search(k, q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, constraint_index=cidx,
                        allow_bitmaps=allow_dev, excl_lists=excl, precision=precision)
    assert np.array_equal(i2.cpu().numpy(), i) and np.array_equal(s2.cpu().numpy(), s)
    # per-row descriptor path (no constraint index -> one descriptor per query) with the same per-call tables
    per_row = cons[cidx]
    s3, i3 = eng.search(k, q_emb=torch.from_numpy(q).cuda(), q_cf=torch.from_numpy(u).cuda(), w_content=0.4, w_cf=0.6,
                        constraints=torch.from_numpy(per_row.view(np.uint8).reshape(b, -1).copy()).cuda(),
                        allow_bitmaps=allow_dev, excl_lists=torch.from_numpy(excl).cuda(), precision=precision)
    assert np.array_equal(i3.cpu().numpy(), i) and np.array_equal(s3.cpu().numpy(), s)
    eng.close()


def test_pilot_seeding_changes_nothing_but_the_work(bb):
    """The pilot pass seeds every row's threshold with a valid lower bound: results are bit-identical with and
    without it, and the scan appends far fewer candidates."""
    n, d, f, b, k = 120000, 384, 64, 512, 50
    themes, emb, itf, parts, year, theme = _catalog(n, d, f, 51)
    rng = np.random.default_rng(9)
    q = synth.embeddings(b, d, 52, "aniso")
    u = synth.cf_factors(b, f, 53)
    self_id = rng.integers(0, n, b).astype(np.int32)
    rated = [np.sort(rng.choice(20000, 40, replace=False)).astype(np.int32) for _ in range(b)]   # inside the pilot range too
    eng = bb.ScoringEngine(n, d, f)
    eng.load_catalog(emb=emb, item_factors=itf, num_parts=parts, year=year, theme_id=theme)
    eng.set_exclusion_lists(rated)
    cons = _mix_constraints(bb, 4)
    cidx = (np.arange(b) % 4).astype(np.int32)
    kw = dict(q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, constraint_index=cidx, self_id=self_id,
              excl_lists=np.arange(b, dtype=np.int32))
    out, appends = {}, {}
    for mode in ("pilot", "cold"):
        if mode == "cold":
            os.environ["BB_NO_PILOT"] = "1"
        try:
            for prec in ("bf16", "bf16_rescore"):
                out[mode, prec] = eng.search(k, precision=prec, **kw)
                appends[mode, prec] = eng.info()["last_appends"]
        finally:
            os.environ.pop("BB_NO_PILOT", None)
    for prec in ("bf16", "bf16_rescore"):
        assert np.array_equal(out["pilot", prec][1], out["cold", prec][1]), prec
        assert np.array_equal(out["pilot", prec][0], out["cold", prec][0]), prec
        assert appends["pilot", prec] < 0.6 * appends["cold", prec], (prec, appends)
    # and against the oracle
    valid = _mix_valid(cidx, parts, year)
    ref = bo.fused_hybrid_scores(q, u, bo.l2_normalize_rows(emb), itf, np.full(b, 0.4, np.float32), np.full(b, 0.6, np.float32))
    rid, rs = bo.fused_hybrid_topk(ref, k, valid=valid, self_ids=self_id, exclusions=rated)
    s, i = out["pilot", "bf16_rescore"]
    rec = np.mean([len(set(i[r]) & set(rid[r])) / k for r in range(b)])
    assert rec >= 0.999, rec
    eng.close()


@pytest.mark.parametrize("world", [2, 4])
def test_collective_search_shards_on_one_gpu(bb, world):
    """``world`` shards of one catalog as ``world`` handles on ONE device, linked with bb_comm_link_local: the
    collective search (union thresholds read from the other shards' blocks, candidates delivered to the owner's
    inbox, owner merge) must return, slice by slice, exactly what one engine over the whole catalog returns."""
    n, d, f, b, k = 200000, 384, 64, 700, 100          # b not a multiple of world * anything: ragged last slice
    themes, emb, itf, parts, year, theme = _catalog(n, d, f, 61)
    rng = np.random.default_rng(8)
    q = synth.embeddings(b, d, 62, "aniso") * 2.5
    u = synth.cf_factors(b, f, 63)
    self_id = rng.integers(0, n, b).astype(np.int32)
    rated = [np.sort(rng.choice(n, 30, replace=False)).astype(np.int32) for _ in range(b)]
    cons = _mix_constraints(bb, 4)
    cidx = (np.arange(b) % 4).astype(np.int32)
    kw = dict(q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, constraint_index=cidx, self_id=self_id,
              excl_lists=np.arange(b, dtype=np.int32))
    full = bb.ScoringEngine(n, d, f, max_batch=b)
    full.load_catalog(emb=emb, item_factors=itf, num_parts=parts, year=year, theme_id=theme)
    full.set_exclusion_lists(rated)
    shards = []
    for r in range(world):
        lo, hi = bb.shard_bounds(n, world, r)
        e = bb.ScoringEngine(hi - lo, d, f, id_offset=lo, max_batch=b)
        e.load_catalog(emb=emb[lo:hi], item_factors=itf[lo:hi], num_parts=parts[lo:hi], year=year[lo:hi], theme_id=theme[lo:hi])
        e.set_exclusion_lists(rated)
        shards.append(e)
    for r, e in enumerate(shards):
        e.comm_link_local(r, shards)
    from brickbrain_rec_engine_b200.sharded import slice_bounds
    for rep in range(3):                      # the inbox halves alternate, the arrival counters accumulate
        for prec in ("fp32", "bf16_rescore"):
            fs, fi = full.search(k, precision=prec, **kw)
            mine = [e.sharded_scatter(k, precision=prec, **kw) for e in shards]
            appended = [e.info()["last_appends"] for e in shards]
            got = [e.sharded_collect(k, m) for e, m in zip(shards, mine)]
            for r in range(world):
                lo, hi = slice_bounds(b, world, r)
                assert mine[r] == hi - lo
                s, i = got[r]
                if prec == "fp32":
                    # ids bit-identical; scores to fp32 rounding (the single engine's exact search runs on the tensor
                    # pipe and re-scores with another summation order than the shards' FFMA scan)
                    _same_exact_ranking(fi[lo:hi], fs[lo:hi], i, s, f"fp32 shard {r}")
                else:
                    rec = np.mean([len(set(i[x]) & set(fi[lo + x])) / k for x in range(hi - lo)])
                    assert rec >= 0.999, (prec, r, rec)
                    m = i == fi[lo:hi]
                    np.testing.assert_allclose(s[m], fs[lo:hi][m], rtol=1e-5, atol=2e-6)
            if prec == "bf16_rescore" and world == 2 and rep == 0:
                union_appends = appended
    # Shards that scan later read the earlier shards' claims (here: shard 1 starts after shard 0 has finished).  On a
    # catalog that is long enough for the poller warp to matter, the union bound must cut the later shard's appends.
    if world == 2:
        n2 = 1_200_000
        themes2, emb2, itf2, parts2, year2, theme2 = _catalog(n2, d, f, 65)
        big = []
        for r in range(2):
            lo, hi = bb.shard_bounds(n2, 2, r)
            e = bb.ScoringEngine(hi - lo, d, f, id_offset=lo, max_batch=b)
            e.load_catalog(emb=emb2[lo:hi], item_factors=itf2[lo:hi], num_parts=parts2[lo:hi], year=year2[lo:hi], theme_id=theme2[lo:hi])
            big.append(e)
        for r, e in enumerate(big):
            e.comm_link_local(r, big)
        kw2 = dict(q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, constraint_index=cidx)
        mine = [e.sharded_scatter(k, precision="bf16_rescore", **kw2) for e in big]
        appended = [e.info()["last_appends"] for e in big]
        for e, m in zip(big, mine):
            e.sharded_collect(k, m)
        assert appended[1] < 0.85 * appended[0], appended
        for e in big:
            e.close()
    # q_item resolved by the owner shard is not available in-process (ShardedSearch does it); an id outside the
    # shard must still be reported on the device-output path, at the next call
    import torch
    bad = shards[1].search(5, q_item=torch.tensor([0], dtype=torch.int32).cuda(), precision="fp32")
    torch.cuda.synchronize()
    with pytest.raises(bb.BrickscoreError):
        shards[1].search(5, q_emb=q[:1], precision="fp32")
    del bad
    for e in shards + [full]:
        e.close()


def test_pipelined_collective_searches_on_one_gpu(bb):
    """bb_sharded_scatter_after: the owner-side merge of search s is launched inside search s+1 (after its prologue, before
    its scan).  Four different batches back to back on three in-process shards: every search's slices must equal the
    single-engine result of ITS batch (a stale or early merge would return another batch's rows), the inbox halves
    alternate, the last search is finished by a plain collect."""
    import torch
    from brickbrain_rec_engine_b200.sharded import slice_bounds
    world, n, d, f, b, k = 3, 150_000, 384, 64, 520, 60
    themes, emb, itf, parts, year, theme = _catalog(n, d, f, 71)
    cons = _mix_constraints(bb, 4)
    cidx = (np.arange(b) % 4).astype(np.int32)
    full = bb.ScoringEngine(n, d, f, max_batch=b)
    full.load_catalog(emb=emb, item_factors=itf, num_parts=parts, year=year, theme_id=theme)
    shards = []
    for r in range(world):
        lo, hi = bb.shard_bounds(n, world, r)
        e = bb.ScoringEngine(hi - lo, d, f, id_offset=lo, max_batch=b)
        e.load_catalog(emb=emb[lo:hi], item_factors=itf[lo:hi], num_parts=parts[lo:hi], year=year[lo:hi], theme_id=theme[lo:hi])
        shards.append(e)
    for r, e in enumerate(shards):
        e.comm_link_local(r, shards)
    batches = []
    for step in range(4):
        q = torch.from_numpy(synth.embeddings(b, d, 80 + step, "aniso") * 2.0).cuda()
        u = torch.from_numpy(synth.cf_factors(b, f, 90 + step)).cuda()
        batches.append(dict(q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, constraint_index=cidx,
                            precision="bf16_rescore" if step % 2 else "fp32"))
    outs = []                                               # outs[step][r] = (scores, ids) CUDA tensors of shard r's slice
    for step, kw in enumerate(batches):
        cur = []
        for r, e in enumerate(shards):
            lo, hi = slice_bounds(b, world, r)
            prev = None if step == 0 else outs[step - 1][r]
            mine = e.sharded_scatter(k, collect_previous_into=prev, **kw)
            assert mine == hi - lo
            cur.append((torch.empty((hi - lo, k), dtype=torch.float32, device="cuda"),
                        torch.empty((hi - lo, k), dtype=torch.int32, device="cuda")))
        outs.append(cur)
    for r, e in enumerate(shards):
        lo, hi = slice_bounds(b, world, r)
        e.sharded_collect(k, hi - lo, out=outs[-1][r])
    torch.cuda.synchronize()
    for step, kw in enumerate(batches):
        fs, fi = full.search(k, **kw)
        fs, fi = fs.cpu().numpy(), fi.cpu().numpy()
        for r in range(world):
            lo, hi = slice_bounds(b, world, r)
            s, i = outs[step][r][0].cpu().numpy(), outs[step][r][1].cpu().numpy()
            if kw["precision"] == "fp32":
                _same_exact_ranking(fi[lo:hi], fs[lo:hi], i, s, f"step {step} shard {r}")
            else:
                rec = np.mean([len(set(i[x]) & set(fi[lo + x])) / k for x in range(hi - lo)])
                assert rec >= 0.999, (step, r, rec)
    # a chained scatter without a pending one is an error, not a hang
    with pytest.raises(bb.BrickscoreError):
        shards[0].sharded_scatter(k, collect_previous_into=outs[0][0], **batches[0])
    for e in shards + [full]:
        e.close()


def test_twenty_threads_each_with_its_own_filter(bb):
    """tests/performance/production_scalability_test.py:247-281 drives 20 threads against the recommender: every
    request carries its own valid_set_filter and must never see another request's."""
    n, d = 25216, 384
    themes, emb, _, parts, year, theme = _catalog(n, d, 0, 71)
    cat = bb.CatalogTable(synth.set_nums(n), [f"set {i}" for i in range(n)], year, parts, theme,
                          [themes.get(int(t), "Unknown") for t in theme])
    hyb = bb.HybridRecommender.from_arrays(cat, emb, themes)
    rec = hyb.content_recommender
    rng = np.random.default_rng(5)
    filters = [[cat.set_nums[j] for j in np.nonzero(rng.random(n) < 0.02 + 0.01 * t)[0]] for t in range(20)]
    targets = [cat.set_nums[int(x)] for x in rng.integers(0, n, 20)]
    want = [[r.set_num for r in rec.get_similar_sets(targets[t], 10, valid_set_filter=filters[t])] for t in range(20)]
    got, errs = [None] * 20, []

    def work(t):
        try:
            for _ in range(5):
                got[t] = [r.set_num for r in rec.get_similar_sets(targets[t], 10, valid_set_filter=filters[t])]
        except Exception as e:      # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(20)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs
    for t in range(20):
        assert got[t] == want[t]
        assert set(got[t]) <= set(filters[t])
    hyb.engine.close()


def test_train_svd_then_score_through_the_engine(bb):
    """An accelerate()d / from_arrays() recommender that (re-)trains its factors must score with the NEW item factors
    (factor-only reload of the engine operand), rebuilt user lookup and rated lists (recommendation_system.py:402-407)."""
    n, d, n_users, f = 600, 16, 80, 12
    themes, emb, _, parts, year, theme = _catalog(n, d, 0, 81)
    cat = bb.CatalogTable(synth.set_nums(n), [f"set {i}" for i in range(n)], year, parts, theme,
                          [themes.get(int(t), "Unknown") for t in theme])
    hyb = bb.HybridRecommender.from_arrays(cat, emb, themes, cf_dim=f)
    cf = hyb.collaborative_recommender
    ratings = synth.rating_matrix(n_users, n, 7).astype(np.float64)
    cf.user_item_matrix = ratings
    cf.train_svd_model(n_components=f)
    assert cf.svd_model is not None and len(cf.user_lookup) == n_users
    uf, itf, _ = bo.truncated_svd_factors(ratings, n_components=f)
    for user in (0, 17, 79):
        recs = cf.get_recommendations(user, top_k=10)
        sc = (uf[user] @ itf.T).astype(np.float64)
        sc[ratings[user] > 0] = -np.inf
        want = np.argsort(-sc, kind="stable")[:10]
        assert [r.set_num for r in recs] == [cat.set_nums[j] for j in want]
        np.testing.assert_allclose([r.score for r in recs], sc[want], rtol=2e-5, atol=2e-6)
    hyb.engine.close()


def test_device_query_producer_keeps_embeddings_on_the_gpu(bb):
    """encoder output (a CUDA tensor) goes straight into bb_search; micro-batched queries share one scan;
    the search_recommendations seam applies the reference's result filters."""
    import torch
    n, d = 6000, 384
    emb = synth.embeddings(n, d, 91, "aniso")
    themes = synth.themes_table()
    names = [themes[t] for t in sorted(themes)]
    docs = [bb.Document(f"set {i}", {"set_num": f"{i:05d}-1", "name": f"set {i}", "year": 2000 + i % 25, "num_parts": 10 + (i * 37) % 3000,
                                     "theme": names[i % len(names)], "theme_id": i % len(names)}) for i in range(n)]
    store = bb.BrickVectorStore(emb, docs, precision="fp32")
    encode = bb.minilm_shaped_encoder("cuda:0")
    prod = bb.DeviceQueryProducer(encode, store, max_pending=8)
    texts = [f"space ship with {j} engines and a crew of minifigures" for j in range(5)]
    s, i = prod.search_tensor(texts, 10)
    assert s.is_cuda and i.is_cuda and tuple(i.shape) == (5, 10)
    q_host = encode(texts).cpu().numpy()
    rs, ri = store.similarity_search_batch(q_host, 10)
    assert np.array_equal(i.cpu().numpy(), ri)
    np.testing.assert_allclose(s.cpu().numpy(), rs, rtol=1e-5, atol=2e-6)
    tickets = [prod.submit(t) for t in texts]
    got = [prod.result(t, k=10) for t in tickets]
    assert all(np.array_equal(g[1].cpu().numpy(), ri[j]) for j, g in enumerate(got))
    pq = {"semantic_query": texts[0], "filters": {"min_pieces": 500}, "confidence": 0.7, "intent": "search"}
    recs = prod.search_recommendations(pq, top_k=10)
    want = [int(x) for x in ri[0] if docs[int(x)].metadata["num_parts"] >= 500]
    assert [int(r["set_num"][:5]) for r in recs] == want and all(r["intent"] == "search" for r in recs)
    store.engine.close()


@pytest.mark.parametrize("k", [10, 100])
def test_exact_fp32_on_the_tensor_pipe_equals_the_ffma_scan(bb, k):
    """BB_PREC_FP32 with the split plane runs on tcgen05 (split-bf16 operands), re-scores in exact fp32 and proves each
    row's top-k complete; rows it cannot prove (ties at the over-fetch boundary) fall back to the FFMA scan.  The result
    must equal the FFMA path on (a) tie-free data -- where almost no row may need the fallback -- and (b) a catalog
    made of 64 copies of each row, where the fallback carries most rows."""
    n, d, f, b = 61440, 384, 64, 300
    themes, emb, itf, parts, year, theme = _catalog(n, d, f, 131)
    rng = np.random.default_rng(2)
    q = synth.embeddings(b, d, 132, "aniso") * 1.3
    u = synth.cf_factors(b, f, 133)
    self_id = rng.integers(0, n, b).astype(np.int32)
    rated = [np.sort(rng.choice(n, 30, replace=False)).astype(np.int32) for _ in range(b)]
    cons = _mix_constraints(bb, 4)
    cidx = (np.arange(b) % 4).astype(np.int32)
    # (b): every catalog row exists 64 times -> every score is a 64-way exact tie, the boundary of any over-fetch cuts a tie group
    dup = n // 64
    for label, e_mat, f_mat in (("tie-free", emb, itf), ("64-way ties", np.tile(emb[:dup], (64, 1))[:n], np.tile(itf[:dup], (64, 1))[:n])):
        eng = bb.ScoringEngine(n, d, f)
        eng.load_catalog(emb=np.ascontiguousarray(e_mat), item_factors=np.ascontiguousarray(f_mat), num_parts=parts, year=year, theme_id=theme)
        eng.set_exclusion_lists(rated)
        kw = dict(q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6,
                  constraints=cons, constraint_index=cidx, self_id=self_id, excl_lists=np.arange(b, dtype=np.int32), precision="fp32")
        s_t, i_t = eng.search(k, **kw)
        flagged = eng.info()["last_flagged_rows"]
        os.environ["BB_FP32_FFMA"] = "1"
        try:
            s_f, i_f = eng.search(k, **kw)
        finally:
            os.environ.pop("BB_FP32_FFMA", None)
        _same_exact_ranking(i_f, s_f, i_t, s_t, f"{label} k={k} flagged={flagged}")
        if label == "tie-free":
            assert flagged <= b // 20, (label, flagged)
        else:
            assert flagged >= 10, ("tied data must exercise the fallback", flagged)
        eng.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16_rescore"])
def test_pinned_host_buffers_skip_the_staging_mirror(bb, precision):
    """Host arrays that already live in pinned memory (>= 64 KB) are DMA'd straight from / into the caller's buffers,
    small or pageable ones are packed into the workspace's pinned mirror; a call may mix both.  Same results either way,
    also when the same workspace alternates between the two kinds."""
    import torch
    n, d, f, b, k = 20_000, 384, 64, 600, 40
    _, emb, itf, parts, year, theme = _catalog(n, d, f, 31)
    eng = bb.ScoringEngine(n, d, f)
    eng.load_catalog(emb=emb, item_factors=itf, num_parts=parts, year=year, theme_id=theme)
    q = synth.embeddings(b, d, 32, "aniso")
    u = synth.cf_factors(b, f, 33)
    self_id = np.arange(b, dtype=np.int32)
    cons = _mix_constraints(bb)
    cidx = (np.arange(b) % 4).astype(np.int32)
    kw = dict(w_content=0.4, w_cf=0.6, constraints=cons, constraint_index=cidx, precision=precision)
    ref_s, ref_i = eng.search(k, q_emb=q, q_cf=u, self_id=self_id, **kw)               # pageable in, pageable out
    pq, pu = torch.from_numpy(q).pin_memory(), torch.from_numpy(u).pin_memory()        # 900 KB / 150 KB: direct
    ps = torch.from_numpy(self_id).pin_memory()                                        # 2.4 KB: mirror although pinned
    out_s, out_i = torch.empty((b, k), dtype=torch.float32).pin_memory(), torch.empty((b, k), dtype=torch.int32).pin_memory()
    for rep in range(3):
        out_s.zero_(); out_i.zero_()
        eng.search(k, q_emb=pq.numpy(), q_cf=pu.numpy(), self_id=ps.numpy(), out=(out_s.numpy(), out_i.numpy()), **kw)
        assert np.array_equal(out_i.numpy(), ref_i) and np.array_equal(out_s.numpy(), ref_s), rep
        s2, i2 = eng.search(k, q_emb=pq.numpy(), q_cf=u, self_id=self_id, **kw)         # mixed: pinned + pageable in one call
        assert np.array_equal(i2, ref_i) and np.array_equal(s2, ref_s), rep
    eng.close()
