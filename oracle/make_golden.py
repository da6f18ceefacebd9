"""Generates tests/golden/*.npz by EXECUTING the unmodified reference.

Run in the dev container only (needs /root/reference):

    python oracle/make_golden.py

Every output array in the golden files is produced by the reference's own
functions (``get_similar_sets``, ``get_recommendations``,
``_combine_recommendations``, ``HybridRecommender.get_recommendations``,
``HardConstraintFilter.apply_constraints`` -- the latter executing the
reference's generated SQL on SQLite, see ref_harness.py).  Inputs are stored in
exactly-representable form so the GPU box can replay them without the reference.
"""
from __future__ import annotations

import json
import os
import sys
from datetime import datetime

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402
import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def pad_lists(lists, k, fill):
    out = np.full((len(lists), k), fill, dtype=np.asarray(fill).dtype if not isinstance(fill, int) else np.int64)
    for i, l in enumerate(lists):
        out[i, : len(l)] = l
    return out


def recs_to_arrays(recs, index_of, k):
    ids = [index_of[r.set_num] for r in recs]
    sc = [r.score for r in recs]
    a = np.full(k, -1, dtype=np.int64)
    s = np.full(k, np.nan, dtype=np.float64)
    a[: len(ids)] = ids
    s[: len(sc)] = sc
    return a, s


def ragged(list_of_arrays):
    off = np.zeros(len(list_of_arrays) + 1, dtype=np.int64)
    for i, a in enumerate(list_of_arrays):
        off[i + 1] = off[i] + len(a)
    flat = np.concatenate([np.asarray(a, dtype=np.int64) for a in list_of_arrays]) if list_of_arrays else np.zeros(0, np.int64)
    return flat, off


def main():
    os.makedirs(OUT, exist_ok=True)
    rs, hcf = rh.load_reference()
    themes = synth.themes_table()
    now_year = datetime.now().year

    # ------------------------------------------------------------------ masks (hcf.py SQL on SQLite)
    n = 4000
    num_parts, year, theme_id = synth.metadata(n, seed=11, themes=themes, null_theme_frac=0.03)
    num_parts[:7] = [0, 0, 50, 51, 500, 501, 12000]        # include num_parts = 0 rows (fail "> 0")
    year[:4] = [now_year - 5, now_year - 6, 2018, 2017]
    set_nums = synth.set_nums(n)
    names = [f"Set {i}" for i in range(n)]
    owned = synth.user_lists(6, n, 40, seed=21)
    wish = synth.user_lists(6, n, 25, seed=22)
    con = rh.SqliteShimConnection(set_nums, names, year, theme_id, num_parts, themes, owned, wish)
    index_of = {s: i for i, s in enumerate(set_nums)}
    filt = hcf.HardConstraintFilter(con)
    scen_names, scen_specs, valid_lists = [], [], []
    for name, kw in synth.SCENARIO_KWARGS:
        cons = filt.create_constraint_set(**kw)
        res = filt.apply_constraints(cons)
        scen_names.append("kw:" + name)
        scen_specs.append({"kwargs": kw, "pairs": [[c.constraint_type.value, c.value] for c in cons]})
        valid_lists.append([index_of[s] for s in res.valid_set_nums])
    for name, pairs in synth.SCENARIO_LISTS:
        cons = [hcf.HardConstraint(hcf.ConstraintType(t), v) for t, v in pairs]
        res = filt.apply_constraints(cons)
        scen_names.append("list:" + name)
        scen_specs.append({"pairs": [[t, v] for t, v in pairs]})
        valid_lists.append([index_of[s] for s in res.valid_set_nums])
    flat, off = ragged(valid_lists)
    np.savez_compressed(
        os.path.join(OUT, "mask_golden.npz"),
        num_parts=num_parts, year=year, theme_id=theme_id,
        themes_json=json.dumps({str(k): v for k, v in themes.items()}),
        owned_json=json.dumps({str(k): v.tolist() for k, v in owned.items()}),
        wish_json=json.dumps({str(k): v.tolist() for k, v in wish.items()}),
        scenario_names=np.array(scen_names), scenario_specs=json.dumps(scen_specs),
        valid_flat=flat, valid_off=off, now_year=now_year)
    print("mask_golden:", len(scen_names), "scenarios; valid counts",
          [len(v) for v in valid_lists])

    # ------------------------------------------------------------------ content, fp32 384-d (int8-exact inputs)
    n = 1536
    d = 384
    x = synth.embeddings(n, d, seed=31, kind="aniso")
    q8 = np.clip(np.rint(x * 600.0), -127, 127).astype(np.int8)
    feat = (q8.astype(np.float32) / np.float32(64.0))
    np_, yr_, th_ = synth.metadata(n, seed=32, themes=themes)
    sn = synth.set_nums(n)
    frame = rh.make_set_frame(sn, [f"Set {i}" for i in range(n)], yr_, np.maximum(th_, 0), np_,
                              [themes.get(int(t), "Unknown") for t in th_])
    idx_of = {s: i for i, s in enumerate(sn)}
    content = rh.build_content(rs, None, frame, feat)
    targets = np.arange(0, n, 64)[:24]
    k = 10
    rng = np.random.default_rng(33)
    valid_half = np.sort(rng.choice(n, size=n // 3, replace=False))
    ids_nf, sc_nf, ids_f, sc_f = [], [], [], []
    for t in targets:
        a, s = recs_to_arrays(content.get_similar_sets(sn[t], k), idx_of, k)
        ids_nf.append(a); sc_nf.append(s)
        a, s = recs_to_arrays(content.get_similar_sets(sn[t], k, valid_set_filter=[sn[i] for i in valid_half]), idx_of, k)
        ids_f.append(a); sc_f.append(s)
    np.savez_compressed(os.path.join(OUT, "content_fp32_golden.npz"), q8=q8, scale=np.float32(64.0),
                        targets=targets, k=k, valid_ids=valid_half,
                        ids_nofilter=np.array(ids_nf), scores_nofilter=np.array(sc_nf),
                        ids_filter=np.array(ids_f), scores_filter=np.array(sc_f))
    print("content_fp32_golden: targets", len(targets))

    # ------------------------------------------------------------------ content, float64 ~10-d features (reference's own shape)
    n64 = 2000
    rng = np.random.default_rng(41)
    np64, yr64, th64 = synth.metadata(n64, seed=42, themes=themes)
    numeric = np.stack([np64, rng.integers(1, 400, n64), rng.integers(1, 40, n64),
                        rng.random(n64), rng.random(n64), th64], axis=1).astype(np.float64)
    numeric = (numeric - numeric.mean(0)) / numeric.std(0)             # StandardScaler (rs.py:186)
    cats = np.array([rh.size_category(int(p)) for p in np64])
    onehot = np.stack([(cats == c) for c in ("large", "medium", "small", "xl")], axis=1).astype(np.float64)
    feat64 = np.hstack([numeric, onehot])                              # rs.py:191
    sn64 = synth.set_nums(n64)
    frame64 = rh.make_set_frame(sn64, [f"Set {i}" for i in range(n64)], yr64, np.maximum(th64, 0), np64,
                                [themes.get(int(t), "Unknown") for t in th64])
    idx64 = {s: i for i, s in enumerate(sn64)}
    c64 = rh.build_content(rs, None, frame64, feat64)
    t64 = np.arange(5, n64, 97)[:20]
    ids64, sc64 = [], []
    for t in t64:
        a, s = recs_to_arrays(c64.get_similar_sets(sn64[t], k), idx64, k)
        ids64.append(a); sc64.append(s)
    np.savez_compressed(os.path.join(OUT, "content_f64_golden.npz"), feat=feat64, targets=t64, k=k,
                        ids=np.array(ids64), scores=np.array(sc64))

    # ------------------------------------------------------------------ FAISS fixture rows (real MiniLM embeddings; inputs only)
    raw = open(os.path.join(rh.REFERENCE_ROOT, "test_embeddings", "index.faiss"), "rb").read()
    assert raw[:4] == b"IxF2"
    fx = np.frombuffer(raw, dtype="<f4", count=3840, offset=45).reshape(10, 384).copy()
    snf = synth.set_nums(10)
    framef = rh.make_set_frame(snf, [f"Set {i}" for i in range(10)], np.full(10, 2020), np.ones(10), np.full(10, 100),
                               ["x"] * 10)
    cf_ = rh.build_content(rs, None, framef, fx)
    idsf, scf = [], []
    for t in range(10):
        a, s = recs_to_arrays(cf_.get_similar_sets(snf[t], 5), {s_: i for i, s_ in enumerate(snf)}, 5)
        idsf.append(a); scf.append(s)
    np.savez_compressed(os.path.join(OUT, "faiss_fixture_golden.npz"), emb=fx, ids=np.array(idsf), scores=np.array(scf))

    # ------------------------------------------------------------------ CF: real TruncatedSVD from the reference, then get_recommendations
    n_users = 64
    ratings = synth.rating_matrix(n_users, n, seed=51)
    user_ids = list(range(1, n_users + 1))
    cf = rh.build_cf(rs, None, frame, user_ids, ratings, None, None)
    cf.svd_model = None
    cf.train_svd_model()                                            # rs.py:368-409, random_state=42
    uf, itf = np.asarray(cf.user_factors), np.asarray(cf.item_factors)
    users = user_ids[::5][:12]
    k_cf = 10
    ids_c, sc_c, ids_cf_f, sc_cf_f = [], [], [], []
    for u in users:
        a, s = recs_to_arrays(cf.get_recommendations(u, k_cf), idx_of, k_cf)
        ids_c.append(a); sc_c.append(s)
        a, s = recs_to_arrays(cf.get_recommendations(u, k_cf, valid_set_filter=[sn[i] for i in valid_half]), idx_of, k_cf)
        ids_cf_f.append(a); sc_cf_f.append(s)
    np.savez_compressed(os.path.join(OUT, "cf_golden.npz"), ratings=ratings.astype(np.int8), user_ids=np.array(user_ids),
                        user_factors=uf, item_factors=itf, users=np.array(users), k=k_cf, valid_ids=valid_half,
                        ids_nofilter=np.array(ids_c), scores_nofilter=np.array(sc_c),
                        ids_filter=np.array(ids_cf_f), scores_filter=np.array(sc_cf_f))
    print("cf_golden: factors", uf.shape, itf.shape)

    # ------------------------------------------------------------------ CF training, more users than items (sklearn's
    # randomized SVD does NOT transpose then): the reference's train_svd_model on a 300 x 96 pivot
    n_tall_users, n_tall_items = 300, 96
    tall = synth.rating_matrix(n_tall_users, n_tall_items, seed=52, lo=8, hi=24)
    frame_t = rh.make_set_frame(synth.set_nums(n_tall_items), [f"Set {i}" for i in range(n_tall_items)],
                                np.full(n_tall_items, 2020), np.ones(n_tall_items), np.full(n_tall_items, 100), ["x"] * n_tall_items)
    cf_t = rh.build_cf(rs, None, frame_t, list(range(1, n_tall_users + 1)), tall, None, None)
    cf_t.svd_model = None
    cf_t.train_svd_model()
    np.savez_compressed(os.path.join(OUT, "svd_tall_golden.npz"), ratings=tall.astype(np.int8),
                        user_factors=np.asarray(cf_t.user_factors), item_factors=np.asarray(cf_t.item_factors))
    print("svd_tall_golden: factors", np.asarray(cf_t.user_factors).shape, np.asarray(cf_t.item_factors).shape)

    # ------------------------------------------------------------------ hybrid: _combine_recommendations + end-to-end with constraints
    con2 = rh.SqliteShimConnection(sn, [f"Set {i}" for i in range(n)], yr_, th_, np_, themes,
                                   synth.user_lists(8, n, 30, seed=61), synth.user_lists(8, n, 20, seed=62))
    hyb = rh.build_hybrid(rs, con2, content, cf)
    hyb.constraint_filter = hcf.HardConstraintFilter(con2)
    cases = []
    k_h = 10
    e2e = [
        dict(user_id=1, liked=int(targets[1]), kw=None, weights=None),
        dict(user_id=6, liked=int(targets[3]), kw=dict(pieces_max=1000, year_min=2010), weights=None),
        dict(user_id=11, liked=int(targets[5]), kw=dict(pieces_min=100, required_themes=["City", "Creator", "Star Wars"]), weights=None),
        dict(user_id=16, liked=int(targets[7]), kw=dict(price_max=150.0, age_min=8, excluded_themes=["Duplo"]), weights=(0.7, 0.5)),
        dict(user_id=21, liked=int(targets[9]), kw=dict(pieces_min=5000, price_max=5.0), weights=None),      # nothing valid
        dict(user_id=None, liked=int(targets[2]), kw=dict(pieces_max=800), weights=None),                    # content only
        dict(user_id=26, liked=None, kw=dict(year_min=2000), weights=None),                                  # CF only
        dict(user_id=5, liked=int(targets[4]), kw=dict(exclude_owned=True, exclude_wishlisted=True, user_id=5, pieces_max=3000), weights=None),
    ]
    for c in e2e:
        if c["weights"]:
            hyb.set_weights(*c["weights"])
        else:
            hyb.content_weight, hyb.collaborative_weight = 0.4, 0.6
        cons = hyb.constraint_filter.create_constraint_set(**c["kw"]) if c["kw"] else None
        recs, cres = hyb.get_recommendations(user_id=c["user_id"], liked_set=(sn[c["liked"]] if c["liked"] is not None else None),
                                             top_k=k_h, constraints=cons)
        a, s = recs_to_arrays(recs, idx_of, k_h)
        cases.append(dict(user_id=c["user_id"], liked=c["liked"], kwargs=c["kw"], weights=[hyb.content_weight, hyb.collaborative_weight],
                          ids=a.tolist(), scores=[None if np.isnan(v) else float(v) for v in s],
                          n_valid=(len(cres.valid_set_nums) if cres is not None else None)))
    # pure combine on explicit lists
    hyb.content_weight, hyb.collaborative_weight = 0.4, 0.6
    cl = content.get_similar_sets(sn[int(targets[1])], 2 * k_h)
    fl = cf.get_recommendations(1, 2 * k_h)
    comb = hyb._combine_recommendations(cl, fl, k_h)
    combine = dict(content_ids=[idx_of[r.set_num] for r in cl], content_scores=[r.score for r in cl],
                   cf_ids=[idx_of[r.set_num] for r in fl], cf_scores=[r.score for r in fl],
                   out_ids=[idx_of[r.set_num] for r in comb], out_scores=[r.score for r in comb])
    np.savez_compressed(os.path.join(OUT, "hybrid_golden.npz"), num_parts=np_, year=yr_, theme_id=th_,
                        owned_json=json.dumps({str(k_): v.tolist() for k_, v in synth.user_lists(8, n, 30, seed=61).items()}),
                        wish_json=json.dumps({str(k_): v.tolist() for k_, v in synth.user_lists(8, n, 20, seed=62).items()}),
                        themes_json=json.dumps({str(k_): v for k_, v in themes.items()}),
                        cases_json=json.dumps(cases), combine_json=json.dumps(combine), k=k_h, now_year=now_year)
    print("hybrid_golden:", [(c["n_valid"], sum(i >= 0 for i in c["ids"])) for c in cases])
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
