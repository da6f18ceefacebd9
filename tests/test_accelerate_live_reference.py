"""``accelerate()`` on a LIVE reference object (dev container only: needs ``/root/reference``).  The unmodified
``HybridRecommender`` of ``recommendation_system.py`` is built over the SQLite shim with injected features, the reference's
own ``train_svd_model`` trains its factors, ``accelerate(reference_hybrid, themes)`` adopts that state, and the same calls
are made on both objects: the accelerated adapters must return what the reference returns (same sets in the same order
wherever the reference's scores are not tied, same scores, same violation / empty-result behaviour).  The compute behind
the adapters is the oracle-backed ``FakeEngine`` here (CPU); the GPU engine is held to the same goldens in
``test_gpu_engine.py``.  Test infrastructure: nothing here is imported by the product."""
import os

import numpy as np
import pytest

import synth
import brickbrain_rec_engine_b200 as bb
from fake_engine import FakeEngine

REF = "/root/reference/src/scripts"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not present on this machine")


@pytest.fixture(scope="module")
def live():
    import ref_harness as rh
    rs, hcf = rh.load_reference()
    themes = synth.themes_table()
    n, d, n_users = 900, 48, 40
    x = synth.embeddings(n, d, seed=131, kind="aniso")
    feat = np.clip(np.rint(x * 600.0), -127, 127).astype(np.int8).astype(np.float32) / np.float32(64.0)
    parts, year, theme = synth.metadata(n, seed=132, themes=themes)
    sn = synth.set_nums(n)
    frame = rh.make_set_frame(sn, [f"Set {i}" for i in range(n)], year, np.maximum(theme, 0), parts,
                              [themes.get(int(t), "Unknown") for t in theme])
    owned, wish = synth.user_lists(8, n, 30, seed=161), synth.user_lists(8, n, 20, seed=162)
    ratings = synth.rating_matrix(n_users, n, seed=151)
    user_ids = list(range(1, n_users + 1))
    triples = [(user_ids[u], i, ratings[u, i]) for u in range(n_users) for i in np.nonzero(ratings[u] > 0)[0]]
    con = rh.SqliteShimConnection(sn, [f"Set {i}" for i in range(n)], year, np.maximum(theme, 0), parts, themes, owned, wish,
                                  ratings=triples)
    content = rh.build_content(rs, con, frame, feat)
    cf = rh.build_cf(rs, con, frame, user_ids, ratings, None, None)
    cf.svd_model = None
    cf.train_svd_model()                                     # the reference's own training (recommendation_system.py:368-409)
    hyb = rh.build_hybrid(rs, con, content, cf)
    hyb.constraint_filter = hcf.HardConstraintFilter(con)
    return dict(hyb=hyb, themes=themes, sn=sn, owned=owned, wish=wish, ratings=ratings, user_ids=user_ids)


@pytest.fixture()
def accelerated(live, monkeypatch):
    import brickbrain_rec_engine_b200.recommenders as rec
    monkeypatch.setattr(rec, "ScoringEngine", type("E", (FakeEngine,), {"pack_bitmap": staticmethod(bb.ScoringEngine.pack_bitmap),
                                                                       "unpack_bitmap": staticmethod(bb.ScoringEngine.unpack_bitmap)}))
    return bb.accelerate(live["hyb"], live["themes"], owned=live["owned"], wishlisted=live["wish"])


def _same_ranking(ref, got, what):
    """Same scores position by position; same sets in the same order except inside runs of (near-)equal reference
    scores -- the reference sorts float64 scores whose ties (items with identical rating columns get identical factors)
    it breaks by iteration order, and a tie run that straddles the k-th place may admit an item from beyond the cut."""
    rs_, gs_ = [r.score for r in ref], [r.score for r in got]
    assert len(ref) == len(got), what
    np.testing.assert_allclose(gs_, rs_, rtol=2e-5, atol=2e-6, err_msg=what)
    names = [r.set_num for r in ref]
    tol = lambda v: 2e-5 * max(1.0, abs(v))
    for p, (a, b) in enumerate(zip(ref, got)):
        if a.set_num != b.set_num:
            inside = b.set_num in names and abs(rs_[p] - rs_[names.index(b.set_num)]) <= tol(rs_[p])
            trailing = abs(rs_[p] - rs_[-1]) <= tol(rs_[p])
            assert inside or trailing, (what, p, a.set_num, b.set_num)
    for a, b in zip(ref, got):
        if a.set_num == b.set_num:
            assert (a.name, a.year, a.num_parts, a.theme_name) == (b.name, b.year, b.num_parts, b.theme_name), what
            assert a.reasons == b.reasons and a.constraint_violations == b.constraint_violations, (what, a.reasons, b.reasons)


def test_adopted_state(live, accelerated):
    ref, acc = live["hyb"], accelerated
    assert isinstance(acc, bb.HybridRecommender)
    assert acc.content_weight == ref.content_weight and acc.collaborative_weight == ref.collaborative_weight
    assert acc.content_recommender.catalog.set_nums == live["sn"]
    cfr = acc.collaborative_recommender
    assert cfr.user_lookup == {u: i for i, u in enumerate(live["user_ids"])}
    for i, u in enumerate(live["user_ids"][:10]):
        assert np.array_equal(np.asarray(cfr.rated[u]), np.nonzero(live["ratings"][i] > 0)[0])


def test_content_and_cf_calls_match_the_live_reference(live, accelerated):
    ref, acc, sn = live["hyb"], accelerated, live["sn"]
    rng = np.random.default_rng(5)
    valid = [sn[i] for i in np.sort(rng.choice(len(sn), size=len(sn) // 3, replace=False))]
    for t in (3, 77, 410, 899):
        _same_ranking(ref.content_recommender.get_similar_sets(sn[t], 10), acc.content_recommender.get_similar_sets(sn[t], 10), f"content {t}")
        _same_ranking(ref.content_recommender.get_similar_sets(sn[t], 7, valid_set_filter=valid),
                      acc.content_recommender.get_similar_sets(sn[t], 7, valid_set_filter=valid), f"content filtered {t}")
    for u in (1, 9, 23, 40):
        _same_ranking(ref.collaborative_recommender.get_recommendations(u, 10), acc.collaborative_recommender.get_recommendations(u, 10), f"cf {u}")
        _same_ranking(ref.collaborative_recommender.get_recommendations(u, 6, valid_set_filter=valid),
                      acc.collaborative_recommender.get_recommendations(u, 6, valid_set_filter=valid), f"cf filtered {u}")
    assert acc.content_recommender.get_similar_sets("no-such-set", 5) == ref.content_recommender.get_similar_sets("no-such-set", 5) == []


@pytest.mark.parametrize("case", [
    dict(user_id=2, liked=5, kw=None),
    dict(user_id=7, liked=120, kw=dict(pieces_max=1000, year_min=2010)),
    dict(user_id=12, liked=300, kw=dict(pieces_min=100, required_themes=["City", "Creator", "Star Wars"])),
    dict(user_id=17, liked=640, kw=dict(price_max=150.0, age_min=8, excluded_themes=["Duplo"])),
    dict(user_id=21, liked=33, kw=dict(pieces_min=5000, price_max=5.0)),                  # nothing valid
    dict(user_id=None, liked=48, kw=dict(pieces_max=800)),                                # content only
    dict(user_id=26, liked=None, kw=dict(year_min=2000)),                                 # CF only
    dict(user_id=5, liked=71, kw=dict(exclude_owned=True, exclude_wishlisted=True, user_id=5, pieces_max=3000)),
])
def test_hybrid_calls_match_the_live_reference(live, accelerated, case):
    ref, acc, sn = live["hyb"], accelerated, live["sn"]
    liked = sn[case["liked"]] if case["liked"] is not None else None
    rc = ref.constraint_filter.create_constraint_set(**case["kw"]) if case["kw"] else None
    ac = acc.constraint_filter.create_constraint_set(**case["kw"]) if case["kw"] else None
    r_recs, r_res = ref.get_recommendations(user_id=case["user_id"], liked_set=liked, top_k=10, constraints=rc)
    a_recs, a_res = acc.get_recommendations(user_id=case["user_id"], liked_set=liked, top_k=10, constraints=ac)
    _same_ranking(r_recs, a_recs, str(case))
    if rc:
        assert sorted(a_res.valid_set_nums) == sorted(r_res.valid_set_nums), case
        assert len(a_res.applied_constraints) == len(r_res.applied_constraints)
        for key in ("input_sets", "output_sets", "constraint_count"):
            assert a_res.performance_stats[key] == r_res.performance_stats[key], (case, key)
        assert len(a_res.violations) == len(r_res.violations)
        for rv, av in zip(r_res.violations, a_res.violations):          # hard_constraint_filter.py:534-637: counts, texts, advice
            assert (av.constraint.constraint_type.value, av.constraint.value) == (rv.constraint.constraint_type.value, rv.constraint.value)
            assert (av.violating_count, av.total_count, av.message) == (rv.violating_count, rv.total_count, rv.message), case
            assert av.suggested_alternatives == rv.suggested_alternatives, case
    else:
        assert (a_res is None) == (r_res is None)


def test_request_objects_are_interchangeable(live, accelerated):
    """``get_recommendations_from_request`` takes the reference's own ``RecommendationRequest`` (and the package's twin has
    the same fields, soft preferences included)."""
    import dataclasses
    import ref_harness as rh
    rs, _ = rh.load_reference()
    ref, acc, sn = live["hyb"], accelerated, live["sn"]
    kw = dict(user_id=7, liked_set=sn[120], top_k=6, pieces_max=1000, year_min=2005, preferred_themes=["City"], budget_preference=80.0)
    assert [f.name for f in dataclasses.fields(bb.RecommendationRequest)] == [f.name for f in dataclasses.fields(rs.RecommendationRequest)]
    r_recs, r_res = ref.get_recommendations_from_request(rs.RecommendationRequest(**kw))
    a_recs, a_res = acc.get_recommendations_from_request(rs.RecommendationRequest(**kw))       # the reference's class
    b_recs, b_res = acc.get_recommendations_from_request(bb.RecommendationRequest(**kw))       # the package's twin
    _same_ranking(r_recs, a_recs, "request (reference class)")
    _same_ranking(r_recs, b_recs, "request (package class)")
    assert sorted(a_res.valid_set_nums) == sorted(r_res.valid_set_nums) == sorted(b_res.valid_set_nums)


def test_public_types_have_the_reference_fields():
    """Dataclass fields (names, order) and enum members of the adapter types equal the reference's; the package's
    ``ConstraintResult`` only APPENDS its device bitmap."""
    import dataclasses
    import ref_harness as rh
    rs, hcf = rh.load_reference()
    names = lambda c: [f.name for f in dataclasses.fields(c)]
    assert names(bb.RecommendationResult) == names(rs.RecommendationResult)
    assert names(bb.RecommendationRequest) == names(rs.RecommendationRequest)
    assert names(bb.HardConstraint) == names(hcf.HardConstraint)
    assert names(bb.ConstraintViolation) == names(hcf.ConstraintViolation)
    assert names(bb.ConstraintResult)[:len(names(hcf.ConstraintResult))] == names(hcf.ConstraintResult)
    assert [(e.name, e.value) for e in bb.ConstraintType] == [(e.name, e.value) for e in hcf.ConstraintType]
    assert [(e.name, e.value) for e in bb.ConstraintSeverity] == [(e.name, e.value) for e in hcf.ConstraintSeverity]


def test_method_signatures_equal_the_reference():
    """the seams callers already use keep their parameter names, order, kinds and defaults."""
    import inspect
    import ref_harness as rh
    rs, hcf = rh.load_reference()
    pairs = [(rs.ContentBasedRecommender.get_similar_sets, bb.ContentBasedRecommender.get_similar_sets),
             (rs.CollaborativeFilteringRecommender.get_recommendations, bb.CollaborativeFilteringRecommender.get_recommendations),
             (rs.CollaborativeFilteringRecommender.train_svd_model, bb.CollaborativeFilteringRecommender.train_svd_model),
             (rs.HybridRecommender.get_recommendations, bb.HybridRecommender.get_recommendations),
             (rs.HybridRecommender.get_recommendations_from_request, bb.HybridRecommender.get_recommendations_from_request),
             (rs.HybridRecommender._combine_recommendations, bb.HybridRecommender._combine_recommendations),
             (rs.HybridRecommender.set_weights, bb.HybridRecommender.set_weights),
             (hcf.HardConstraintFilter.create_constraint_set, bb.HardConstraintFilter.create_constraint_set),
             (hcf.HardConstraintFilter.apply_constraints, bb.HardConstraintFilter.apply_constraints)]
    for r, o in pairs:
        a = [(p.name, p.default, p.kind) for p in inspect.signature(r).parameters.values()]
        b = [(p.name, p.default, p.kind) for p in inspect.signature(o).parameters.values()]
        assert a == b, r.__qualname__


def test_violation_analysis_matches_the_live_reference(live, accelerated):
    """``apply_constraints`` alone, on constraint sets that eliminate most or all of the catalog (the reference analyses
    every constraint whose elimination rate is high: counts, message and suggested alternatives)."""
    ref, acc = live["hyb"].constraint_filter, accelerated.constraint_filter
    for kw in (dict(required_themes=["Nonexistent Theme"], pieces_max=100), dict(pieces_min=5000, price_max=5.0),
               dict(year_min=2024, age_max=6), dict(max_complexity="simple", min_complexity="complex"),
               dict(excluded_themes=["City"], must_be_available=True, pieces_max=60)):
        r = ref.apply_constraints(ref.create_constraint_set(**kw))
        a = acc.apply_constraints(acc.create_constraint_set(**kw))
        assert sorted(a.valid_set_nums) == sorted(r.valid_set_nums), kw
        assert len(a.violations) == len(r.violations), kw
        for rv, av in zip(r.violations, a.violations):
            assert (av.violating_count, av.total_count, av.message, av.suggested_alternatives) == \
                   (rv.violating_count, rv.total_count, rv.message, rv.suggested_alternatives), kw


def test_cold_start_matches_the_live_reference(live, accelerated):
    """Unknown user: the reference answers from SQL aggregates over ``user_interactions`` (recommendation_system.py:485-534,
    executed here on the SQLite shim); the adapter answers from the rating summary it adopted from the pivot table."""
    import warnings
    ref, acc, sn = live["hyb"].collaborative_recommender, accelerated.collaborative_recommender, live["sn"]
    rated_any = np.nonzero((live["ratings"] > 0).sum(0) >= 3)[0]
    valid = [sn[i] for i in rated_any[::2]] + [sn[i] for i in range(0, len(sn), 9)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                       # pandas: "only supports SQLAlchemy connectable"
        for flt in (None, valid):
            r = ref.get_recommendations(987654, 8, valid_set_filter=flt)
            a = acc.get_recommendations(987654, 8, valid_set_filter=flt)
            assert len(r) == len(a) > 0
            assert [(x.set_num, x.score, x.reasons, x.name, x.year, x.num_parts, x.theme_name) for x in r] == \
                   [(x.set_num, x.score, x.reasons, x.name, x.year, x.num_parts, x.theme_name) for x in a], flt
