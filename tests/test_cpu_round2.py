"""CPU tests of the round-2 host logic: post-filters of the semantic-search callers against goldens produced by
executing the reference's own functions, the recommendation cache, theme-bitmap sizing, query-slice / user-shard
arithmetic, and the sharded search (q_item resolution, slice results) over gloo with the oracle-backed fake engine."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import brick_oracle as bo
import brickbrain_rec_engine_b200 as bb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nlp_golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "nlp_filter_golden.json")))


def test_oracle_post_filters_match_reference_outputs(nlp_golden):
    g = nlp_golden
    for c in g["nlp_cases"]:
        assert bo.nlp_post_filter(g["metas"], c["filters"]) == c["kept"], c["filters"]
    for c in g["hf_cases"]:
        assert bo.hf_post_filter(g["hf_results"], c["filters"]) == c["kept"], c["filters"]


def test_product_post_filters_match_reference_outputs(nlp_golden):
    g = nlp_golden
    for c in g["nlp_cases"]:
        docs = [bb.Document("", m) for m in g["metas"]]
        got = [int(d.metadata["set_num"][:5]) for d in bb.apply_document_filters(docs, c["filters"])]
        assert got == c["kept"], c["filters"]
    for c in g["hf_cases"]:
        got = [int(r["set_num"][:5]) for r in bb.apply_result_filters(g["hf_results"], c["filters"])]
        assert got == c["kept"], c["filters"]


def test_semantic_search_post_keeps_reference_order_and_shortfall(nlp_golden):
    """top-20 first, THEN the filter, then [:top_k]: fewer than top_k results when the filter removes retrieved hits."""
    metas = nlp_golden["metas"]
    ranked = list(np.random.default_rng(0).permutation(len(metas)))
    out = bo.semantic_search_post(ranked, metas, {"themes": ["city"]}, top_k=10, retrieve_k=20)
    want = [i for i in ranked[:20] if "city" in metas[i]["theme"].lower()][:10]
    assert out == want and len(out) < 10


def test_recommendation_cache_semantics():
    now = [1000.0]
    c = bb.RecommendationCache(k_max=8, clock=lambda: now[0])
    assert c.get(5) is None
    c.bulk_upsert(np.array([5, 9]), np.array([[3, 1, 2, -1], [7, 8, -1, -1]], np.int32),
                  np.array([[0.5, 0.9, 0.7, 0], [0.2, 0.4, 0, 0]], np.float32), "hybrid")
    ids, sc = c.get(5, "hybrid", top_k=2)
    assert ids.tolist() == [1, 2] and np.allclose(sc, [0.9, 0.7])          # ORDER BY score DESC LIMIT top_k
    assert c.get(5, "content") is None                                      # keyed by recommendation_type
    c.put(5, [4], [0.1], "hybrid")                                          # DELETE + INSERT: replaces the user's rows
    assert c.get(5, "hybrid", 10)[0].tolist() == [4]
    assert c.get(0) is None and c.get(None) is None
    now[0] += 24 * 3600 + 1                                                 # expires_at = created + 24 h
    assert c.get(9) is None
    c.put(0, [1], [1.0])                                                    # falsy user id: not cached
    assert len(c) == 2


def test_theme_bits_are_sized_or_rejected():
    themes = {5: "City", 1500: "Far Theme", 40: "Space"}
    assert bb.theme_bits_for(themes) == 1504
    comp = bb.ConstraintCompiler(themes, theme_bits=1024)
    with pytest.raises(ValueError):
        comp.compile([bb.HardConstraint(bb.ConstraintType.THEMES_REQUIRED, ["Far"])])
    big = bb.ConstraintCompiler(themes, theme_bits=bb.theme_bits_for(themes))
    cc = big.compile([bb.HardConstraint(bb.ConstraintType.THEMES_REQUIRED, ["Far"])])
    assert cc.theme_bitmap.shape[0] == 47 and (cc.theme_bitmap[1500 >> 5] >> (1500 & 31)) & 1
    from fake_engine import FakeEngine
    with pytest.raises(ValueError):
        bb.HardConstraintFilter(FakeEngine(10, 4), [f"{i}" for i in range(10)], themes)


def test_slice_and_user_shard_bounds():
    for b, w in ((8192, 8), (700, 4), (5, 8), (1, 2)):
        s = [bb.slice_bounds(b, w, r) for r in range(w)]
        assert s[0][0] == 0 and s[-1][1] == b and all(s[i][1] == s[i + 1][0] for i in range(w - 1))
        assert all(hi - lo <= -(-b // w) for lo, hi in s)
    u = [bb.user_shard(1_000_003, 8, r) for r in range(8)]
    assert u[0][0] == 0 and u[-1][1] == 1_000_003 and all(u[i][1] == u[i + 1][0] for i in range(7))


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "oracle")); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import brick_oracle as bo, synth
import brickbrain_rec_engine_b200 as bb
from fake_engine import FakeEngine
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
n, d, f, b, k = 2003, 32, 8, 10, 6
emb = synth.embeddings(n, d, 1, "aniso"); itf = synth.cf_factors(n, f, 2)
lo, hi = bb.shard_bounds(n, world, rank)
eng = FakeEngine(hi - lo, d, f, id_offset=lo)
eng.load_catalog(emb=emb[lo:hi], item_factors=itf[lo:hi])
rated = [np.sort(np.random.default_rng(50 + j).choice(n, 20, replace=False)).astype(np.int32) for j in range(b)]
eng.set_exclusion_lists(rated)
u = synth.cf_factors(b, f, 4)
liked = ((np.arange(b, dtype=np.int32) * 211) % n).astype(np.int32)       # rows owned by different shards
ss = bb.ShardedSearch(eng)
assert ss.exchange == "nccl"                                               # (no communication blocks on CPU)
s, i = ss.search(k, q_item=liked, q_cf=u, w_content=0.4, w_cf=0.6, self_id=liked, excl_lists=np.arange(b, dtype=np.int32))
e_unit = bo.l2_normalize_rows(emb)
ref = bo.fused_hybrid_scores(e_unit[liked], u, e_unit, itf, np.full(b, .4, np.float32), np.full(b, .6, np.float32))
rid, rs = bo.fused_hybrid_topk(ref, k, self_ids=liked, exclusions=rated)
assert np.array_equal(np.asarray(i), rid), (rank, i, rid)
np.testing.assert_allclose(np.asarray(s), rs, rtol=1e-5)
s2, i2 = ss.search(k, result="slice", q_item=liked, q_cf=u, w_content=0.4, w_cf=0.6, self_id=liked, excl_lists=np.arange(b, dtype=np.int32))
a, z = bb.slice_bounds(b, world, rank)
assert np.array_equal(np.asarray(i2), rid[a:z])
# without communication blocks (CPU / gloo) a pipelined search is a plain slice search and flush() has nothing to do
s3, i3 = ss.search_pipelined(k, q_item=liked, q_cf=u, w_content=0.4, w_cf=0.6, self_id=liked, excl_lists=np.arange(b, dtype=np.int32))
ss.flush()
assert np.array_equal(np.asarray(i3), rid[a:z])
try:
    ss.search(k, q_item=np.array([n + 5], dtype=np.int32))
    raise SystemExit("an id that no shard owns must be rejected")
except ValueError:
    pass
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_q_item_and_slices_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out[-3000:]
