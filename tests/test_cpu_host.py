"""CPU-only tests: the C ABI loads and exports every declared symbol, host-side logic
(constraint compilation, adapters, sharding plumbing) behaves like the reference.
Compute never runs here -- a ``FakeEngine`` (oracle-backed, tests only) stands in."""
import json
import os
import re
import struct
import subprocess
import sys

import numpy as np
import pytest

import brick_oracle as bo
import synth
import brickbrain_rec_engine_b200 as bb
from brickbrain_rec_engine_b200 import _lib as L
from fake_engine import FakeEngine, eval_descriptor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "brickscore.h")).read()
    declared = set(re.findall(r"BB_API\s+[\w\s\*]+?\b(bb_\w+)\s*\(", header))
    assert len(declared) >= 17
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in brickscore.h but not exported"
    assert declared == set(L.PROTOTYPES), "ctypes prototypes and the header diverge"
    assert lib.bb_version().startswith(b"brickscore")


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(L.bb_constraint) == 36 and L.CONSTRAINT_DTYPE.itemsize == 36
    assert L.bb_constraint.excl_list.offset == 20 and L.bb_constraint.flags.offset == 32
    assert C.sizeof(L.bb_config) == 56
    d = L.no_constraint()
    assert d["parts_min"] == -2**31 and d["parts_max"] == 2**31 - 1 and d["theme_set"] == -1


def test_no_cpu_fallback():
    """Without a device bb_create must fail loudly (BB_ERR_CUDA); nothing computes on the CPU."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(bb.BrickscoreError) as e:
        bb.ScoringEngine(16, 8)
    assert e.value.code == L.BB_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "brickbrain_rec_engine_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "brick_oracle" not in src and "ref_harness" not in src and "import synth" not in src, fn


# ------------------------------------------------------------------------------------------ constraint compilation
def _golden_masks(golden_dir):
    g = np.load(os.path.join(golden_dir, "mask_golden.npz"), allow_pickle=False)
    themes = {int(k): v for k, v in json.loads(str(g["themes_json"])).items()}
    owned = {int(k): np.array(v, dtype=np.int32) for k, v in json.loads(str(g["owned_json"])).items()}
    wish = {int(k): np.array(v, dtype=np.int32) for k, v in json.loads(str(g["wish_json"])).items()}
    return g, themes, owned, wish


def test_compiled_descriptors_reproduce_reference_sql(golden_dir):
    """HardConstraintFilter (host compile + mask) vs the valid sets the reference's SQL returned."""
    g, themes, owned, wish = _golden_masks(golden_dir)
    n = g["num_parts"].shape[0]
    eng = FakeEngine(n, 8)
    eng.load_catalog(emb=np.ones((n, 8), np.float32), num_parts=g["num_parts"], year=g["year"], theme_id=g["theme_id"])
    filt = bb.HardConstraintFilter(eng, synth.set_nums(n), themes, owned=owned, wishlisted=wish)
    filt.compiler.now_year = int(g["now_year"])
    specs = json.loads(str(g["scenario_specs"]))
    off = g["valid_off"]
    for j, name in enumerate(g["scenario_names"]):
        if "kwargs" in specs[j]:
            cons = filt.create_constraint_set(**specs[j]["kwargs"])
            assert [[c.constraint_type.value, c.value] for c in cons] == specs[j]["pairs"], name
        else:
            cons = [bb.HardConstraint(bb.ConstraintType(t), v) for t, v in specs[j]["pairs"]]
        res = filt.apply_constraints(cons)
        got = np.array([filt.index_of[s] for s in res.valid_set_nums], dtype=np.int64)
        assert np.array_equal(got, g["valid_flat"][off[j]:off[j + 1]]), f"scenario {name}"
        assert res.applied_constraints == cons and res.performance_stats["constraint_count"] == len(cons)


def test_violation_messages_and_candidates(golden_dir):
    g, themes, owned, wish = _golden_masks(golden_dir)
    n = g["num_parts"].shape[0]
    eng = FakeEngine(n, 8)
    eng.load_catalog(emb=np.ones((n, 8), np.float32), num_parts=g["num_parts"], year=g["year"], theme_id=g["theme_id"])
    sn = synth.set_nums(n)
    filt = bb.HardConstraintFilter(eng, sn, themes)
    cons = [bb.HardConstraint(bb.ConstraintType.PIECES_MIN, 5000, description="Minimum 5,000 pieces"),
            bb.HardConstraint(bb.ConstraintType.PRICE_MAX, 0.5, description="Maximum $0.50")]
    res = filt.apply_constraints(cons)
    assert res.valid_set_nums == []
    assert len(res.violations) == 2 and "eliminated" in res.violations[0].message
    assert res.violations[1].suggested_alternatives[0] == "Try increasing budget to $0.65"
    cand = sn[:500]
    res = filt.apply_constraints([bb.HardConstraint(bb.ConstraintType.PIECES_MAX, 500)], candidate_set_nums=cand)
    assert set(res.valid_set_nums) <= set(cand) and res.performance_stats["input_sets"] == 500
    assert filt.get_performance_report()["operation_count"] == 2
    assert filt.create_constraint_set(exclude_owned=True) == []          # silently dropped without user_id
    assert bb.create_age_appropriate_constraints(8)[0].value == 6
    assert [c.value for c in bb.create_size_constraints("Medium")] == [201, 800]
    assert bb.create_size_constraints("gigantic") == []
    assert [c.constraint_type for c in bb.create_budget_constraints(100.0, 20.0)] == [bb.ConstraintType.PRICE_MAX, bb.ConstraintType.PRICE_MIN]


def test_descriptor_matches_oracle_spec():
    themes = synth.themes_table()
    comp = bb.ConstraintCompiler(themes, now_year=2026)
    num_parts, year, theme = synth.metadata(5000, 3, themes, null_theme_frac=0.05)
    for name, pairs in synth.SCENARIO_LISTS:
        cons = [bb.HardConstraint(bb.ConstraintType(t), v) for t, v in pairs]
        cc = comp.compile(cons)
        d = cc.desc.copy()
        sets = []
        if cc.theme_bitmap is not None:
            d["theme_set"] = 0
            sets = [cc.theme_bitmap]
        got = eval_descriptor(d, num_parts, year, theme, sets, [], [])
        want = bo.constraint_mask(bo.compile_constraints(pairs, themes, 2026), num_parts, year, theme)
        assert np.array_equal(got, want), name


def test_bitmap_roundtrip():
    rng = np.random.default_rng(0)
    for n in (1, 31, 32, 33, 1000):
        m = rng.random(n) < 0.3
        w = bb.ScoringEngine.pack_bitmap(m)
        assert w.dtype == np.uint32 and w.shape[0] == (n + 31) // 32
        assert np.array_equal(bb.ScoringEngine.unpack_bitmap(w, n), m)
        assert all(((w[i // 32] >> (i % 32)) & 1) == int(m[i]) for i in range(n))


# ------------------------------------------------------------------------------------------ adapters (host logic)
def _fake_hybrid(golden_dir, monkeypatch):
    import brickbrain_rec_engine_b200.recommenders as rec
    monkeypatch.setattr(rec, "ScoringEngine", type("E", (FakeEngine,), {"pack_bitmap": staticmethod(bb.ScoringEngine.pack_bitmap),
                                                                       "unpack_bitmap": staticmethod(bb.ScoringEngine.unpack_bitmap)}))
    gc = np.load(os.path.join(golden_dir, "content_fp32_golden.npz"))
    gf = np.load(os.path.join(golden_dir, "cf_golden.npz"))
    gh = np.load(os.path.join(golden_dir, "hybrid_golden.npz"))
    feat = gc["q8"].astype(np.float32) / gc["scale"]
    n = feat.shape[0]
    themes = {int(k): v for k, v in json.loads(str(gh["themes_json"])).items()}
    owned = {int(k): np.array(v, dtype=np.int32) for k, v in json.loads(str(gh["owned_json"])).items()}
    wish = {int(k): np.array(v, dtype=np.int32) for k, v in json.loads(str(gh["wish_json"])).items()}
    cat = bb.CatalogTable(synth.set_nums(n), [f"Set {i}" for i in range(n)], gh["year"], gh["num_parts"], gh["theme_id"],
                          [themes.get(int(t), "Unknown") for t in gh["theme_id"]])
    user_ids = [int(u) for u in gf["user_ids"]]
    rated = {u: np.nonzero(gf["ratings"][i] > 0)[0].astype(np.int32) for i, u in enumerate(user_ids)}
    hyb = bb.HybridRecommender.from_arrays(cat, feat, themes, user_ids=user_ids, user_factors=gf["user_factors"],
                                           item_factors=gf["item_factors"], rated=rated, owned=owned, wishlisted=wish)
    hyb.constraint_filter.compiler.now_year = int(gh["now_year"])
    return hyb, cat, gc, gf, gh


def test_hybrid_flow_matches_reference_golden(golden_dir, monkeypatch):
    hyb, cat, gc, gf, gh = _fake_hybrid(golden_dir, monkeypatch)
    k = int(gh["k"])
    for case in json.loads(str(gh["cases_json"])):
        hyb.content_weight, hyb.collaborative_weight = case["weights"]
        cons = hyb.constraint_filter.create_constraint_set(**case["kwargs"]) if case["kwargs"] else None
        liked = cat.set_nums[case["liked"]] if case["liked"] is not None else None
        recs, cres = hyb.get_recommendations(user_id=case["user_id"], liked_set=liked, top_k=k, constraints=cons)
        want = [i for i in case["ids"] if i >= 0]
        got = [cat.index_of[r.set_num] for r in recs]
        assert got == want, case["kwargs"]
        np.testing.assert_allclose([r.score for r in recs], [s for s in case["scores"] if s is not None], rtol=1e-5)
        if case["n_valid"] == 0:
            assert recs == [] and cres.valid_set_nums == []
        for r in recs:
            assert isinstance(r, bb.RecommendationResult) and r.name == f"Set {cat.index_of[r.set_num]}"
    # request object path + reference error conventions
    req = bb.RecommendationRequest(user_id=1, liked_set=cat.set_nums[int(gc["targets"][1])], top_k=5, pieces_max=1000)
    recs, cres = hyb.get_recommendations_from_request(req)
    assert len(recs) == 5 and cres.performance_stats["constraint_count"] == 1
    assert any(r.startswith("Content: ") or r.startswith("Community: ") for r in recs[0].reasons)
    assert hyb.content_recommender.get_similar_sets("nope", 5) == []
    cold = hyb.collaborative_recommender.get_recommendations(987654, 5)          # unknown user -> cold start, no raise
    assert all(isinstance(r, bb.RecommendationResult) for r in cold)
    hyb.set_weights(0.7, 0.5)
    assert hyb.content_weight == 0.7 / 1.2 and hyb.collaborative_weight == 0.5 / 1.2


def test_recommend_batch_builds_one_fused_call(golden_dir, monkeypatch):
    hyb, cat, gc, gf, gh = _fake_hybrid(golden_dir, monkeypatch)
    users = [int(u) for u in gf["users"][:4]]
    liked = [cat.set_nums[int(t)] for t in gc["targets"][:4]]
    cons = [hyb.constraint_filter.create_constraint_set(pieces_max=1000), None,
            hyb.constraint_filter.create_constraint_set(required_themes=["City"]), None]
    hyb.engine.calls.clear()
    scores, ids = hyb.recommend_batch(users, liked, 10, constraints=cons, precision="fp32")
    assert hyb.engine.calls == ["search"] and ids.shape == (4, 10)
    rated = {int(u): set(np.nonzero(gf["ratings"][i] > 0)[0]) for i, u in enumerate(gf["user_ids"])}
    for j, u in enumerate(users):
        assert not (set(ids[j]) & rated[u]) and cat.index_of[liked[j]] not in set(ids[j])
    assert (gh["num_parts"][ids[0]] <= 1000).all()


# ------------------------------------------------------------------------------------------ vector store
def test_read_faiss_flat(tmp_path, golden_dir):
    emb = np.load(os.path.join(golden_dir, "faiss_fixture_golden.npz"))["emb"]
    n, d = emb.shape
    raw = b"IxF2" + struct.pack("<i", d) + struct.pack("<q", n) + struct.pack("<qq", 1 << 20, 1 << 20) + b"\x01" + \
        struct.pack("<i", 1) + struct.pack("<q", n * d) + emb.astype("<f4").tobytes()
    p = tmp_path / "index.faiss"
    p.write_bytes(raw)
    rows, metric = bb.read_faiss_flat(str(p))
    assert metric == "l2" and np.array_equal(rows, emb)
    ref = "/root/reference/test_embeddings/index.faiss"
    if os.path.exists(ref):
        rows2, m2 = bb.read_faiss_flat(ref)
        assert m2 == "l2" and np.array_equal(rows2, emb)
    (tmp_path / "bad").write_bytes(b"IVFx" + raw[4:])
    with pytest.raises(ValueError):
        bb.read_faiss_flat(str(tmp_path / "bad"))


# ------------------------------------------------------------------------------------------ sharding (gloo, world_size 2)
def test_shard_bounds():
    for n, w in ((10, 3), (25216, 8), (10_000_000, 8), (5, 8)):
        b = [bb.shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "oracle")); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import brick_oracle as bo, synth
import brickbrain_rec_engine_b200 as bb
from fake_engine import FakeEngine
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
n, d, f, b, k = 3001, 32, 8, 9, 7
emb = synth.embeddings(n, d, 1, "aniso"); itf = synth.cf_factors(n, f, 2)
lo, hi = bb.shard_bounds(n, world, rank)
eng = FakeEngine(hi - lo, d, f, id_offset=lo)
eng.load_catalog(emb=emb[lo:hi], item_factors=itf[lo:hi])
q = synth.embeddings(b, d, 3, "aniso"); u = synth.cf_factors(b, f, 4)
s, i = bb.ShardedSearch(eng).search(k, q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6)
ref = bo.fused_hybrid_scores(q, u, bo.l2_normalize_rows(emb), itf, np.full(b, .4, np.float32), np.full(b, .6, np.float32))
rid, rs = bo.fused_hybrid_topk(ref, k)
assert np.array_equal(np.asarray(i), rid), (rank, i, rid)
np.testing.assert_allclose(np.asarray(s), rs, rtol=1e-6)
# batch divisible by the world size: every rank uploads 1/world of the host batch and the slices are all-gathered
b2 = 8
sid = (np.arange(b2, dtype=np.int32) * 37) % n
s2, i2 = bb.ShardedSearch(eng).search(k, q_emb=q[:b2], q_cf=u[:b2], w_content=0.4, w_cf=0.6, self_id=sid)
rid2, rs2 = bo.fused_hybrid_topk(ref[:b2], k, self_ids=sid)
assert np.array_equal(np.asarray(i2), rid2), (rank, i2, rid2)
np.testing.assert_allclose(np.asarray(s2), rs2, rtol=1e-6)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_search_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out[-3000:]


# ------------------------------------------------------------------------------------------ bench CPU arm
def test_bench_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--n-items", "20000", "--cpu-sample-items", "20000", "--cpu-sample-queries", "16", "--batch", "16"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_cf_training_host_logic():
    """cf_training: the LU normalizer (P @ L without a dense permutation) equals scipy's permute_l form, the CSR pair
    is built as the kernel expects it, and training refuses to run without the CUDA engine (no CPU fallback)."""
    import torch
    from scipy import linalg
    from brickbrain_rec_engine_b200 import cf_training
    rng = np.random.default_rng(0)
    for shape in ((300, 7), (64, 64), (2000, 60)):
        x = rng.standard_normal(shape)
        ref, _ = linalg.lu(x, permute_l=True)
        np.testing.assert_allclose(cf_training._lu_pl(torch.from_numpy(x)).numpy(), ref, rtol=0, atol=1e-12)
    dense = ((rng.random((40, 25)) < 0.2) * rng.integers(1, 6, (40, 25))).astype(np.float32)
    pair = cf_training.CsrPair(dense, "cpu")
    assert pair.shape == (40, 25) and pair.nnz == int((dense != 0).sum())
    ip, ix, v = (t.numpy() for t in pair.a)
    assert ip.dtype == np.int64 and ix.dtype == np.int32 and v.dtype == np.float32 and ip[-1] == pair.nnz
    back = np.zeros_like(dense)
    for r in range(40):
        back[r, ix[ip[r]:ip[r + 1]]] = v[ip[r]:ip[r + 1]]
    assert np.array_equal(back, dense)
    ipt, ixt, vt = (t.numpy() for t in pair.at)
    assert ipt.shape[0] == 26 and ipt[-1] == pair.nnz and (np.diff(ixt[ipt[3]:ipt[4]]) > 0).all()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            cf_training.truncated_svd(None, dense, 5)
