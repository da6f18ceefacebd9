"""Timings of the other BASELINE.json configs on one B200 (parity for these shapes lives in tests/).

    python tools/config_probe.py            # configs 1, 2, 3, 5 (per-GPU share)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth  # noqa: E402
import torch  # noqa: E402
import brickbrain_rec_engine_b200 as bb  # noqa: E402

dev = "cuda:0"
out = {}


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---------------- configs 1 and 2: 25,216 x 384 catalog
n, d, f = 25216, 384, 64
themes = synth.themes_table()
emb = synth.embeddings(n, d, 1, "aniso")
itf = synth.cf_factors(n, f, 2)
parts, year, theme = synth.metadata(n, 3, themes)
eng = bb.ScoringEngine(n, d, f)
eng.load_catalog(emb=emb, item_factors=itf, num_parts=parts, year=year, theme_id=theme)
filt = bb.HardConstraintFilter(eng, synth.set_nums(n), themes)
desc = filt.descriptor(filt.create_constraint_set(pieces_max=1000, year_min=2010))
# config 1: single query, cosine top-10 with filter, host buffers in and out (the reference's own CPU-runnable case)
q1 = synth.embeddings(1, d, 5, "aniso")
t0 = time.perf_counter()
for _ in range(200):
    eng.search(10, q_emb=q1, constraints=desc, precision="fp32")
out["config1_single_query_fp32_host_ms"] = (time.perf_counter() - t0) / 200 * 1e3
t0 = time.perf_counter()
for _ in range(200):
    eng.search(10, q_emb=q1, constraints=desc, precision="bf16_rescore")
out["config1_single_query_bf16rescore_host_ms"] = (time.perf_counter() - t0) / 200 * 1e3
# config 2: batch 4096, content + CF + mask, top-50, device resident
b = 4096
q = torch.from_numpy(synth.embeddings(b, d, 6, "aniso")).to(dev)
u = torch.from_numpy(synth.cf_factors(b, f, 7)).to(dev)
cons = torch.from_numpy(np.atleast_1d(desc).view(np.uint8).reshape(1, -1).copy()).to(dev)
for prec in ("bf16_rescore", "bf16", "fp32"):
    ms = timed(lambda: eng.search(50, q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, precision=prec))
    out[f"config2_batch4096_{prec}_ms"] = ms
    out[f"config2_batch4096_{prec}_qps"] = b / ms * 1e3
    out[f"config2_batch4096_{prec}_tflops"] = 2.0 * b * n * (d + f) / (ms * 1e-3) / 1e12
eng.close()

# ---------------- config 3: item-item CF, 1M users x 25,216 items, ~20 ratings per user, top-20
n_users = 1_000_000
t0 = time.perf_counter()
indptr, indices, data = synth.zipf_ratings_csr(n_users, n, 20, seed=11)
out["config3_gen_seconds"] = time.perf_counter() - t0
ip, ix, dv = (torch.from_numpy(a).to(dev) for a in (indptr, indices, data))
eng3 = bb.ScoringEngine(0, 8, 0)
deg = np.diff(indptr)
out["config3_nnz"] = int(indptr[-1])
out["config3_sum_nnz_u_squared"] = int((deg.astype(np.int64) ** 2).sum())
torch.cuda.synchronize()
t0 = time.perf_counter()
s, i = eng3.item_item_topk(ip, ix, dv, n, 20)
torch.cuda.synchronize()
out["config3_first_call_ms"] = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
s, i = eng3.item_item_topk(ip, ix, dv, n, 20)
torch.cuda.synchronize()
out["config3_item_item_ms"] = (time.perf_counter() - t0) * 1e3
out["config3_gmacs_per_s"] = out["config3_sum_nnz_u_squared"] / (out["config3_item_item_ms"] * 1e-3) / 1e9
eng3.close()

# ---------------- config 5 (per-GPU share of the 8-GPU job): 1M items x (384+64), 125k users, top-100
n5, users = 1_000_000, 125_000
eng5 = bb.ScoringEngine(n5, d, 64, max_batch=8192)
g = torch.Generator(device=dev).manual_seed(3)
x = torch.nn.functional.normalize(torch.randn(n5, d, device=dev, generator=g) + 0.8, dim=1)
v = torch.randn(n5, 64, device=dev, generator=g) / 8
eng5.load_catalog(emb=x, item_factors=v)
del x, v
uf = torch.randn(users, 64, device=dev, generator=g) / 8
liked = torch.randint(0, n5, (users,), device=dev, generator=g, dtype=torch.int32)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = bb.bulk_recommend(eng5, uf, liked, 100, batch=8192, precision="bf16_rescore")
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out["config5_share_users"] = users
out["config5_share_seconds"] = dt
out["config5_users_per_s_per_gpu"] = users / dt
out["config5_tflops"] = 2.0 * users * n5 * 448 / dt / 1e12
print(json.dumps(out, indent=1))
