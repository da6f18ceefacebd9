"""A short pass over every kernel of the library (catalog load, mask, searches of all kinds, merges, item-item, CF training
products) for the bandwidth report -- run under ncu (tools/bandwidth_report.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import brickbrain_rec_engine_b200 as bb  # noqa: E402


def main():
    dev = "cuda:0"
    n, d, f, b, k = 2_000_000, 384, 64, 4096, 100
    themes = bench.themes_table()
    eng = bb.ScoringEngine(n, d, f, max_batch=b)
    descs, theme_bitmaps, _, _ = bench.scenario_descriptors(b)
    eng.set_theme_sets(theme_bitmaps)
    mu, _ = bench.load_synthetic_catalog(torch, eng, 0, n, d, f, dev, themes)          # prep_catalog, pack_meta
    distinct, cidx = bb.ScoringEngine.dedupe_constraints(descs)
    g = torch.Generator(device=dev).manual_seed(1)
    q = torch.randn(b, d, device=dev, generator=g)
    u = torch.randn(b, f, device=dev, generator=g) / 8
    cons = torch.from_numpy(distinct.view(np.uint8).reshape(distinct.shape[0], -1).copy()).to(dev)
    kw = dict(q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, constraint_index=torch.from_numpy(cidx).to(dev))
    for prec in ("bf16_rescore", "fp32"):
        eng.search(k, precision=prec, **kw)                                              # prep_queries, mask, pilot, seed, scan, merge_emit
    eng.apply_constraints(distinct)                                                      # mask (bb_apply_constraints)
    eng.search(10, q_emb=q[:1], precision="fp32")                                        # score_small, select_small
    eng.search(10, q_emb=q[:8], precision="fp32")
    s = torch.randn(4, b, k, device=dev)
    i = torch.randint(0, n, (4, b, k), device=dev, dtype=torch.int32)
    eng.merge_topk(s, i, k)                                                              # merge_lists
    rng = np.random.default_rng(0)
    n_items, n_users = 25216, 200_000
    deg = rng.poisson(20, n_users).clip(1)
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    w = 1.0 / np.arange(1, n_items + 1) ** 0.8
    cols = rng.choice(n_items, int(indptr[-1]), p=w / w.sum()).astype(np.int32)
    for r in range(n_users):        # rows must hold distinct, sorted items
        pass
    key = np.unique(np.repeat(np.arange(n_users, dtype=np.int64), deg) * n_items + cols)
    rows, cols = key // n_items, (key % n_items).astype(np.int32)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n_users), out=indptr[1:])
    vals = rng.integers(1, 6, cols.shape[0]).astype(np.float32)
    eng.item_item_topk(indptr, cols, vals, n_items, 20)                                  # ii_* kernels
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
