"""How much can the union thresholds buy?  G shards of one catalog as G handles on ONE GPU (bb_comm_link_local), searched
one after the other: shard g scans with the FINAL claims of shards 0..g-1 in place, i.e. the last shard sees (almost) the
global threshold from its first tile on.  Prints per shard: scan ms, appends, per-round cycles (profiling build).

    python tools/union_probe.py [n_items_total] [G] [batch]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import brickbrain_rec_engine_b200 as bb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    b = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
    k, d, f, dev = 100, 384, 64, "cuda:0"
    descs, theme_bitmaps, themes, _ = bench.scenario_descriptors(b)
    distinct, cidx = bb.ScoringEngine.dedupe_constraints(descs)
    shards = []
    for r in range(G):
        lo, hi = bb.shard_bounds(n, G, r)
        e = bb.ScoringEngine(hi - lo, d, f, id_offset=lo, max_batch=b)
        e.set_theme_sets(theme_bitmaps)
        mu, _ = bench.load_synthetic_catalog(torch, e, lo, hi, d, f, dev, themes)
        shards.append(e)
    for r, e in enumerate(shards):
        e.comm_link_local(r, shards)
    g = torch.Generator(device="cpu").manual_seed(100)
    q = torch.nn.functional.normalize(torch.randn(b, d, generator=g), dim=1)
    q = (torch.nn.functional.normalize(0.8 * mu.cpu()[None, :] + 0.6 * q, dim=1) * 3.0).to(dev)
    u = (torch.randn(b, f, generator=g) / (f ** 0.5)).to(dev)
    cons = torch.from_numpy(distinct.view(np.uint8).reshape(distinct.shape[0], -1).copy()).to(dev)
    kw = dict(q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, constraint_index=torch.from_numpy(cidx).to(dev),
              precision="bf16_rescore")
    for rep in range(2):
        mine = []
        for r, e in enumerate(shards):
            e.set_profiling(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mine.append(e.sharded_scatter(k, **kw))
            e1.record()
            torch.cuda.synchronize()
            i = e.info()
            if rep == 1:
                print(f"shard {r}: scatter {e0.elapsed_time(e1):.3f} ms, scan {i['scan_ms'] / max(1, i['scan_launches']):.3f} ms units={i['last_units']} appends={i['last_appends']} "
                      f"slow_chunks={i['last_slow_chunks']} compactions={i['last_compactions']} "
                      f"mma_wait_epi={i['mma_wait_epilogue_cycles'] / max(1, i['mma_total_cycles']):.3f} "
                      f"rounds_kcyc={[round(x / 74 / 1e3) for x in i['round_cycles'] if x]}", flush=True)
            e.set_profiling(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for e, m in zip(shards, mine):
            e.sharded_collect(k, m, device=True)
        e1.record()
        torch.cuda.synchronize()
        if rep == 1:
            print(f"collect (all {G} shards): {e0.elapsed_time(e1):.3f} ms")


if __name__ == "__main__":
    main()
