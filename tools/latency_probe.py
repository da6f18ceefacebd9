"""Single-query / small-batch latency through the host-buffer API (BASELINE config 1 shape and a 1M-row catalog):
per-call wall time, and the scan kernel's share (CUDA events, bb_set_profiling)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth  # noqa: E402
import brickbrain_rec_engine_b200 as bb  # noqa: E402


def main():
    d = 384
    for n in (25216, 1000000):
        emb = synth.embeddings(n, d, 3, "aniso")
        eng = bb.ScoringEngine(n, d, 0, max_batch=64)
        eng.load_catalog(emb=emb)
        for b in (1, 8, 64):
            rows = np.arange(b, dtype=np.int32) * 7
            for prec in ("bf16_rescore", "fp32"):
                for _ in range(5):
                    eng.search(10, q_item=rows, self_id=rows, precision=prec)
                eng.set_profiling(True)
                t0 = time.perf_counter()
                reps = 50
                for _ in range(reps):
                    eng.search(10, q_item=rows, self_id=rows, precision=prec)
                dt = (time.perf_counter() - t0) / reps * 1e3
                i = eng.info()
                eng.set_profiling(False)
                print(f"n={n} b={b} {prec}: {dt:.3f} ms/call, scan kernel {i['scan_ms'] / max(1, i['scan_launches']):.3f} ms, "
                      f"units={i['last_units']} appends={i['last_appends']} compactions={i['last_compactions']}", flush=True)
        if n == 25216:
            # BASELINE config 1: one query, cosine top-10 WITH a hard-constraint filter -- the filter travels with the call
            mask = np.random.default_rng(1).random(n) < 0.3
            bm = eng.pack_bitmap(mask).reshape(1, -1)
            cons = bb.no_constraint()
            cons["allow_bitmap"] = 0
            rows = np.array([77], dtype=np.int32)
            for prec in ("bf16_rescore", "fp32"):
                for _ in range(5):
                    eng.search(10, q_item=rows, self_id=rows, constraints=cons, allow_bitmaps=bm, precision=prec)
                t0 = time.perf_counter()
                for _ in range(200):
                    eng.search(10, q_item=rows, self_id=rows, constraints=cons, allow_bitmaps=bm, precision=prec)
                print(f"config1 engine call (host bitmap filter, host in/out) {prec}: {(time.perf_counter() - t0) / 200 * 1e3:.3f} ms/call", flush=True)
            themes = synth.themes_table()
            parts, year, theme = synth.metadata(n, 8, themes)
            cat = bb.CatalogTable(synth.set_nums(n), [f"set {i}" for i in range(n)], year, parts, theme, [themes.get(int(t), "Unknown") for t in theme])
            hyb = bb.HybridRecommender.from_arrays(cat, emb, themes)
            valid = [cat.set_nums[j] for j in np.nonzero(mask)[0]]
            rec = hyb.content_recommender
            for _ in range(3):
                rec.get_similar_sets(cat.set_nums[77], 10, valid_set_filter=valid)
            t0 = time.perf_counter()
            for _ in range(50):
                rec.get_similar_sets(cat.set_nums[77], 10, valid_set_filter=valid)
            print(f"config1 get_similar_sets(set_num, 10, valid_set_filter=<{len(valid)} set_nums>) incl. hydration: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms/call", flush=True)
            cons_list = hyb.constraint_filter.create_constraint_set(pieces_max=1000, year_min=2010)
            for _ in range(3):
                hyb.get_recommendations(liked_set=cat.set_nums[77], top_k=10, constraints=cons_list)
            t0 = time.perf_counter()
            for _ in range(50):
                hyb.get_recommendations(liked_set=cat.set_nums[77], top_k=10, constraints=cons_list)
            print(f"config1 HybridRecommender.get_recommendations(liked_set, constraints) incl. mask kernel + violation analysis: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms/call", flush=True)
            hyb.engine.close()
        eng.close()


if __name__ == "__main__":
    main()
