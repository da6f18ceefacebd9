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
        eng.close()


if __name__ == "__main__":
    main()
