"""Item-item CF (config 3) probe: one workload, one child process per BB_II_CHUNK value (raters per work unit); set
BRICKSCORE_LIB to time another build of the library on the same box.
``python tools/ii_probe.py [--users 1000000] [--chunks 8192,16384]``"""
import argparse
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def workload(n_items, n_users):
    rng = np.random.default_rng(3)
    w = 1.0 / np.arange(1, n_items + 1) ** 0.8
    w /= w.sum()
    perm = rng.permutation(n_items)
    counts = np.clip(rng.poisson(20, size=n_users), 1, None)
    draws = perm[rng.choice(n_items, size=int(counts.sum()), p=w)].astype(np.int64)
    rows = np.repeat(np.arange(n_users, dtype=np.int64), counts)
    key = np.unique(rows * n_items + draws)
    rows, cols = key // n_items, (key % n_items).astype(np.int32)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n_users), out=indptr[1:])
    vals = rng.integers(1, 6, size=cols.shape[0]).astype(np.float32)
    return indptr, cols, vals


def child(path, n_items, k, steps):
    import torch
    from brickbrain_rec_engine_b200 import ScoringEngine
    z = np.load(path)
    eng = ScoringEngine(1, 1, 0, device=0)
    d = [torch.from_numpy(z[x]).cuda() for x in ("indptr", "cols", "vals")]
    import time
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.3:                    # the SM clocks start from idle: warm up by time
        out = eng.item_item_topk(*d, n_items, k, device_out=True)
        torch.cuda.synchronize()
    t = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.item_item_topk(*d, n_items, k, device_out=True)
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    s, i = out
    chk = float(s.double().sum().item()), int(i.long().sum().item())
    print(f"chunk={os.environ.get('BB_II_CHUNK', 'default')} lib={os.environ.get('BRICKSCORE_LIB', 'in-tree')} "
          f"ms={np.mean(t):.3f} min={np.min(t):.3f} checksum={chk[0]:.6f}/{chk[1]}", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--n-items", type=int, default=25_216)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--chunks", default="16384")
    ap.add_argument("--child", default=None)
    a = ap.parse_args()
    if a.child:
        child(a.child, a.n_items, a.k, a.steps)
        sys.exit(0)
    t0 = time.time()
    indptr, cols, vals = workload(a.n_items, a.users)
    path = "/tmp/ii_probe.npz"
    np.savez(path, indptr=indptr, cols=cols, vals=vals)
    print(f"workload: {cols.shape[0]} ratings, generated in {time.time() - t0:.1f} s", flush=True)
    for c in a.chunks.split(","):
        env = dict(os.environ, BB_II_CHUNK=c)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", path, "--n-items", str(a.n_items), "--k", str(a.k),
                        "--steps", str(a.steps)], env=env, check=False)
