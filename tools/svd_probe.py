"""Times CF factor training (cf_training.truncated_svd) and its CSR x dense kernel at BASELINE config-3 scale:
1M users x 25,216 items, ~2e7 ratings (Zipf item popularity), 50 components (L = 60)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth  # noqa: E402
import brickbrain_rec_engine_b200 as bb  # noqa: E402
from brickbrain_rec_engine_b200 import cf_training  # noqa: E402


def main():
    import scipy.sparse as sp
    n_users, n_items = int(os.environ.get("SVD_USERS", "1000000")), 25216
    indptr, indices, data = synth.zipf_ratings_csr(n_users, n_items, 20, seed=9)
    a = sp.csr_matrix((data.astype(np.float32), indices, indptr), shape=(n_users, n_items))
    eng = bb.ScoringEngine(64, 64, 0)
    m = cf_training.CsrPair(a, "cuda:0")
    nnz, l = m.nnz, 60
    out = {"n_users": n_users, "n_items": n_items, "nnz": nnz, "L": l}
    for name, csr, n_rows, n_cols in (("A@Q", m.a, n_users, n_items), ("A.T@Q", m.at, n_items, n_users)):
        x = torch.randn((n_cols, l), dtype=torch.float64, device="cuda")
        y = eng.csr_spmm(csr[0], csr[1], csr[2], n_cols, x)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(5):
            eng.csr_spmm(csr[0], csr[1], csr[2], n_cols, x, out=y)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 5
        alg = nnz * (8 + l * 8) + n_rows * (8 + l * 8)
        out[name] = {"ms": round(ms, 3), "algorithmic_GB": round(alg / 1e9, 3), "GB_per_s": round(alg / ms / 1e6, 1),
                     "GFLOP_per_s": round(2 * nnz * l / ms / 1e6, 1)}
    t0 = time.perf_counter()
    uf, itf, s = cf_training.truncated_svd(eng, a, 50)
    torch.cuda.synchronize()
    out["train_s_incl_host_csr_build"] = round(time.perf_counter() - t0, 2)
    out["singular_values_head"] = [round(float(v), 3) for v in s[:4]]
    if n_users <= 200000:
        from sklearn.decomposition import TruncatedSVD
        t0 = time.perf_counter()
        ref = TruncatedSVD(n_components=50, random_state=42).fit(a.astype(np.float64))
        out["sklearn_fit_s"] = round(time.perf_counter() - t0, 2)
        out["max_abs_diff_item_factors_vs_sklearn"] = float(np.abs(ref.components_.T - itf).max())
    print(out)


if __name__ == "__main__":
    main()
