"""Collaborative-filtering factor training on the GPU: the reference's ``train_svd_model``
(recommendation_system.py:368-409) = ``TruncatedSVD(n_components, random_state=42).fit_transform`` on the
user x item rating matrix, i.e. scikit-learn's randomized truncated SVD:

    Q  = RandomState(42).normal(size=(cols, n_components + 10))
    5x:  Q = lu(A @ Q).PL ;  Q = lu(A.T @ Q).PL            (power iterations, LU-normalized)
    Q  = qr(A @ Q).Q ;  B = Q.T @ A ;  Uhat, s, Vt = svd(B)  (A transposed first when it has more columns than rows)
    sign convention svd_flip(u_based_decision=False);  item_factors = Vt.T ;  user_factors = X @ item_factors

The sparse products (the only work that scales with the number of ratings) run in the hand-written CSR kernel
(``bb_csr_spmm_f64``, csrc/bb_spmm.cu) in fp64 like the reference; the thin (rows x 60) LU / QR factorizations
and the 60-wide SVD are cuSOLVER calls through torch.  The random test matrix is drawn on the host from the same
NumPy stream as scikit-learn, so the factors reproduce the reference's to rounding (tests/golden/cf_golden.npz).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class CsrPair:
    """A rating matrix resident on the device as CSR(A) and CSR(A.T) (both needed by the power iterations)."""

    def __init__(self, ratings, device: str):
        import scipy.sparse as sp
        a = sp.csr_matrix(ratings, dtype=np.float32)
        a.eliminate_zeros()
        a.sort_indices()
        at = a.T.tocsr()
        at.sort_indices()
        self.shape = a.shape
        self.nnz = int(a.nnz)
        self.a = self._to_device(a, device)
        self.at = self._to_device(at, device)

    @staticmethod
    def _to_device(m, device):
        return (torch.from_numpy(m.indptr.astype(np.int64)).to(device), torch.from_numpy(m.indices.astype(np.int32)).to(device),
                torch.from_numpy(m.data.astype(np.float32)).to(device))


def _lu_pl(x):
    """scipy.linalg.lu(x, permute_l=True)[0]: the row-permuted unit-lower factor P @ L of a tall matrix, without
    materializing the rows x rows permutation (getrf pivots are n sequential row swaps: at most 2n rows move)."""
    rows, n = x.shape
    lu, piv = torch.linalg.lu_factor(x)
    l = torch.tril(lu, diagonal=-1)
    l[:n] += torch.eye(n, dtype=x.dtype, device=x.device)
    where = {}                                         # position -> original row currently held there (sparse view of perm)
    for i, p in enumerate((piv.cpu().numpy() - 1).tolist()):
        a, b = where.get(i, i), where.get(p, p)
        where[i], where[p] = b, a
    src = [k for k, v in where.items() if k != v]
    if src:                                            # A[perm] = L U  =>  (P L)[perm[i]] = L[i]
        dst = torch.tensor([where[k] for k in src], device=x.device)
        src_t = torch.tensor(src, device=x.device)
        rows_l = l[src_t].clone()
        l[dst] = rows_l
    return l


def truncated_svd(engine, ratings, n_components: int = 50, n_iter: int = 5, n_oversamples: int = 10,
                  random_state: int = 42) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``(user_factors [U, F], item_factors [N, F], singular_values [F])`` of a user x item rating matrix
    (dense ndarray / DataFrame values or scipy sparse; zeros = missing), as the reference's ``train_svd_model``
    stores them (recommendation_system.py:402-407)."""
    if torch is None or not torch.cuda.is_available():
        raise RuntimeError("cf_training.truncated_svd needs the CUDA engine; there is no CPU fallback")
    dev = f"cuda:{engine.device}"
    m = CsrPair(ratings, dev)
    n_users, n_items = m.shape
    if n_components > n_items:
        raise ValueError(f"n_components({n_components}) must be <= n_features({n_items}).")
    n_random = n_components + n_oversamples
    if n_random > 128:
        raise ValueError("n_components + n_oversamples must be <= 128")
    transpose = n_users < n_items                      # sklearn: work on the matrix with fewer columns
    w, wt = (m.at, m.a) if transpose else (m.a, m.at)
    rows, cols = (n_items, n_users) if transpose else (n_users, n_items)
    q = torch.from_numpy(np.random.RandomState(random_state).normal(size=(cols, n_random))).to(dev)

    def mul(csr, n_cols, x):
        return engine.csr_spmm(csr[0], csr[1], csr[2], n_cols, x.contiguous())

    with torch.cuda.device(engine.device):
        for _ in range(n_iter):
            q = _lu_pl(mul(w, cols, q))                # [rows, n_random]
            q = _lu_pl(mul(wt, rows, q))               # [cols, n_random]
        q, _ = torch.linalg.qr(mul(w, cols, q), mode="reduced")
        b = mul(wt, rows, q).T.contiguous()            # Q.T @ W  [n_random, cols]
        uhat, s, vt = torch.linalg.svd(b, full_matrices=False)
        u = q @ uhat
        if transpose:
            vt_f, s_f = u[:, :n_components].T, s[:n_components]
        else:
            vt_f, s_f = vt[:n_components], s[:n_components]
        # svd_flip(u_based_decision=False): the largest-magnitude entry of every component is positive
        idx = vt_f.abs().argmax(dim=1)
        sign = torch.sign(vt_f[torch.arange(vt_f.shape[0], device=dev), idx])
        sign = torch.where(sign == 0, torch.ones_like(sign), sign)
        item_factors = (vt_f * sign[:, None]).T.contiguous()          # [N, F] = components_.T
        user_factors = mul(m.a, n_items, item_factors)                # X @ components_.T  (sklearn >= 1.x transform)
        return user_factors.cpu().numpy(), item_factors.cpu().numpy(), s_f.cpu().numpy()
