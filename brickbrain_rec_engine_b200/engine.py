"""``ScoringEngine`` -- thin object wrapper over the C ABI (one handle = one catalog shard on one GPU).

Arrays may be NumPy (host) or torch tensors (host or CUDA); CUDA tensors are
passed by device pointer and the call is enqueued on torch's current stream
without host synchronization.  PyTorch is used only for device memory and
streams -- all arithmetic happens inside libbrickscore.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib as L

try:  # plumbing only
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _np_dtype_to_torch(dt):
    return {np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32, np.dtype(np.int16): torch.int16,
            np.dtype(np.int64): torch.int64, np.dtype(np.float64): torch.float64, np.dtype(np.uint32): torch.int32,
            np.dtype(np.uint8): torch.uint8}[np.dtype(dt)]


class _Args:
    """Collects raw pointers and keeps the backing arrays alive for the call."""

    def __init__(self):
        self.keep = []
        self.any_cuda = False

    def ptr(self, x, dtype, shape=None):
        if x is None:
            return None
        if _is_torch(x):
            t = x
            want = _np_dtype_to_torch(dtype)
            if t.dtype != want:
                t = t.to(want)
            t = t.contiguous()
            if shape is not None and tuple(t.shape) != tuple(shape):
                raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
            self.keep.append(t)
            self.any_cuda |= t.is_cuda
            return t.data_ptr()
        a = np.ascontiguousarray(x, dtype=dtype)
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
        self.keep.append(a)
        return a.ctypes.data


class ScoringEngine:
    """Resident catalog shard + the fused scoring / mask / top-k path.

    Parameters mirror ``bb_config`` (include/brickscore.h).
    """

    def __init__(self, n_items: int, dim: int, cf_dim: int = 0, device: int = 0, id_offset: int = 0,
                 k_max: int = 128, max_batch: int = 0, metric: str = "cosine", keep_fp32: bool = True,
                 build_bf16: bool = True, theme_bits: int = 1024, build_split: Optional[bool] = None):
        """``build_split``: the second bf16 term of the operands, which puts ``precision="fp32"`` searches on the tensor
        pipe (None = the library's default: built when it costs at most 4 GiB)."""
        self._lib = L.load()
        cfg = L.bb_config()
        cfg.struct_size = C.sizeof(L.bb_config)
        cfg.device = device
        cfg.n_items = n_items
        cfg.id_offset = id_offset
        cfg.dim = dim
        cfg.cf_dim = cf_dim
        cfg.k_max = k_max
        cfg.max_batch = max_batch
        cfg.theme_bits = theme_bits
        cfg.metric = L.METRICS[metric]
        cfg.flags = (L.BB_CFG_KEEP_FP32 if keep_fp32 else 0) | (L.BB_CFG_BUILD_BF16 if build_bf16 else 0)
        if build_split is None:
            build_split = keep_fp32 and build_bf16 and n_items * ((dim + cf_dim + 63) // 64 * 64) * 2 <= (4 << 30)
        if build_split:
            cfg.flags |= L.BB_CFG_BUILD_SPLIT
        h = C.c_void_p()
        rc = self._lib.bb_create(C.byref(h), C.byref(cfg))
        if rc != L.BB_OK:
            raise L.BrickscoreError(rc, self._lib.bb_last_error(None).decode())
        self._h = h
        self.n_items, self.dim, self.cf_dim = n_items, dim, cf_dim
        self.device, self.id_offset, self.k_max = device, id_offset, k_max
        self.theme_words = theme_bits // 32
        self.n_theme_sets = 0
        self.n_allow = 0
        self.n_excl = 0
        self.comm_rank, self.comm_world = 0, 1

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.bb_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != L.BB_OK:
            raise L.BrickscoreError(rc, self._lib.bb_last_error(self._h).decode())

    def _stream(self, stream) -> Optional[int]:
        if stream is not None:
            return int(stream)
        if torch is not None and torch.cuda.is_available():
            return int(torch.cuda.current_stream(self.device).cuda_stream)
        return None

    def info(self) -> dict:
        i = L.bb_info()
        i.struct_size = C.sizeof(L.bb_info)
        self._check(self._lib.bb_get_info(self._h, C.byref(i)))
        d = {f: getattr(i, f) for f, _ in L.bb_info._fields_ if f != "struct_size"}
        for f in ("epilogue_cycles", "round_cycles", "round_wait_cycles"):
            d[f] = list(d[f])
        return d

    def set_profiling(self, enable: bool = True):
        """Bracket every scan launch with CUDA events; totals appear in ``info()['scan_ms']``."""
        self._check(self._lib.bb_set_profiling(self._h, 1 if enable else 0))

    def comm_export(self) -> bytes:
        """cudaIpc handle of this shard's communication block (row-sharded catalogs, see ShardedSearch)."""
        buf = C.create_string_buffer(L.BB_IPC_HANDLE_BYTES)
        self._check(self._lib.bb_comm_export(self._h, buf))
        return buf.raw

    def comm_import(self, rank: int, handles: Sequence[bytes]):
        """Maps the communication blocks of all ranks (``handles[rank]`` is this rank's own and is ignored): the scan
        then prunes against the union of all shards' thresholds and the merge kernel delivers candidates to the
        queries' owner ranks through NVLink peer memory."""
        blob = b"".join(h if h is not None else b"\0" * L.BB_IPC_HANDLE_BYTES for h in handles)
        self._check(self._lib.bb_comm_import(self._h, rank, len(handles), C.c_char_p(blob)))
        self.comm_rank, self.comm_world = rank, len(handles)

    def comm_link_local(self, rank: int, engines: Sequence["ScoringEngine"]):
        """Same as ``comm_import`` for shards that live in THIS process (several GPUs driven by one process, or several
        shards on one GPU in tests)."""
        arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
        self._check(self._lib.bb_comm_link_local(self._h, rank, len(engines), arr))
        self.comm_rank, self.comm_world = rank, len(engines)

    @property
    def launches(self) -> int:
        return self.info()["launches"]

    # ------------------------------------------------------------------ catalog
    def load_catalog(self, emb=None, item_factors=None, num_parts=None, year=None, theme_id=None,
                     row_begin: int = 0, stream=None):
        ref = emb if emb is not None else item_factors
        n = int(ref.shape[0])
        a = _Args()
        self._check(self._lib.bb_load_catalog(
            self._h, row_begin, n,
            a.ptr(emb, np.float32, (n, self.dim)) if self.dim else None,
            a.ptr(item_factors, np.float32, (n, self.cf_dim)) if self.cf_dim else None,
            a.ptr(num_parts, np.int32, (n,)), a.ptr(year, np.int16, (n,)), a.ptr(theme_id, np.int16, (n,)),
            self._stream(stream)))

    def fetch_embeddings(self, ids) -> np.ndarray:
        a = _Args()
        n = int(np.asarray(ids).shape[0]) if not _is_torch(ids) else int(ids.shape[0])
        out = np.empty((n, self.dim), dtype=np.float32)
        self._check(self._lib.bb_fetch_embeddings(self._h, n, a.ptr(ids, np.int32, (n,)), out.ctypes.data,
                                                  self._stream(None)))
        return out

    def set_exclusion_lists(self, lists: Sequence[Sequence[int]]):
        """lists[i] = sorted global item ids excluded for exclusion-list index i."""
        off = np.zeros(len(lists) + 1, dtype=np.int64)
        for i, l in enumerate(lists):
            off[i + 1] = off[i] + len(l)
        ids = (np.concatenate([np.sort(np.asarray(l, dtype=np.int32)) for l in lists]) if len(lists) and off[-1] > 0
               else np.zeros(0, dtype=np.int32))
        self._check(self._lib.bb_set_exclusion_lists(self._h, len(lists), off.ctypes.data,
                                                     ids.ctypes.data if ids.size else None))
        self.n_excl = len(lists)

    def set_theme_sets(self, bitmaps):
        bm = np.ascontiguousarray(bitmaps, dtype=np.uint32).reshape(-1, self.theme_words)
        self._check(self._lib.bb_set_theme_sets(self._h, bm.shape[0], bm.ctypes.data if bm.size else None))
        self.n_theme_sets = bm.shape[0]

    def set_allow_bitmaps(self, bitmaps):
        """bitmaps: [n, ceil(n_items/32)] uint32 over LOCAL rows (numpy or CUDA tensor)."""
        words = (self.n_items + 31) // 32
        a = _Args()
        n = int(bitmaps.shape[0])
        self._check(self._lib.bb_set_allow_bitmaps(self._h, n, a.ptr(bitmaps, np.uint32, (n, words)) if n else None))
        self.n_allow = n

    @staticmethod
    def pack_bitmap(mask: np.ndarray) -> np.ndarray:
        """bool[n] -> uint32[ceil(n/32)] (bit i of word i//32 = mask[i])."""
        m = np.asarray(mask, dtype=bool)
        pad = (-m.shape[0]) % 32
        if pad:
            m = np.concatenate([m, np.zeros(pad, dtype=bool)])
        return np.packbits(m.reshape(-1, 32), axis=1, bitorder="little").view("<u4").reshape(-1)

    @staticmethod
    def unpack_bitmap(words: np.ndarray, n: int) -> np.ndarray:
        b = np.unpackbits(np.ascontiguousarray(words, dtype="<u4").view(np.uint8), bitorder="little")
        return b[:n].astype(bool)

    # ------------------------------------------------------------------ the hot path
    def _batch(self, a: _Args, q_emb, q_item, q_cf, w_content, w_cf, constraints, self_id, precision, as_l2,
               constraint_index=None, allow_bitmaps=None, excl_lists=None, prenormalized=False):
        ref = next(x for x in (q_emb, q_item, q_cf) if x is not None)
        b = int(ref.shape[0])
        qb = L.bb_query_batch()
        qb.struct_size = C.sizeof(L.bb_query_batch)
        qb.n_queries = b
        qb.q_emb = a.ptr(q_emb, np.float32, (b, self.dim))
        qb.q_item = a.ptr(q_item, np.int32, (b,))
        qb.q_cf = a.ptr(q_cf, np.float32, (b, self.cf_dim))
        if np.isscalar(w_content):
            qb.w_content_all = float(w_content)
        else:
            qb.w_content = a.ptr(w_content, np.float32, (b,))
        if np.isscalar(w_cf):
            qb.w_cf_all = float(w_cf)
        else:
            qb.w_cf = a.ptr(w_cf, np.float32, (b,))
        if constraints is not None and _is_torch(constraints):
            # device-resident descriptors: uint8 tensor [n, 36] holding bb_constraint records
            if constraints.dtype != torch.uint8 or constraints.dim() != 2 or constraints.shape[1] != L.CONSTRAINT_DTYPE.itemsize:
                raise TypeError("device constraints must be a uint8 tensor of shape [n, 36]")
            if constraint_index is None and constraints.shape[0] not in (1, b):
                raise ValueError("constraints must have 1 or n_queries entries")
            t = constraints.contiguous()
            a.keep.append(t)
            a.any_cuda |= t.is_cuda
            qb.constraints = t.data_ptr()
            qb.n_constraints = int(t.shape[0])
        elif constraints is not None:
            c = np.ascontiguousarray(np.atleast_1d(constraints), dtype=L.CONSTRAINT_DTYPE)
            if constraint_index is None and c.shape[0] not in (1, b):
                raise ValueError("constraints must have 1 or n_queries entries")
            a.keep.append(c)
            qb.constraints = c.ctypes.data
            qb.n_constraints = c.shape[0]
        qb.constraint_index = a.ptr(constraint_index, np.int32, (b,))
        qb.self_id = a.ptr(self_id, np.int32, (b,))
        qb.precision = L.PRECISIONS[precision]
        qb.flags = (L.BB_Q_SCORES_AS_L2 if as_l2 else 0) | (L.BB_Q_EMB_PRENORMALIZED if prenormalized else 0)
        if allow_bitmaps is not None:
            # per-CALL validity bitmaps [n, ceil(n_items/32)] (uint32 NumPy, or int32 / CUDA tensor as written by
            # apply_constraints(device_out=True)); bb_constraint.allow_bitmap indexes this array for this call
            words = (self.n_items + 31) // 32
            n_allow = int(allow_bitmaps.shape[0]) if allow_bitmaps.ndim == 2 else 1
            qb.allow_bitmaps = a.ptr(allow_bitmaps.reshape(n_allow, words), np.uint32, (n_allow, words))
            qb.n_allow = n_allow
        if excl_lists is not None:
            # per-QUERY exclusion-list indices [B] or [B, n<=3] (-1 = none): the user's already-rated items
            e = excl_lists if _is_torch(excl_lists) else np.asarray(excl_lists, dtype=np.int32)
            e = e.reshape(b, -1)
            qb.n_excl_per_query = int(e.shape[1])
            qb.excl_lists = a.ptr(e, np.int32, (b, int(e.shape[1])))
        return qb, b

    @staticmethod
    def dedupe_constraints(constraints: np.ndarray):
        """[B] descriptors -> (distinct [C], index int32 [B]) for the ``constraint_index`` form of ``search``."""
        c = np.ascontiguousarray(np.atleast_1d(constraints), dtype=L.CONSTRAINT_DTYPE)
        raw = c.view(np.uint8).reshape(c.shape[0], -1)
        uniq, inv = np.unique(raw, axis=0, return_inverse=True)
        return uniq.reshape(-1).view(L.CONSTRAINT_DTYPE).copy(), inv.astype(np.int32).reshape(-1)

    def search(self, k: int, q_emb=None, q_item=None, q_cf=None, w_content=1.0, w_cf=0.0, constraints=None,
               self_id=None, precision: str = "fp32", as_l2: bool = False, out=None, stream=None,
               constraint_index=None, share_thresholds: bool = False, allow_bitmaps=None, excl_lists=None,
               prenormalized: bool = False, sharded: bool = False) -> Tuple[object, object]:
        """Fused score + blend + mask + top-k.  Returns (scores[B,k] float32, ids[B,k] int32).
        Outputs are CUDA tensors when any input is a CUDA tensor (or ``out`` is given), else NumPy.
        ``allow_bitmaps`` / ``excl_lists``: this call's validity bitmaps and the queries' own exclusion lists.
        ``share_thresholds`` marks a collective search of a row-sharded catalog whose candidates the caller exchanges
        (NCCL path of ``ShardedSearch``); ``sharded=True`` runs the whole collective inside the library
        (``bb_search_sharded``) and returns the results of THIS rank's slice of the queries, [n_mine, k]."""
        a = _Args()
        qb, b = self._batch(a, q_emb, q_item, q_cf, w_content, w_cf, constraints, self_id, precision, as_l2,
                            constraint_index, allow_bitmaps, excl_lists, prenormalized)
        if share_thresholds:
            qb.flags |= L.BB_Q_SHARE_THRESHOLDS
        rows = b
        if sharded and getattr(self, "comm_world", 1) > 1:
            sl = -(-b // self.comm_world)
            rows = max(0, min(sl, b - self.comm_rank * sl))
        alloc = max(rows, 1)     # (an empty slice still takes part in the collective; the library needs valid pointers)
        if out is not None:
            scores, ids = out
            ps, pi = a.ptr(scores, np.float32, (rows, k)), a.ptr(ids, np.int32, (rows, k))
        elif a.any_cuda:
            fs = torch.empty((alloc, k), dtype=torch.float32, device=f"cuda:{self.device}")
            fi = torch.empty((alloc, k), dtype=torch.int32, device=f"cuda:{self.device}")
            ps, pi = fs.data_ptr(), fi.data_ptr()
            scores, ids = fs[:rows], fi[:rows]
        else:
            fs = np.empty((alloc, k), dtype=np.float32)
            fi = np.empty((alloc, k), dtype=np.int32)
            ps, pi = fs.ctypes.data, fi.ctypes.data
            scores, ids = fs[:rows], fi[:rows]
        fn = self._lib.bb_search_sharded if sharded else self._lib.bb_search
        self._check(fn(self._h, C.byref(qb), k, ps, pi, self._stream(stream)))
        return scores, ids

    def sharded_scatter(self, k: int, stream=None, **kwargs) -> int:
        """First half of a collective search (``bb_sharded_scatter``): same query arguments as ``search``.  Returns the
        number of query rows this rank owns.  For one thread that drives several shards: scatter on every shard,
        then ``sharded_collect`` on every shard."""
        a = _Args()
        qb, b = self._batch(a, kwargs.get("q_emb"), kwargs.get("q_item"), kwargs.get("q_cf"), kwargs.get("w_content", 1.0),
                            kwargs.get("w_cf", 0.0), kwargs.get("constraints"), kwargs.get("self_id"),
                            kwargs.get("precision", "fp32"), kwargs.get("as_l2", False), kwargs.get("constraint_index"),
                            kwargs.get("allow_bitmaps"), kwargs.get("excl_lists"), kwargs.get("prenormalized", False))
        prev = kwargs.get("collect_previous_into")
        if prev is not None:
            # pipelined: the previous scatter's owner-side merge (into ``prev``: CUDA tensors [n_mine, k]) is launched
            # inside this call, after this search's prologue and before its scan (``bb_sharded_scatter_after``)
            ps, pi = a.ptr(prev[0], np.float32), a.ptr(prev[1], np.int32)
            self._check(self._lib.bb_sharded_scatter_after(self._h, C.byref(qb), k, ps, pi, self._stream(stream)))
        else:
            self._check(self._lib.bb_sharded_scatter(self._h, C.byref(qb), k, self._stream(stream)))
        sl = -(-b // self.comm_world)
        return max(0, min(sl, b - self.comm_rank * sl))

    def sharded_collect(self, k: int, n_mine: int, device: bool = False, stream=None, out=None):
        """Second half: (scores, ids) [n_mine, k] of this rank's slice of the queries (``out``: CUDA tensors to fill)."""
        alloc = max(n_mine, 1)
        if out is not None:
            a = _Args()
            self._check(self._lib.bb_sharded_collect(self._h, a.ptr(out[0], np.float32), a.ptr(out[1], np.int32), self._stream(stream)))
            return out[0][:n_mine], out[1][:n_mine]
        if device:
            fs = torch.empty((alloc, k), dtype=torch.float32, device=f"cuda:{self.device}")
            fi = torch.empty((alloc, k), dtype=torch.int32, device=f"cuda:{self.device}")
            ps, pi = fs.data_ptr(), fi.data_ptr()
        else:
            fs = np.empty((alloc, k), dtype=np.float32)
            fi = np.empty((alloc, k), dtype=np.int32)
            ps, pi = fs.ctypes.data, fi.ctypes.data
        self._check(self._lib.bb_sharded_collect(self._h, ps, pi, self._stream(stream)))
        return fs[:n_mine], fi[:n_mine]

    def debug_scores(self, q_emb=None, q_item=None, q_cf=None, w_content=1.0, w_cf=0.0,
                     precision: str = "fp32") -> np.ndarray:
        a = _Args()
        qb, b = self._batch(a, q_emb, q_item, q_cf, w_content, w_cf, None, None, precision, False)
        out = np.empty((b, self.n_items), dtype=np.float32)
        self._check(self._lib.bb_debug_scores(self._h, C.byref(qb), out.ctypes.data, self._stream(None)))
        return out

    # ------------------------------------------------------------------ mask / combine / merge
    def apply_constraints(self, constraints, device_out: bool = False):
        """n descriptors -> (bitmaps [n, words] uint32, counts [n] int64)."""
        c = np.ascontiguousarray(np.atleast_1d(constraints), dtype=L.CONSTRAINT_DTYPE)
        n = c.shape[0]
        words = (self.n_items + 31) // 32
        counts = np.zeros(n, dtype=np.int64)
        if device_out:
            bm = torch.empty((n, words), dtype=torch.int32, device=f"cuda:{self.device}")
            pb = bm.data_ptr()
        else:
            bm = np.empty((n, words), dtype=np.uint32)
            pb = bm.ctypes.data
        self._check(self._lib.bb_apply_constraints(self._h, c.ctypes.data, n, pb, counts.ctypes.data,
                                                   self._stream(None)))
        return bm, counts

    def hybrid_combine(self, c_ids, c_scores, f_ids, f_scores, weights, k: int):
        """Reference-semantics combine (recommendation_system.py:789-843).  Inputs [R, len] (id -1 = empty)."""
        a = _Args()
        c_ids = np.atleast_2d(np.asarray(c_ids, dtype=np.int32))
        f_ids = np.atleast_2d(np.asarray(f_ids, dtype=np.int32))
        r = c_ids.shape[0]
        w = np.ascontiguousarray(weights, dtype=np.float64)
        per_row = 1 if w.ndim == 2 else 0
        out_s = np.empty((r, k), dtype=np.float64)
        out_i = np.empty((r, k), dtype=np.int32)
        self._check(self._lib.bb_hybrid_combine(
            self._h, r, c_ids.shape[1], a.ptr(c_ids, np.int32) if c_ids.size else None,
            a.ptr(np.atleast_2d(c_scores), np.float32) if c_ids.size else None, f_ids.shape[1],
            a.ptr(f_ids, np.int32) if f_ids.size else None,
            a.ptr(np.atleast_2d(f_scores), np.float32) if f_ids.size else None, w.ctypes.data, per_row, k,
            out_s.ctypes.data, out_i.ctypes.data, self._stream(None)))
        return out_s, out_i

    def merge_topk(self, scores, ids, k: int, stream=None):
        """scores / ids: [n_lists, B, len] -> top-k per query (cross-shard merge after the all-gather)."""
        a = _Args()
        n_lists, b, ln = (int(s) for s in scores.shape)
        ps, pi = a.ptr(scores, np.float32, (n_lists, b, ln)), a.ptr(ids, np.int32, (n_lists, b, ln))
        if a.any_cuda:
            out_s = torch.empty((b, k), dtype=torch.float32, device=f"cuda:{self.device}")
            out_i = torch.empty((b, k), dtype=torch.int32, device=f"cuda:{self.device}")
            po_s, po_i = out_s.data_ptr(), out_i.data_ptr()
        else:
            out_s = np.empty((b, k), dtype=np.float32)
            out_i = np.empty((b, k), dtype=np.int32)
            po_s, po_i = out_s.ctypes.data, out_i.ctypes.data
        self._check(self._lib.bb_merge_topk(self._h, b, n_lists, ln, ps, pi, k, po_s, po_i, self._stream(stream)))
        return out_s, out_i

    def csr_spmm(self, indptr, indices, values, n_cols: int, x, out=None, stream=None):
        """``Y = A @ x`` for a CSR matrix ``A`` (int64 indptr, int32 indices, float32 values) and a dense float64
        ``x`` [n_cols, L]; all CUDA tensors on this engine's device.  The sparse products of ``cf_training``."""
        for t, dt in ((indptr, torch.int64), (indices, torch.int32), (values, torch.float32), (x, torch.float64)):
            if not (torch.is_tensor(t) and t.is_cuda and t.dtype == dt and t.is_contiguous()):
                raise ValueError("csr_spmm expects contiguous CUDA tensors: int64 indptr, int32 indices, float32 values, float64 x")
        n_rows, l = int(indptr.shape[0]) - 1, int(x.shape[1])
        if int(x.shape[0]) != n_cols:
            raise ValueError("x must have n_cols rows")
        if out is None:
            out = torch.empty((n_rows, l), dtype=torch.float64, device=x.device)
        if indices.numel() == 0:                       # no stored entries: the product is zero
            return out.zero_()
        self._check(self._lib.bb_csr_spmm_f64(self._h, n_rows, n_cols, int(indices.numel()), indptr.data_ptr(), indices.data_ptr(), values.data_ptr(),
                                              x.data_ptr(), l, out.data_ptr(), self._stream(stream)))
        return out

    def item_item_topk(self, indptr, indices, values, n_items: int, k: int, device_out: bool = False):
        """Item-item cosine top-k from a user x item CSR rating matrix (BASELINE config 3).  The CSR arrays may be
        NumPy or CUDA tensors; ``device_out`` leaves the lists on the device."""
        a = _Args()
        n_users = int(indptr.shape[0]) - 1
        if device_out:
            out_s = torch.empty((n_items, k), dtype=torch.float32, device=f"cuda:{self.device}")
            out_i = torch.empty((n_items, k), dtype=torch.int32, device=f"cuda:{self.device}")
            ps, pi = out_s.data_ptr(), out_i.data_ptr()
        else:
            out_s = np.empty((n_items, k), dtype=np.float32)
            out_i = np.empty((n_items, k), dtype=np.int32)
            ps, pi = out_s.ctypes.data, out_i.ctypes.data
        self._check(self._lib.bb_item_item_topk(self._h, n_users, n_items, a.ptr(indptr, np.int64),
                                                a.ptr(indices, np.int32), a.ptr(values, np.float32), k,
                                                ps, pi, self._stream(None)))
        return out_s, out_i
