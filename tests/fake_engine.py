"""Oracle-backed stand-in for ``ScoringEngine`` used ONLY by the CPU test-suite to
exercise host-side logic (adapters, sharding, descriptor compilation) where no GPU
exists.  It evaluates ``bb_constraint`` descriptors exactly as include/brickscore.h
specifies and ranks with the oracle.  Never imported by the product package."""
import numpy as np

import brick_oracle as bo
from brickbrain_rec_engine_b200 import _lib as L


def eval_descriptor(d, num_parts, year, theme_id, theme_sets, allow_bitmaps, excl_lists, id_offset=0):
    n = num_parts.shape[0]
    if int(d["flags"]) & L.BB_C_ALWAYS_FALSE:
        return np.zeros(n, dtype=bool)
    m = (num_parts >= int(d["parts_min"])) & (num_parts <= int(d["parts_max"]))
    m &= (year >= int(d["year_min"])) & (year <= int(d["year_max"]))
    ts = int(d["theme_set"])
    if ts >= 0:
        bits = np.unpackbits(np.ascontiguousarray(theme_sets[ts], dtype="<u4").view(np.uint8), bitorder="little")
        t = theme_id.astype(np.int64)
        ok = (t >= 0) & (t < bits.shape[0])
        ok[ok] = bits[t[ok]].astype(bool)
        m &= ok
    ab = int(d["allow_bitmap"])
    if ab >= 0:
        m &= np.unpackbits(np.ascontiguousarray(allow_bitmaps[ab], dtype="<u4").view(np.uint8), bitorder="little")[:n].astype(bool)
    for e in d["excl_list"]:
        if int(e) >= 0:
            ids = np.asarray(excl_lists[int(e)], dtype=np.int64) - id_offset
            ids = ids[(ids >= 0) & (ids < n)]
            m[ids] = False
    return m


class FakeEngine:
    def __init__(self, n_items, dim, cf_dim=0, device=0, id_offset=0, k_max=128, theme_bits=1024, **_):
        self.n_items, self.dim, self.cf_dim = n_items, dim, cf_dim
        self.device, self.id_offset, self.k_max = device, id_offset, k_max
        self.theme_words = theme_bits // 32
        self.emb = np.zeros((n_items, dim), np.float32)
        self.itf = np.zeros((n_items, cf_dim), np.float32)
        self.num_parts = np.ones(n_items, np.int32)
        self.year = np.zeros(n_items, np.int16)
        self.theme = np.full(n_items, -1, np.int16)
        self.theme_sets, self.allow, self.excl = [], [], []
        self.calls = []

    def load_catalog(self, emb=None, item_factors=None, num_parts=None, year=None, theme_id=None, row_begin=0, stream=None):
        n = (emb if emb is not None else item_factors).shape[0]
        sl = slice(row_begin, row_begin + n)
        if emb is not None:
            self.emb[sl] = bo.l2_normalize_rows(np.asarray(emb, np.float32))
        if item_factors is not None:
            self.itf[sl] = item_factors
        if num_parts is not None:
            self.num_parts[sl] = num_parts
        if year is not None:
            self.year[sl] = year
        if theme_id is not None:
            self.theme[sl] = theme_id

    def fetch_embeddings(self, ids):
        return self.emb[np.asarray(ids) - self.id_offset].copy()

    def set_exclusion_lists(self, lists):
        self.excl = [np.asarray(l) for l in lists]

    def set_theme_sets(self, bitmaps):
        self.theme_sets = list(np.asarray(bitmaps, dtype=np.uint32).reshape(-1, self.theme_words))

    def set_allow_bitmaps(self, bitmaps):
        self.allow = list(np.asarray(bitmaps, dtype=np.uint32))

    pack_bitmap = staticmethod(lambda mask: __import__("brickbrain_rec_engine_b200").ScoringEngine.pack_bitmap(mask))
    unpack_bitmap = staticmethod(lambda w, n: __import__("brickbrain_rec_engine_b200").ScoringEngine.unpack_bitmap(w, n))

    def _valid(self, d):
        return eval_descriptor(d, self.num_parts, self.year, self.theme, self.theme_sets, self.allow, self.excl, self.id_offset)

    dedupe_constraints = staticmethod(lambda c: __import__("brickbrain_rec_engine_b200").ScoringEngine.dedupe_constraints(c))

    def search(self, k, q_emb=None, q_item=None, q_cf=None, w_content=1.0, w_cf=0.0, constraints=None, self_id=None,
               precision="fp32", as_l2=False, out=None, stream=None, constraint_index=None, share_thresholds=False,
               allow_bitmaps=None, excl_lists=None, prenormalized=False, sharded=False):
        self.calls.append("search")
        q_emb, q_cf, q_item, self_id, constraint_index, excl_lists = (
            None if x is None else np.asarray(x) for x in (q_emb, q_cf, q_item, self_id, constraint_index, excl_lists))
        ref = next(x for x in (q_emb, q_item, q_cf) if x is not None)
        b = ref.shape[0]
        qe = None
        if q_item is not None:
            qe = self.emb[np.asarray(q_item) - self.id_offset]
        elif q_emb is not None:
            qe = np.asarray(q_emb, np.float32)
        wc = np.full(b, w_content, np.float32) if np.isscalar(w_content) else np.asarray(w_content, np.float32)
        wf = np.full(b, w_cf, np.float32) if np.isscalar(w_cf) else np.asarray(w_cf, np.float32)
        s = bo.fused_hybrid_scores(qe, None if q_cf is None else np.asarray(q_cf, np.float32), self.emb if qe is not None else None,
                                   self.itf if q_cf is not None else None, wc, wf)
        valid = None
        if constraints is not None:
            c = np.atleast_1d(constraints)
            pick = (lambda j: int(constraint_index[j])) if constraint_index is not None else (lambda j: j if c.shape[0] > 1 else 0)
            allow = self.allow if allow_bitmaps is None else list(np.asarray(allow_bitmaps, dtype=np.uint32).reshape(-1, (self.n_items + 31) // 32))
            valid = np.stack([eval_descriptor(c[pick(j)], self.num_parts, self.year, self.theme, self.theme_sets, allow,
                                              self.excl, self.id_offset) for j in range(b)])
        if excl_lists is not None:                      # the queries' own exclusion lists (already-rated items)
            if valid is None:
                valid = np.ones((b, self.n_items), dtype=bool)
            e = excl_lists.reshape(b, -1)
            for j in range(b):
                for l in e[j]:
                    if l >= 0:
                        ids_l = np.asarray(self.excl[int(l)]) - self.id_offset
                        ids_l = ids_l[(ids_l >= 0) & (ids_l < self.n_items)]
                        valid[j, ids_l] = False
        sid = None
        if self_id is not None:                         # a global id outside this shard excludes nothing here
            sid = np.asarray(self_id, dtype=np.int64) - self.id_offset
            sid = np.where((sid >= 0) & (sid < self.n_items), sid, -1)
        ids, sc = bo.fused_hybrid_topk(s, k, valid=valid, self_ids=sid)
        ids = np.where(ids >= 0, ids + self.id_offset, -1).astype(np.int32)
        sc = sc.astype(np.float32)
        if as_l2:
            sc = np.where(ids >= 0, 2.0 - 2.0 * sc, np.inf).astype(np.float32)
        return sc, ids

    def apply_constraints(self, constraints, device_out=False):
        c = np.atleast_1d(constraints)
        masks = [self._valid(c[j]) for j in range(c.shape[0])]
        return np.stack([self.pack_bitmap(m) for m in masks]), np.array([int(m.sum()) for m in masks], dtype=np.int64)

    def hybrid_combine(self, c_ids, c_scores, f_ids, f_scores, weights, k):
        c_ids, f_ids = np.atleast_2d(c_ids), np.atleast_2d(f_ids)
        c_scores, f_scores = np.atleast_2d(c_scores), np.atleast_2d(f_scores)
        out_s = np.full((c_ids.shape[0], k), -np.inf)
        out_i = np.full((c_ids.shape[0], k), -1, np.int32)
        for r in range(c_ids.shape[0]):
            ids, sc = bo.hybrid_combine(c_ids[r], np.asarray(c_scores[r], np.float32), f_ids[r], np.asarray(f_scores[r], np.float32),
                                        weights[0], weights[1], k)
            out_i[r, :len(ids)], out_s[r, :len(ids)] = ids, sc
        return out_s, out_i

    def merge_topk(self, scores, ids, k, stream=None):
        scores, ids = np.asarray(scores), np.asarray(ids)
        g, b, ln = scores.shape
        out_s = np.full((b, k), -np.inf, np.float32)
        out_i = np.full((b, k), -1, np.int32)
        for r in range(b):
            cand = [(float(scores[l, r, j]), int(ids[l, r, j])) for l in range(g) for j in range(ln) if ids[l, r, j] >= 0]
            cand.sort(key=lambda x: (-x[0], x[1]))
            for j, (s, i) in enumerate(cand[:k]):
                out_s[r, j], out_i[r, j] = s, i
        return out_s, out_i

    def close(self):
        pass
