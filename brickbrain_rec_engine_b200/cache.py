"""Recommendation cache: the ``user_recommendations`` table of the reference
(src/db/user_interaction_schema.sql:72-85) and its two accessors in the API
(``_get_cached_recommendations`` recommendation_api.py:1964-2003, ``_cache_recommendations`` :2005-2036) as a
host-side columnar store, so that bulk generation (BASELINE config 5: 1M users x top-100) lands with one vectorized
upsert per result batch instead of one SQL INSERT per row.

Semantics kept: rows are keyed by (user_id, recommendation_type); caching a user REPLACES that user's rows of the type
(DELETE + INSERT, :2013-2030); rows expire 24 hours after creation (``expires_at`` default, schema :80); a read returns
the unexpired rows ordered by score descending, at most ``top_k`` (:1979-1981), or ``None`` when there are none.
"""
from __future__ import annotations

import time
from typing import Dict, Optional, Tuple

import numpy as np

TTL_SECONDS = 24 * 3600


class RecommendationCache:
    def __init__(self, k_max: int = 128, ttl_seconds: float = TTL_SECONDS, clock=time.time):
        self.k_max = k_max
        self.ttl = ttl_seconds
        self.clock = clock
        self._tables: Dict[str, dict] = {}

    def _table(self, rec_type: str, min_rows: int = 0) -> dict:
        t = self._tables.get(rec_type)
        if t is None:
            t = {"row_of": {}, "n": 0, "ids": np.empty((0, self.k_max), np.int32), "scores": np.empty((0, self.k_max), np.float32),
                 "count": np.empty(0, np.int32), "created": np.empty(0, np.float64), "user": np.empty(0, np.int64)}
            self._tables[rec_type] = t
        if min_rows > t["ids"].shape[0]:
            cap = max(min_rows, 2 * t["ids"].shape[0], 1024)
            for name, fill in (("ids", -1), ("scores", 0), ("count", 0), ("created", 0), ("user", -1)):
                old = t[name]
                new = np.full((cap,) + old.shape[1:], fill, dtype=old.dtype)
                new[: old.shape[0]] = old
                t[name] = new
        return t

    def bulk_upsert(self, user_ids: np.ndarray, ids: np.ndarray, scores: np.ndarray, rec_type: str = "hybrid"):
        """``ids`` / ``scores`` [n, k] (id -1 = empty slot) for ``user_ids`` [n]: replaces each user's cached rows of
        ``rec_type``.  This is the consumer of ``BulkSink`` (``lambda lo, s, i: cache.bulk_upsert(users[lo:lo+len(s)], i, s)``)."""
        user_ids = np.asarray(user_ids, dtype=np.int64)
        n, k = ids.shape
        if k > self.k_max:
            raise ValueError(f"k={k} exceeds the cache's k_max={self.k_max}")
        t = self._table(rec_type)
        row_of = t["row_of"]
        rows = np.empty(n, dtype=np.int64)
        nxt = t["n"]
        for j, u in enumerate(user_ids.tolist()):
            r = row_of.get(u)
            if r is None:
                r = row_of[u] = nxt
                nxt += 1
            rows[j] = r
        t = self._table(rec_type, nxt)
        t["n"] = nxt
        t["ids"][rows, :k] = ids
        t["ids"][rows, k:] = -1
        t["scores"][rows, :k] = scores
        t["count"][rows] = (ids >= 0).sum(axis=1)
        t["created"][rows] = self.clock()
        t["user"][rows] = user_ids

    def put(self, user_id, set_rows, scores, rec_type: str = "hybrid"):
        """One user (``_cache_recommendations``).  A falsy user id is not cached (:2007-2008)."""
        if not user_id:
            return
        ids = np.asarray(set_rows, dtype=np.int32).reshape(1, -1)
        self.bulk_upsert(np.array([user_id]), ids, np.asarray(scores, dtype=np.float32).reshape(1, -1), rec_type)

    def get(self, user_id, rec_type: str = "hybrid", top_k: int = 10) -> Optional[Tuple[np.ndarray, np.ndarray]]:
        """(catalog rows, scores) ordered by score descending, or None (``_get_cached_recommendations``)."""
        if not user_id:
            return None
        t = self._tables.get(rec_type)
        if t is None:
            return None
        r = t["row_of"].get(int(user_id))
        if r is None or self.clock() - t["created"][r] >= self.ttl:
            return None
        c = int(t["count"][r])
        if c == 0:
            return None
        ids, sc = t["ids"][r, : self.k_max], t["scores"][r, : self.k_max]
        keep = ids >= 0
        ids, sc = ids[keep], sc[keep]
        order = np.argsort(-sc, kind="stable")[:top_k]
        return ids[order].copy(), sc[order].copy()

    def __len__(self):
        return sum(t["n"] for t in self._tables.values())
