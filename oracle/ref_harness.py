"""Runs the UNMODIFIED reference (``/root/reference/src/scripts``) in the dev container.

TEST INFRASTRUCTURE ONLY (see ``oracle/brick_oracle.py`` header).  Used by
``oracle/make_golden.py`` to produce ``tests/golden/*.npz`` and by
``tests/test_oracle_vs_reference.py`` (skipped when ``/root/reference`` is
absent, e.g. on the GPU box).

* ``recommendation_system.py`` / ``hard_constraint_filter.py`` import
  ``psycopg2`` at module scope; the package is not installed here, so a stub is
  placed in ``sys.modules`` before import.  No reference source is copied.
* The reference talks to Postgres.  ``SqliteShimConnection`` gives its code a
  psycopg2-shaped connection over an in-memory SQLite database so that the SQL
  the reference GENERATES (``hard_constraint_filter.py:318-480``) is executed
  under real SQL three-valued logic.  Only the placeholder style is rewritten
  (``%s`` -> ``?``; ``= ANY(%s)`` with a list parameter -> ``IN (?,..)``).
"""
from __future__ import annotations

import logging
import os
import re
import sqlite3
import sys
import types
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

REFERENCE_ROOT = os.environ.get("BRICKBRAIN_REFERENCE", "/root/reference")
_SCRIPTS = os.path.join(REFERENCE_ROOT, "src", "scripts")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_SCRIPTS, "recommendation_system.py"))


def load_reference():
    """Returns (recommendation_system, hard_constraint_filter) modules of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "psycopg2" not in sys.modules:
        stub = types.ModuleType("psycopg2")
        extras = types.ModuleType("psycopg2.extras")

        class RealDictCursor:  # marker only; the shim checks identity by name
            pass

        extras.RealDictCursor = RealDictCursor
        stub.extras = extras
        stub.connect = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no postgres here"))
        sys.modules["psycopg2"] = stub
        sys.modules["psycopg2.extras"] = extras
    if _SCRIPTS not in sys.path:
        sys.path.insert(0, _SCRIPTS)
    import importlib
    rs = importlib.import_module("recommendation_system")
    hcf = importlib.import_module("hard_constraint_filter")
    for name in ("recommendation_system", "hard_constraint_filter", "root"):
        logging.getLogger(name if name != "root" else None).setLevel(logging.ERROR)
    return rs, hcf


# --------------------------------------------------------------------------------------
# psycopg2-shaped shim over sqlite3
# --------------------------------------------------------------------------------------

class _ShimCursor:
    def __init__(self, con: sqlite3.Connection, dict_rows: bool):
        self._cur = con.cursor()
        self._dict = dict_rows

    @staticmethod
    def _rewrite(sql: str, params: Sequence[Any]):
        out_sql: List[str] = []
        out_params: List[Any] = []
        pos = 0
        it = iter(params or [])
        for m in re.finditer(r"=\s*ANY\(%s\)|%s", sql):
            out_sql.append(sql[pos:m.start()])
            p = next(it)
            if m.group(0).startswith("="):
                vals = list(p)
                out_sql.append("IN (" + ",".join("?" * len(vals)) + ")")
                out_params.extend(vals)
            else:
                out_sql.append("?")
                out_params.append(p)
            pos = m.end()
        out_sql.append(sql[pos:])
        return "".join(out_sql), out_params

    def execute(self, sql: str, params: Sequence[Any] = ()):  # noqa: D401
        q, p = self._rewrite(sql, params)
        self._cur.execute(q, p)
        return self

    def _wrap(self, row):
        if row is None or not self._dict:
            return row
        return {d[0]: v for d, v in zip(self._cur.description, row)}

    def fetchall(self):
        return [self._wrap(r) for r in self._cur.fetchall()]

    def fetchone(self):
        return self._wrap(self._cur.fetchone())

    @property
    def description(self):          # pandas.read_sql on a DBAPI connection reads the column names from here
        return self._cur.description

    def close(self):
        self._cur.close()


class SqliteShimConnection:
    """In-memory ``sets`` / ``themes`` / ``user_collections`` / ``user_wishlists`` / ``user_interactions``
    (column types as in src/db/rebrickable_schema.sql:38-46 and
    user_interaction_schema.sql)."""

    def __init__(self, set_nums: Sequence[str], names: Sequence[str], year: np.ndarray,
                 theme_id: np.ndarray, num_parts: np.ndarray, themes: Dict[int, str],
                 owned: Optional[Dict[int, Sequence[int]]] = None,
                 wishlisted: Optional[Dict[int, Sequence[int]]] = None,
                 ratings: Optional[Sequence[tuple]] = None):
        self.con = sqlite3.connect(":memory:")
        c = self.con.cursor()
        c.execute("CREATE TABLE themes (id INTEGER PRIMARY KEY, name TEXT NOT NULL, parent_id INTEGER)")
        c.execute("CREATE TABLE sets (set_num TEXT PRIMARY KEY, name TEXT NOT NULL, year INTEGER NOT NULL,"
                  " theme_id INTEGER, num_parts INTEGER, img_url TEXT)")
        c.execute("CREATE TABLE user_collections (user_id INTEGER, set_num TEXT)")
        c.execute("CREATE TABLE user_wishlists (user_id INTEGER, set_num TEXT)")
        c.executemany("INSERT INTO themes VALUES (?,?,NULL)", [(int(k), v) for k, v in themes.items()])
        rows = []
        for i, sn in enumerate(set_nums):
            t = int(theme_id[i])
            rows.append((sn, names[i], int(year[i]), None if t < 0 else t, int(num_parts[i]), None))
        c.executemany("INSERT INTO sets VALUES (?,?,?,?,?,?)", rows)
        for table, d in (("user_collections", owned), ("user_wishlists", wishlisted)):
            if d:
                c.executemany(f"INSERT INTO {table} VALUES (?,?)",
                              [(int(u), set_nums[int(i)]) for u, ids in d.items() for i in ids])
        # user_interactions (user_interaction_schema.sql): the cold-start aggregates of recommendation_system.py:485-534
        c.execute("CREATE TABLE user_interactions (user_id INTEGER, set_num TEXT, rating INTEGER)")
        if ratings:
            c.executemany("INSERT INTO user_interactions VALUES (?,?,?)", [(int(u), set_nums[int(i)], int(r)) for u, i, r in ratings])
        self.con.commit()

    def cursor(self, cursor_factory=None):
        return _ShimCursor(self.con, dict_rows=cursor_factory is not None)

    def commit(self):
        self.con.commit()

    def rollback(self):
        self.con.rollback()


# --------------------------------------------------------------------------------------
# recommenders with injected synthetic state
# --------------------------------------------------------------------------------------

def size_category(num_parts: int) -> str:
    # bucket edges of ContentBasedRecommender._categorize_by_size (recommendation_system.py:164-177)
    return "small" if num_parts < 100 else "medium" if num_parts < 500 else "large" if num_parts < 1000 else "xl"


def make_set_frame(set_nums, names, year, theme_id, num_parts, theme_names):
    import pandas as pd
    return pd.DataFrame({
        "set_num": list(set_nums),
        "name": list(names),
        "year": np.asarray(year, dtype=np.int64),
        "theme_id": np.asarray(theme_id, dtype=np.int64),
        "num_parts": np.asarray(num_parts, dtype=np.int64),
        "img_url": [None] * len(set_nums),
        "theme_name": list(theme_names),
        "size_category": [size_category(int(p)) for p in num_parts],
        "complexity_score": np.asarray(num_parts, dtype=np.float64) / max(1, int(np.max(num_parts))),
    })


def build_content(rs, dbcon, set_frame, feat_matrix: np.ndarray):
    """``ContentBasedRecommender`` with ``set_feat`` / ``feat_matrix`` / ``set_lookup``
    injected (what ``prepare_features`` would build, recommendation_system.py:91-142)."""
    rec = rs.ContentBasedRecommender(dbcon)
    rec.set_feat = set_frame
    rec.feat_matrix = feat_matrix
    rec.set_lookup = dict(zip(range(len(set_frame)), set_frame["set_num"]))
    return rec


def build_cf(rs, dbcon, set_frame, user_ids: Sequence[int], rating_matrix: np.ndarray,
             user_factors: np.ndarray, item_factors: np.ndarray):
    """``CollaborativeFilteringRecommender`` with the pivot table, lookups and
    SVD factors injected (recommendation_system.py:325-337, 399-407)."""
    import pandas as pd
    rec = rs.CollaborativeFilteringRecommender(dbcon)
    set_nums = list(set_frame["set_num"])
    rec.user_item_matrix = pd.DataFrame(rating_matrix, index=list(user_ids), columns=set_nums)
    rec.user_lookup = {u: i for i, u in enumerate(user_ids)}
    rec.item_lookup = {s: i for i, s in enumerate(set_nums)}
    rec.reverse_user_lookup = {i: u for u, i in rec.user_lookup.items()}
    rec.reverse_item_lookup = {i: s for s, i in rec.item_lookup.items()}
    rec.user_factors = user_factors
    rec.item_factors = item_factors
    rec.svd_model = object()           # "trained" sentinel (recommendation_system.py:420-421)
    details = set_frame.set_index("set_num")

    def _get_set_details(set_num):     # replaces one SQL round trip per hit (:535-550)
        r = details.loc[set_num]
        return {"set_num": set_num, "name": r["name"], "year": int(r["year"]),
                "num_parts": int(r["num_parts"]), "img_url": None, "theme_name": r["theme_name"]}

    rec._get_set_details = _get_set_details
    return rec


def build_hybrid(rs, dbcon, content, cf):
    hyb = rs.HybridRecommender(dbcon)
    hyb.content_recommender = content
    hyb.collaborative_recommender = cf
    return hyb
