"""CPU oracle for the Brickbrain candidate-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``brickbrain_rec_engine_b200/`` imports
this module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may call it, and there only as the
checker or the timed CPU baseline -- never as the product path.

It is a NumPy restatement of the reference's algorithm for the path (the
reference is Python; citations are ``file:line`` under ``/root/reference/``).
Pinning status (see DESIGN.md "Oracle"):

* content cosine / CF dot / hybrid combine: pinned against the UNMODIFIED
  reference functions executed in the dev container (``oracle/ref_harness.py``
  imports ``src/scripts/recommendation_system.py`` with a stub ``psycopg2``);
  the resulting vectors are committed under ``tests/golden/`` by
  ``oracle/make_golden.py``.
* constraint mask: pinned against the reference's OWN generated SQL
  (``hard_constraint_filter.py:318-480``) executed on an in-memory SQLite
  database through a psycopg2-shaped shim (``oracle/ref_harness.py``).
* vector-store search (pgvector / Chroma / FAISS live in third-party code that
  is not under /root/reference: langchain-postgres==0.0.12, chromadb==1.0.15,
  faiss-cpu unpinned) and item-item CF (absent from the reference):
  **parity unpinned** -- restated from the published algorithm (exact flat
  search; sklearn cosine on item columns), anchored on the reference call sites
  ``lego_nlp_recommeder.py:296-305,1394`` and the fixture
  ``test_embeddings/index.faiss`` (inputs only, it pins no outputs).

Spec ordering everywhere: descending score, ties broken by LOWER catalog index.
The reference's own tie order is implementation defined (unstable argsort,
Python set iteration) so parity on ids is asserted on tie-free inputs.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# a1  content cosine            recommendation_system.py:213-214
# --------------------------------------------------------------------------------------

def l2_normalize_rows(x: np.ndarray) -> np.ndarray:
    """sklearn.preprocessing.normalize(norm='l2') as reached from
    ``cosine_similarity`` (recommendation_system.py:214): rows are divided by
    their L2 norm; zero-norm rows are divided by 1 (stay zero).  dtype preserved
    (float32 in -> float32 out)."""
    x = np.asarray(x)
    norms = np.sqrt(np.einsum("ij,ij->i", x, x))
    norms = np.where(norms == 0.0, 1.0, norms).astype(x.dtype)
    return x / norms[:, None]


def cosine_scores(target: np.ndarray, feat_matrix: np.ndarray) -> np.ndarray:
    """``cosine_similarity(target[1xD], X[NxD])[0]`` (recommendation_system.py:213-214)."""
    t = l2_normalize_rows(np.asarray(target).reshape(1, -1))
    x = l2_normalize_rows(feat_matrix)
    return (t @ x.T)[0]


def rank_desc(scores: np.ndarray) -> np.ndarray:
    """Spec ordering: indices sorted by (-score, +index).  Stable sort on the
    negated scores == the reference's ``argsort(...)[::-1]`` on tie-free data
    (recommendation_system.py:217)."""
    return np.argsort(-scores, kind="stable")


# --------------------------------------------------------------------------------------
# a2/a3  content ranking + filter walk     recommendation_system.py:217-247
# --------------------------------------------------------------------------------------

def content_similar(feat_matrix: np.ndarray, target_idx: int, top_k: int,
                    valid_mask: Optional[np.ndarray] = None,
                    drop: str = "rank0") -> Tuple[np.ndarray, np.ndarray]:
    """Restates ``ContentBasedRecommender.get_similar_sets``.

    ``drop='rank0'`` is the reference behaviour: the best-ranked row is dropped
    whatever it is (recommendation_system.py:217, ``[::-1][1:]``);
    ``drop='self'`` is the B200 engine's spec (the target id is excluded).  The
    two agree whenever the target is the unique maximum (cos(x,x)=1).
    ``valid_mask`` restates the ``set_num not in valid_set_filter`` walk
    (recommendation_system.py:229); an empty filter list is falsy in the
    reference, i.e. "no filter" -- callers pass ``None`` for that case.
    Returns (ids int64[<=k], scores[<=k]).
    """
    sim = cosine_scores(feat_matrix[target_idx], feat_matrix)
    order = rank_desc(sim)
    if drop == "rank0":
        order = order[1:]
    elif drop == "self":
        order = order[order != target_idx]
    else:
        raise ValueError(drop)
    if valid_mask is not None:
        order = order[valid_mask[order]]
    ids = order[:top_k]
    return ids.astype(np.int64), sim[ids]


# --------------------------------------------------------------------------------------
# a4/a5  collaborative filtering         recommendation_system.py:432-461
# --------------------------------------------------------------------------------------

def cf_scores(user_factors_u: np.ndarray, item_factors: np.ndarray) -> np.ndarray:
    """``np.dot(user_factors[1xF], item_factors.T).flatten()`` (recommendation_system.py:438)."""
    return np.dot(np.asarray(user_factors_u).reshape(1, -1), item_factors.T).ravel()


def cf_recommend(user_factors_u: np.ndarray, item_factors: np.ndarray, top_k: int,
                 rated_mask: Optional[np.ndarray] = None,
                 valid_mask: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Restates the exclusion / filter loop and the stable descending sort of
    ``CollaborativeFilteringRecommender.get_recommendations``
    (recommendation_system.py:441-461).  ``list.sort(reverse=True)`` is stable,
    so ties keep the lower index first == spec ordering."""
    scores = cf_scores(user_factors_u, item_factors)
    keep = np.ones(scores.shape[0], dtype=bool)
    if rated_mask is not None:
        keep &= ~rated_mask
    if valid_mask is not None:
        keep &= valid_mask
    cand = np.nonzero(keep)[0]
    order = cand[np.argsort(-scores[cand], kind="stable")][:top_k]
    return order.astype(np.int64), scores[order]


# --------------------------------------------------------------------------------------
# a8  hybrid combine (reference semantics)   recommendation_system.py:789-843
# --------------------------------------------------------------------------------------

def hybrid_combine(content_ids: Sequence[int], content_scores: Sequence[float],
                   cf_ids: Sequence[int], cf_scores_: Sequence[float],
                   content_weight: float, collaborative_weight: float,
                   top_k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Union of the two (top-2k) lists; a side that is missing contributes 0
    (recommendation_system.py:812-818); python-float (fp64) blend; stable
    descending sort; slice.  Tie order in the reference follows Python ``set``
    iteration (hash dependent); the spec breaks ties by lower catalog index."""
    c = {int(i): float(s) for i, s in zip(content_ids, content_scores)}
    f = {int(i): float(s) for i, s in zip(cf_ids, cf_scores_)}
    ids = sorted(set(c) | set(f))
    hybrid = [content_weight * c.get(i, 0) + collaborative_weight * f.get(i, 0) for i in ids]
    order = sorted(range(len(ids)), key=lambda j: (-hybrid[j], ids[j]))[:top_k]
    return (np.array([ids[j] for j in order], dtype=np.int64),
            np.array([hybrid[j] for j in order], dtype=np.float64))


def normalize_weights(content_weight: float, collaborative_weight: float) -> Tuple[float, float]:
    """``HybridRecommender.set_weights`` (recommendation_system.py:845-849)."""
    total = content_weight + collaborative_weight
    return content_weight / total, collaborative_weight / total


# --------------------------------------------------------------------------------------
# a9/a10  hard-constraint mask        hard_constraint_filter.py:318-480, 505-516
# --------------------------------------------------------------------------------------

INT32_MAX = 2**31 - 1
INT32_MIN = -2**31


def like_to_regex(pattern: str) -> "re.Pattern[str]":
    """SQL LIKE -> regex ('%' any run, '_' one char), case-insensitive, as in
    ``LOWER(name) LIKE LOWER(%s)`` with ``%name%`` (hard_constraint_filter.py:505-512)."""
    out = []
    for ch in pattern:
        if ch == "%":
            out.append(".*")
        elif ch == "_":
            out.append(".")
        else:
            out.append(re.escape(ch))
    return re.compile("^" + "".join(out) + "$", re.IGNORECASE | re.DOTALL)


def resolve_theme_ids(theme_names: Iterable[str], themes: Dict[int, str]) -> List[int]:
    """``HardConstraintFilter._get_theme_ids`` (hard_constraint_filter.py:482-532):
    OR of case-insensitive substring LIKE matches over the ``themes`` table."""
    pats = [like_to_regex(f"%{n}%") for n in theme_names]
    return sorted(tid for tid, name in themes.items() if any(p.match(name) for p in pats))


@dataclass
class MaskSpec:
    """Integer form of an AND of hard constraints -- what the GPU descriptor encodes."""
    parts_min: int = INT32_MIN
    parts_max: int = INT32_MAX
    year_min: int = -32768
    year_max: int = 32767
    theme_active: bool = False              # any theme predicate present -> NULL theme fails
    required_themes: Optional[set] = None   # None = no THEMES_REQUIRED predicate
    excluded_themes: set = field(default_factory=set)
    always_false: bool = False
    exclude_user_owned: Optional[int] = None
    exclude_user_wishlisted: Optional[int] = None
    notes: List[str] = field(default_factory=list)


def compile_constraints(constraints: Sequence[Tuple[str, object]], themes: Dict[int, str],
                        now_year: int) -> MaskSpec:
    """Folds an ordered list of ``(constraint_type_value, value)`` pairs into one
    ``MaskSpec``, restating ``_build_constraint_sql`` + ``_constraint_to_sql``
    (hard_constraint_filter.py:318-480).  Every numeric predicate is an integer
    compare on num_parts / year; an AND of range predicates is a range
    intersection."""
    m = MaskSpec()
    m.parts_min = 1                      # "s.num_parts > 0" is always added (:343)
    required: Optional[set] = None
    excluded: set = set()

    def pmax(v):
        m.parts_max = min(m.parts_max, int(v))

    def pmin(v):
        m.parts_min = max(m.parts_min, int(v))

    for ctype, value in constraints:
        if ctype == "pieces_max":
            pmax(value)                                           # :379-380
        elif ctype == "pieces_min":
            pmin(value)                                           # :382-383
        elif ctype == "year_min":
            m.year_min = max(m.year_min, int(value))              # :385-386
        elif ctype == "year_max":
            m.year_max = min(m.year_max, int(value))              # :388-389
        elif ctype == "price_max":
            pmax(int(value / 0.10))                               # :391-395
        elif ctype == "price_min":
            pmin(int(value / 0.15))                               # :397-400
        elif ctype == "themes_required":
            ids = resolve_theme_ids(value, themes)                # :402-409
            if ids:
                required = set(ids) if required is None else (required & set(ids))
                m.theme_active = True
            else:
                m.always_false = True                             # "1=0"
        elif ctype == "themes_excluded":
            ids = resolve_theme_ids(value, themes)                # :411-418
            if ids:
                excluded |= set(ids)
                m.theme_active = True                             # NOT(NULL = ANY) is NULL -> row fails
        elif ctype == "age_min":
            v = value                                             # :420-430
            if v <= 4:
                pmax(50)
            elif v <= 8:
                pmax(500)
            elif v <= 12:
                pmax(1500)
            else:
                pmin(500)
        elif ctype == "age_max":
            v = value                                             # :432-439
            if v <= 8:
                pmax(300)
            elif v <= 12:
                pmax(800)
        elif ctype == "exclude_owned":
            m.exclude_user_owned = int(value)                     # :441-445
        elif ctype == "exclude_wishlisted":
            m.exclude_user_wishlisted = int(value)                # :447-451
        elif ctype == "complexity_max":
            pmax({"simple": 200, "moderate": 800, "complex": 999999}.get(value, 999999))   # :453-461
        elif ctype == "complexity_min":
            pmin({"simple": 0, "moderate": 200, "complex": 800}.get(value, 0))             # :463-471
        elif ctype == "availability":
            m.year_min = max(m.year_min, now_year - 5)            # :473-477
        else:
            pass                                                  # default: no constraint (:479-480)
    m.required_themes = required
    m.excluded_themes = excluded
    return m


def constraint_mask(spec: MaskSpec, num_parts: np.ndarray, year: np.ndarray, theme_id: np.ndarray,
                    owned: Optional[Dict[int, np.ndarray]] = None,
                    wishlisted: Optional[Dict[int, np.ndarray]] = None) -> np.ndarray:
    """Boolean mask over catalog rows.  ``theme_id == -1`` encodes SQL NULL: it
    fails BOTH ``theme_id = ANY(ids)`` and ``NOT(theme_id = ANY(ids))``
    (three-valued logic; rebrickable_schema.sql:38-46 makes theme_id nullable).
    ``num_parts <= 0`` encodes NULL/zero parts and fails ``num_parts > 0``."""
    n = num_parts.shape[0]
    if spec.always_false:
        return np.zeros(n, dtype=bool)
    mask = (num_parts >= spec.parts_min) & (num_parts <= spec.parts_max)
    mask &= (year >= spec.year_min) & (year <= spec.year_max)
    if spec.theme_active:
        mask &= theme_id >= 0
        if spec.required_themes is not None:
            mask &= np.isin(theme_id, np.array(sorted(spec.required_themes), dtype=np.int64))
        if spec.excluded_themes:
            mask &= ~np.isin(theme_id, np.array(sorted(spec.excluded_themes), dtype=np.int64))
    if spec.exclude_user_owned is not None and owned is not None:
        ids = owned.get(spec.exclude_user_owned)
        if ids is not None and len(ids):
            mask[np.asarray(ids)] = False
    if spec.exclude_user_wishlisted is not None and wishlisted is not None:
        ids = wishlisted.get(spec.exclude_user_wishlisted)
        if ids is not None and len(ids):
            mask[np.asarray(ids)] = False
    return mask


# --------------------------------------------------------------------------------------
# north-star fused path: full-catalog hybrid blend + mask + top-k
#   (BASELINE.md section 5 gate (iii); components are a1 + a4, weights a8)
# --------------------------------------------------------------------------------------

def topk_desc(scores: np.ndarray, valid: Optional[np.ndarray], k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Exact top-k by (-score, +id) over the valid rows; pads with id -1 / score -inf."""
    idx = np.nonzero(valid)[0] if valid is not None else np.arange(scores.shape[0])
    s = scores[idx]
    if idx.shape[0] > 4 * k + 64:
        # argpartition for speed, but keep every element tied with the k-th score
        kth = np.partition(s, -k)[-k]
        keep = s >= kth
        idx, s = idx[keep], s[keep]
    order = np.argsort(-s, kind="stable")[:k]
    ids = np.full(k, -1, dtype=np.int64)
    out = np.full(k, -np.inf, dtype=scores.dtype)
    ids[: order.shape[0]] = idx[order]
    out[: order.shape[0]] = s[order]
    return ids, out


def fused_hybrid_scores(q_emb: Optional[np.ndarray], q_cf: Optional[np.ndarray],
                        emb_unit: Optional[np.ndarray], item_factors: Optional[np.ndarray],
                        w_content: np.ndarray, w_cf: np.ndarray, metric: str = "cosine") -> np.ndarray:
    """fp32 score matrix  s = w_c * cos(q, e) + w_f * (u . v)  for a batch.
    ``emb_unit`` is the row-normalized catalog (``l2_normalize_rows`` applied
    once -- the reference re-normalizes it per query, recommendation_system.py:214,
    with the same result)."""
    b = (q_emb if q_emb is not None else q_cf).shape[0]
    n = (emb_unit if emb_unit is not None else item_factors).shape[0]
    s = np.zeros((b, n), dtype=np.float32)
    if q_emb is not None and emb_unit is not None:
        q = l2_normalize_rows(q_emb.astype(np.float32)) if metric == "cosine" else q_emb.astype(np.float32)
        s += np.asarray(w_content, dtype=np.float32).reshape(-1, 1) * (q @ emb_unit.T)
    if q_cf is not None and item_factors is not None:
        s += np.asarray(w_cf, dtype=np.float32).reshape(-1, 1) * (q_cf.astype(np.float32) @ item_factors.astype(np.float32).T)
    return s


def fused_hybrid_topk(scores: np.ndarray, k: int, valid: Optional[np.ndarray] = None,
                      self_ids: Optional[np.ndarray] = None,
                      exclusions: Optional[Sequence[Optional[np.ndarray]]] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Row-wise top-k of a score matrix with a per-row validity mask
    (``valid`` is [B,N] bool or [N] bool), per-row self exclusion and per-row
    exclusion id lists (already-rated / owned / wishlisted)."""
    b, n = scores.shape
    ids = np.full((b, k), -1, dtype=np.int64)
    out = np.full((b, k), -np.inf, dtype=scores.dtype)
    for r in range(b):
        v = np.ones(n, dtype=bool) if valid is None else (valid[r] if valid.ndim == 2 else valid).copy()
        if self_ids is not None and self_ids[r] >= 0:
            v[self_ids[r]] = False
        if exclusions is not None and exclusions[r] is not None and len(exclusions[r]):
            v[np.asarray(exclusions[r])] = False
        ids[r], out[r] = topk_desc(scores[r], v, k)
    return ids, out


# --------------------------------------------------------------------------------------
# a11/a12  vector-store search (third-party in the reference; parity unpinned)
# --------------------------------------------------------------------------------------

def vector_search(query: np.ndarray, emb: np.ndarray, k: int, metric: str = "cosine") -> Tuple[np.ndarray, np.ndarray]:
    """Exact flat search as performed by the stores behind
    lego_nlp_recommeder.py:296-305 / :1394.  ``cosine``: pgvector ``<=>`` ranks by
    cosine distance 1-cos -> we return cos, descending.  ``l2``: FAISS
    ``IndexFlatL2`` returns squared L2, ascending; on unit vectors that is
    2-2cos, i.e. the same ranking -- we return (ids, squared distances)."""
    q = np.asarray(query, dtype=np.float32).reshape(1, -1)
    if metric == "cosine":
        s = cosine_scores(q[0], emb.astype(np.float32))
        ids, sc = topk_desc(s, None, k)
        return ids, sc
    if metric == "ip":
        s = (q @ emb.astype(np.float32).T)[0]
        return topk_desc(s, None, k)
    if metric == "l2":
        e = emb.astype(np.float32)
        d = (np.einsum("ij,ij->i", e, e) - 2.0 * (q @ e.T)[0] + float((q @ q.T)[0, 0])).astype(np.float32)
        ids, neg = topk_desc(-d, None, k)
        return ids, -neg
    raise ValueError(metric)


# --------------------------------------------------------------------------------------
# a13  item-item CF (absent in the reference; BASELINE config 3; parity unpinned)
# --------------------------------------------------------------------------------------

def item_item_topk(indptr: np.ndarray, indices: np.ndarray, data: np.ndarray, n_items: int,
                   k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Cosine between item columns of the user x item rating CSR matrix, top-k
    neighbours per item excluding the item itself; the same sklearn primitive
    the reference uses at recommendation_system.py:214 (``cosine_similarity``).
    Dense N x N intermediate -> only for oracle-sized inputs."""
    from scipy.sparse import csr_matrix
    from sklearn.metrics.pairwise import cosine_similarity
    n_users = indptr.shape[0] - 1
    x = csr_matrix((data.astype(np.float64), indices, indptr), shape=(n_users, n_items))
    sim = cosine_similarity(x.T.tocsr(), dense_output=True)
    ids = np.full((n_items, k), -1, dtype=np.int64)
    out = np.full((n_items, k), -np.inf, dtype=np.float64)
    for i in range(n_items):
        v = np.ones(n_items, dtype=bool)
        v[i] = False
        ids[i], out[i] = topk_desc(sim[i], v, k)
    return ids, out


# --------------------------------------------------------------------------------------
# CPU baselines for bench.py (timed, never the product path)
# --------------------------------------------------------------------------------------

def truncated_svd_factors(ratings: np.ndarray, n_components: int = 50, n_iter: int = 5, n_oversamples: int = 10,
                          random_state: int = 42) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """CF factor training: ``train_svd_model`` (recommendation_system.py:368-409) =
    ``TruncatedSVD(n_components, random_state=42).fit_transform(user_item_matrix)`` and ``components_.T``
    (recommendation_system.py:399-407).  The arithmetic is scikit-learn's (third-party, not under the reference
    tree; 1.9.0 here, reference pin >= 1.3): ``sklearn.utils.extmath._randomized_svd`` with the TruncatedSVD
    defaults (algorithm="randomized", n_iter=5, n_oversamples=10, power_iteration_normalizer="auto" -> LU,
    transpose="auto"), ``svd_flip(u_based_decision=False)``, ``X_transformed = X @ components_.T``.
    Pinned: reproduces the factors the reference produced for tests/golden/cf_golden.npz bit for bit.
    Returns (user_factors [U, F], item_factors [N, F], singular_values [F])."""
    from scipy import linalg
    m = np.asarray(ratings, dtype=np.float64)
    rs = np.random.RandomState(random_state)
    n_random = n_components + n_oversamples
    transpose = m.shape[0] < m.shape[1]
    w = m.T if transpose else m
    q = rs.normal(size=(w.shape[1], n_random))
    for _ in range(n_iter):
        q, _ = linalg.lu(w @ q, permute_l=True, check_finite=False)
        q, _ = linalg.lu(w.T @ q, permute_l=True, check_finite=False)
    q, _ = linalg.qr(w @ q, mode="economic", check_finite=False)
    b = q.T @ w
    uhat, s, vt = linalg.svd(b, full_matrices=False, lapack_driver="gesdd")
    u = q @ uhat
    if transpose:
        vt_f, s_f = u[:, :n_components].T.copy(), s[:n_components]
    else:
        vt_f, s_f = vt[:n_components].copy(), s[:n_components]
    idx = np.argmax(np.abs(vt_f), axis=1)
    sign = np.sign(vt_f[np.arange(vt_f.shape[0]), idx])
    sign[sign == 0] = 1.0
    vt_f *= sign[:, None]
    return m @ vt_f.T, vt_f.T.copy(), s_f


def reference_style_query(q_emb: np.ndarray, q_cf: np.ndarray, emb: np.ndarray, item_factors: np.ndarray,
                          valid_ids: Optional[set], w_c: float, w_f: float, top_k: int):
    """One query the way the reference executes it (no batch API): per-query
    ``cosine_similarity`` over the whole catalog (re-normalizing it, :214), full
    argsort (:217), Python filter walk (:222-247), ``np.dot`` + Python loop +
    list.sort (:438-461), dict union blend (:789-843)."""
    from sklearn.metrics.pairwise import cosine_similarity
    sim = cosine_similarity(q_emb.reshape(1, -1), emb)[0]
    order = np.argsort(sim)[::-1]
    content = []
    for i in order:
        if len(content) >= 2 * top_k:
            break
        if valid_ids is not None and int(i) not in valid_ids:
            continue
        content.append((int(i), float(sim[i])))
    scores = np.dot(q_cf.reshape(1, -1), item_factors.T).flatten()
    item_scores = []
    for idx in range(len(scores)):
        if valid_ids is not None and idx not in valid_ids:
            continue
        item_scores.append((idx, scores[idx]))
    item_scores.sort(key=lambda x: x[1], reverse=True)
    collab = [(i, float(s)) for i, s in item_scores[: 2 * top_k]]
    return hybrid_combine([i for i, _ in content], [s for _, s in content],
                          [i for i, _ in collab], [s for _, s in collab], w_c, w_f, top_k)


def batched_cpu_topk(q_cat: np.ndarray, e_cat: np.ndarray, k: int,
                     valid: Optional[np.ndarray] = None, block: int = 65536,
                     class_valid: Optional[np.ndarray] = None, class_index: Optional[np.ndarray] = None,
                     exclusions: Optional[np.ndarray] = None, self_ids: Optional[np.ndarray] = None
                     ) -> Tuple[np.ndarray, np.ndarray]:
    """Best-effort CPU restatement used as ``cpu_baseline`` (kind "port"):
    blocked fp32 sgemm (all BLAS threads) over the concatenated operands
    [w_c*q_hat | w_f*u] x [e_hat | v], validity mask, and a running top-k: the first
    block is partitioned, later blocks only keep scores above each row's current k-th
    best (vectorized compare + one sort of the survivors), so the selection cost stays
    well below the sgemm.  BASELINE.md section 4 item 2.

    Masks: ``valid`` [N] shared by all queries, or -- the GPU arm's workload --
    ``class_valid`` [C, N] (one mask per distinct constraint set,
    hard_constraint_filter.py:318-480) with ``class_index`` [B]; ``exclusions`` [B, R]
    already-rated item ids per query (recommendation_system.py:441-451) and
    ``self_ids`` [B] (the rank-0 drop of :217)."""
    b = q_cat.shape[0]
    n = e_cat.shape[0]
    best_s = np.full((b, k), -np.inf, dtype=np.float32)
    best_i = np.full((b, k), -1, dtype=np.int64)
    ninf = np.float32(-np.inf)
    penalty = None if valid is None else np.where(valid, np.float32(0), ninf).astype(np.float32)
    class_pen = None
    if class_valid is not None:
        class_pen = np.where(class_valid, np.float32(0), ninf).astype(np.float32)
        groups = [np.nonzero(class_index == c)[0] for c in range(class_valid.shape[0])]
    rows_b = np.arange(b)
    for lo in range(0, n, block):
        hi = min(n, lo + block)
        s = q_cat @ e_cat[lo:hi].T
        if penalty is not None:
            s += penalty[None, lo:hi]
        if class_pen is not None:
            for c, g in enumerate(groups):
                if g.size:
                    s[g] += class_pen[c, lo:hi][None, :]
        if exclusions is not None:
            m = (exclusions >= lo) & (exclusions < hi)
            rr = np.broadcast_to(rows_b[:, None], exclusions.shape)
            s[rr[m], exclusions[m] - lo] = ninf
        if self_ids is not None:
            m = (self_ids >= lo) & (self_ids < hi)
            s[rows_b[m], self_ids[m] - lo] = ninf
        if lo == 0:
            kk = min(k, hi - lo)
            part = np.argpartition(-s, kk - 1, axis=1)[:, :kk]
            best_s[:, :kk] = np.take_along_axis(s, part, axis=1)
            best_i[:, :kk] = part
            continue
        thr = best_s.min(axis=1)
        rows, cols = np.nonzero(s > thr[:, None])
        if rows.size == 0:
            continue
        # merge the survivors with the running lists: one sort by (row, -score), keep the first k per row
        all_r = np.concatenate([np.repeat(np.arange(b), k), rows])
        all_s = np.concatenate([best_s.ravel(), s[rows, cols]])
        all_i = np.concatenate([best_i.ravel(), cols + lo])
        order = np.lexsort((-all_s, all_r))
        all_r, all_s, all_i = all_r[order], all_s[order], all_i[order]
        start = np.searchsorted(all_r, np.arange(b))
        rank = np.arange(all_r.size) - start[all_r]
        keep = rank < k
        best_s = all_s[keep].reshape(b, k)
        best_i = all_i[keep].reshape(b, k)
    best_i = np.where(np.isneginf(best_s), -1, best_i)
    order = np.lexsort((best_i, -best_s), axis=1)
    return np.take_along_axis(best_i, order, axis=1), np.take_along_axis(best_s, order, axis=1)


# --------------------------------------------------------------------------------------
# a11 / a12 -- post-filters of the semantic-search callers (results of the vector store -> what the caller returns)
# --------------------------------------------------------------------------------------

def nlp_post_filter(metadatas: Sequence[dict], filters: dict) -> List[int]:
    """``NLPRecommender._apply_filters`` (lego_nlp_recommeder.py:1514-1549) over the metadata dicts of the retrieved
    documents (built at :245-255), returning the indices that survive, in order.  Truthiness rules are the
    reference's: a filter value of 0 / None / '' / [] is "not set"; ``min_age`` of any truthy value means
    ``year >= 2010``; themes match by case-insensitive substring; complexity by equality."""
    keep = []
    for i, m in enumerate(metadatas):
        if filters.get('min_pieces') and m['num_parts'] < filters['min_pieces']:
            continue
        if filters.get('max_pieces') and m['num_parts'] > filters['max_pieces']:
            continue
        if filters.get('themes'):
            if not any(t.lower() in m['theme'].lower() for t in filters['themes']):
                continue
        if filters.get('complexity') and m.get('complexity') != filters['complexity']:
            continue
        if filters.get('min_age') and m['year'] < 2010:
            continue
        keep.append(i)
    return keep


def semantic_search_post(ranked_ids: Sequence[int], metadatas: Sequence[dict], filters: dict, top_k: int,
                         retrieve_k: int = 20) -> List[int]:
    """``semantic_search`` after the embedding step (lego_nlp_recommeder.py:1394-1401): the retriever returns the best
    ``retrieve_k`` documents (k=20, :305), the post-filter runs on THOSE, then ``[:top_k]`` -- so fewer than top_k
    results come back when the filter removes retrieved documents (nothing is re-fetched).  ``ranked_ids``: the exact
    flat-search ranking of the catalog for the query; returns catalog indices."""
    got = list(ranked_ids[:retrieve_k])
    keep = nlp_post_filter([metadatas[i] for i in got], filters)
    return [got[j] for j in keep][:top_k]


def hf_post_filter(results: Sequence[dict], filters: dict) -> List[int]:
    """``HuggingFaceNLPRecommender._apply_filters`` (hf_nlp_recommender.py:1351-1386) over result dicts; returns the
    surviving indices (all of them when ``filters`` is falsy).  Exact theme-name membership; every predicate is skipped
    when either side is falsy (a result without ``min_age`` passes the age filter)."""
    if not filters:
        return list(range(len(results)))
    keep = []
    for i, r in enumerate(results):
        ok = True
        if filters.get('themes') and r.get('theme') and r['theme'] not in filters['themes']:
            ok = False
        if filters.get('min_pieces') and r.get('num_parts') and r['num_parts'] < filters['min_pieces']:
            ok = False
        if filters.get('max_pieces') and r.get('num_parts') and r['num_parts'] > filters['max_pieces']:
            ok = False
        if filters.get('min_age') and r.get('min_age') and r['min_age'] > filters['min_age']:
            ok = False
        if filters.get('max_age') and r.get('max_age') and r['max_age'] < filters['max_age']:
            ok = False
        if ok:
            keep.append(i)
    return keep
