"""Recommender entry points with the reference's signatures, backed by the GPU engine.

Mirrors ``src/scripts/recommendation_system.py``:

* ``RecommendationResult`` / ``RecommendationRequest`` (:25-59)
* ``ContentBasedRecommender.get_similar_sets(set_num, top_k=10, valid_set_filter=None)`` (:194-249)
* ``CollaborativeFilteringRecommender.get_recommendations(user_id, top_k=10, valid_set_filter=None)`` (:411-483)
* ``HybridRecommender.get_recommendations / get_recommendations_from_request /
  _combine_recommendations / set_weights`` (:612-721, :789-849)

What moved to the device: cosine / dot scoring, the rank-0 drop, the
``valid_set_filter`` walk, the already-rated exclusion, top-k and the hybrid
combine.  What stays on the host: result hydration (from an in-memory metadata
table instead of one SQL round trip per hit, :470, :535-550) and the reference's
log-and-fallback behaviour.  Data preparation (SQL feature engineering, SVD
training) is out of scope: state is injected with ``from_arrays`` /
``accelerate``.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from .constraints import ConstraintResult, HardConstraint, HardConstraintFilter
from .engine import ScoringEngine

logger = logging.getLogger(__name__)


@dataclass
class RecommendationResult:
    set_num: str
    name: str
    score: float
    reasons: List[str]
    theme_name: str
    year: int
    num_parts: int
    img_url: Optional[str] = None
    constraint_violations: Optional[List[str]] = None


@dataclass
class RecommendationRequest:
    user_id: Optional[int] = None
    liked_set: Optional[str] = None
    top_k: int = 10
    price_max: Optional[float] = None
    price_min: Optional[float] = None
    pieces_max: Optional[int] = None
    pieces_min: Optional[int] = None
    age_min: Optional[int] = None
    age_max: Optional[int] = None
    year_min: Optional[int] = None
    year_max: Optional[int] = None
    required_themes: Optional[List[str]] = None
    excluded_themes: Optional[List[str]] = None
    max_complexity: Optional[str] = None
    min_complexity: Optional[str] = None
    must_be_available: bool = False
    exclude_owned: bool = False
    exclude_wishlisted: bool = False
    # soft preferences: part of the reference's dataclass (recommendation_system.py:62-65), read by nothing on this path
    preferred_themes: Optional[List[str]] = None
    preferred_complexity: Optional[str] = None
    budget_preference: Optional[float] = None


def _size_category(num_parts: int) -> str:
    return "small" if num_parts < 100 else "medium" if num_parts < 500 else "large" if num_parts < 1000 else "xl"


class CatalogTable:
    """Host metadata for hydration: the columns the reference selects per hit
    (set_num, name, year, num_parts, img_url, theme_name; recommendation_system.py:538-543)
    plus what ``_generate_content_reasons`` reads (:251-276)."""

    def __init__(self, set_nums: Sequence[str], names: Sequence[str], year, num_parts, theme_id,
                 theme_names: Sequence[str], img_urls: Optional[Sequence[Optional[str]]] = None,
                 complexity_score=None, size_category: Optional[Sequence[str]] = None):
        self.set_nums = list(set_nums)
        self.names = list(names)
        self.year = np.asarray(year, dtype=np.int64)
        self.num_parts = np.asarray(num_parts, dtype=np.int64)
        self.theme_id = np.asarray(theme_id, dtype=np.int64)
        self.theme_names = list(theme_names)
        self.img_urls = list(img_urls) if img_urls is not None else [None] * len(self.set_nums)
        self.size_category = (list(size_category) if size_category is not None
                              else [_size_category(int(p)) for p in self.num_parts])
        self.complexity_score = (np.asarray(complexity_score, dtype=np.float64) if complexity_score is not None
                                 else self.num_parts / max(1, int(self.num_parts.max(initial=1))))
        self.index_of = {s: i for i, s in enumerate(self.set_nums)}

    def __len__(self):
        return len(self.set_nums)

    @classmethod
    def from_frame(cls, set_feat) -> "CatalogTable":
        """From the reference's ``set_feat`` DataFrame (recommendation_system.py:115-142)."""
        g = set_feat
        return cls(g["set_num"].tolist(), g["name"].tolist(), g["year"].to_numpy(), g["num_parts"].to_numpy(),
                   g["theme_id"].fillna(-1).to_numpy(), g["theme_name"].tolist(),
                   g["img_url"].tolist() if "img_url" in g else None,
                   g["complexity_score"].to_numpy() if "complexity_score" in g else None,
                   g["size_category"].tolist() if "size_category" in g else None)

    def result(self, idx: int, score: float, reasons: List[str]) -> RecommendationResult:
        return RecommendationResult(set_num=self.set_nums[idx], name=self.names[idx], score=score, reasons=reasons,
                                    theme_name=self.theme_names[idx], year=int(self.year[idx]),
                                    num_parts=int(self.num_parts[idx]), img_url=self.img_urls[idx])


class DeviceFilter:
    """A ``valid_set_filter`` that already lives on the GPU: the validity bitmap written by the mask kernel
    (``HardConstraintFilter.apply_constraints``).  ``HybridRecommender.get_recommendations`` hands it to both
    recommenders, so the request's filter never makes the round trip bitmap -> list of set_nums -> bitmap."""

    def __init__(self, bitmap, count: int):
        self.bitmap = bitmap            # CUDA tensor int32 [1, words] (or uint32 NumPy on CPU test doubles)
        self.count = int(count)

    def __len__(self):
        return self.count

    def __bool__(self):
        return self.count > 0


def _filter_bitmap(catalog: CatalogTable, valid_set_filter):
    """``valid_set_filter`` (list of set_nums, an already packed uint32 bitmap, or a ``DeviceFilter``) -> packed bitmap.
    A falsy filter means "no filter", as in the reference (``if valid_set_filter and ...``, :229, :454)."""
    if valid_set_filter is None:
        return None
    if isinstance(valid_set_filter, DeviceFilter):
        return valid_set_filter.bitmap if valid_set_filter else None
    if isinstance(valid_set_filter, np.ndarray) and valid_set_filter.dtype == np.uint32:
        return valid_set_filter
    if len(valid_set_filter) == 0:
        return None
    mask = np.zeros(len(catalog), dtype=bool)
    mask[[catalog.index_of[s] for s in valid_set_filter if s in catalog.index_of]] = True
    return ScoringEngine.pack_bitmap(mask)


def _clamp_k(top_k: int, engine) -> int:
    """The engine selects at most ``k_max`` (<= 128) results per query; the reference has no such limit (its API caps
    top_k at 50, recommendation_api.py:243, i.e. 100 for the hybrid's top_k * 2).  Larger requests are truncated loudly."""
    if top_k > engine.k_max:
        logger.warning("top_k=%d exceeds the engine's k_max=%d: returning the best %d", top_k, engine.k_max, engine.k_max)
        return engine.k_max
    return top_k


class ContentBasedRecommender:
    """Content similarity over the resident feature / embedding matrix."""

    def __init__(self, dbcon=None, engine: Optional[ScoringEngine] = None, catalog: Optional[CatalogTable] = None,
                 precision: str = "fp32"):
        self.dbcon = dbcon
        self.engine = engine
        self.catalog = catalog
        self.precision = precision
        self.set_lookup = dict(enumerate(catalog.set_nums)) if catalog is not None else {}
        self.feat_matrix = None       # truthiness marker kept for reference-style callers

    def prepare_features(self):
        raise RuntimeError("feature preparation (SQL + feature engineering, recommendation_system.py:91-192) is host-side "
                           "and out of scope: build with from_arrays()/accelerate() instead")

    def get_similar_sets(self, set_num: str, top_k: int = 10,
                         valid_set_filter: Optional[List[str]] = None) -> List[RecommendationResult]:
        if self.engine is None or self.catalog is None:
            self.prepare_features()
        idx = self.catalog.index_of.get(set_num)
        if idx is None:
            logger.error("Set %s not found", set_num)                      # :206-210
            return []
        ids, scores = self._search([idx], top_k, _filter_bitmap(self.catalog, valid_set_filter))
        return self._hydrate(idx, ids[0], scores[0])

    def _search(self, target_rows: Sequence[int], top_k: int, bitmap: Optional[np.ndarray]):
        """Batched: rows of the catalog as queries; self excluded (== the reference's rank-0 drop whenever
        the target is its own best match, :217)."""
        rows = np.asarray(target_rows, dtype=np.int32) + np.int32(self.engine.id_offset)
        cons = allow = None
        if bitmap is not None:                      # the request's filter travels WITH the call (no handle state)
            allow = bitmap.reshape(1, -1)
            cons = L.no_constraint()
            cons["allow_bitmap"] = 0
        scores, ids = self.engine.search(_clamp_k(top_k, self.engine), q_item=rows, self_id=rows, w_content=1.0,
                                         w_cf=0.0, constraints=cons, allow_bitmaps=allow, precision=self.precision)
        if hasattr(scores, "cpu"):                  # (a device-resident filter makes the outputs CUDA tensors)
            scores, ids = scores.cpu().numpy(), ids.cpu().numpy()
        return ids, scores

    def _hydrate(self, target: int, ids: np.ndarray, scores: np.ndarray) -> List[RecommendationResult]:
        c = self.catalog
        out = []
        for i, s in zip(ids, scores):
            if i < 0:
                break
            out.append(c.result(int(i), float(s), self._generate_content_reasons(int(i), target)))
        return out

    def _generate_content_reasons(self, similar: int, target: int) -> List[str]:
        c = self.catalog                                                     # rules of :251-276
        reasons = []
        if c.theme_id[similar] == c.theme_id[target]:
            reasons.append(f"Same theme: {c.theme_names[similar]}")
        if c.size_category[similar] == c.size_category[target]:
            reasons.append(f"Same size category: {c.size_category[similar]}")
        if abs(c.complexity_score[similar] - c.complexity_score[target]) < 0.2:
            reasons.append("Similar complexity level")
        if abs(int(c.year[similar]) - int(c.year[target])) <= 3:
            reasons.append("From similar era")
        return reasons


class CollaborativeFilteringRecommender:
    """User-factor . item-factor scoring with already-rated exclusion."""

    CF_REASON = "Users with similar preferences also liked this set"

    def __init__(self, dbcon=None, engine: Optional[ScoringEngine] = None, catalog: Optional[CatalogTable] = None,
                 user_ids: Optional[Sequence[Any]] = None, user_factors: Optional[np.ndarray] = None,
                 rated: Optional[Dict[Any, np.ndarray]] = None, rated_list_index: Optional[Dict[Any, int]] = None,
                 cf_universe: Optional[np.ndarray] = None, ratings_summary: Optional[dict] = None,
                 precision: str = "fp32"):
        self.dbcon = dbcon
        self.engine = engine
        self.catalog = catalog
        self.user_lookup = {u: i for i, u in enumerate(user_ids)} if user_ids is not None else {}
        self.user_factors = None if user_factors is None else np.ascontiguousarray(user_factors, dtype=np.float32)
        self.rated = rated or {}
        self.rated_list_index = rated_list_index or {}
        self.cf_universe = cf_universe            # bool[N]: rows that have item factors (pivot columns), or None = all
        self.ratings_summary = ratings_summary    # {'avg': float[N], 'count': int[N]} for the cold-start fallback
        self.svd_model = object() if user_factors is not None else None
        self.precision = precision

    def train_svd_model(self, n_components: int = 50):
        """recommendation_system.py:368-409 on the GPU (``cf_training.truncated_svd``): same guards, same factors.
        Needs ``self.user_item_matrix`` (users x catalog rows, 0 = unrated; ndarray, DataFrame or scipy sparse)."""
        from .cf_training import truncated_svd
        m = getattr(self, "user_item_matrix", None)
        if m is None:
            logger.warning("train_svd_model: no user_item_matrix injected")
            self.svd_model = None
            return
        m = m.values if hasattr(m, "values") else m
        if m.shape[0] < 3 or m.shape[1] < 3:
            logger.warning(f"Insufficient data for SVD: {m.shape}. Need at least 3 users and 3 items.")
            self.svd_model = None
            return
        n_components = min(n_components, min(m.shape) - 1)
        if n_components < 2:
            logger.warning("Not enough components for SVD, using fallback recommendations")
            self.svd_model = None
            return
        frame = getattr(self, "user_item_matrix", None)
        self.user_factors, self.item_factors, _ = truncated_svd(self.engine, m, n_components)
        self.user_factors = np.ascontiguousarray(self.user_factors, dtype=np.float32)
        self._adopt_trained_factors(frame, m)
        self.svd_model = object()
        logger.info(f"Trained SVD model with {n_components} components")

    def _adopt_trained_factors(self, frame, m):
        """The reference's ``train_svd_model`` leaves the recommender ready to score (:402-407 + the lookups of
        :332-337).  Here that means: the new item factors go into the engine's operand (factor-only reload), the user
        lookup follows the matrix index, and the already-rated lists / the CF universe are rebuilt from the matrix."""
        eng = self.engine
        f = int(self.item_factors.shape[1])
        self._engine_mismatch = None
        if eng.cf_dim < f or self.catalog is None:
            # the factors were trained but cannot be loaded: scoring must fail loudly instead of using stale operands
            self._engine_mismatch = (f"engine was created with cf_dim={eng.cf_dim} but the trained model has {f} components: "
                                     "create the engine with cf_dim >= n_components (HybridRecommender.from_arrays(cf_dim=...))"
                                     if self.catalog is not None else "no catalog table attached to this recommender")
            logger.warning("train_svd_model: %s", self._engine_mismatch)
            return
        n = len(self.catalog)
        # columns of the matrix -> catalog rows (a pivot table names them by set_num, a plain array is positional)
        if hasattr(frame, "columns") and hasattr(frame, "index"):
            cols = np.array([self.catalog.index_of.get(c, -1) for c in frame.columns], dtype=np.int64)
            users = list(frame.index)
        else:
            cols = np.arange(m.shape[1], dtype=np.int64)
            users = list(self.user_lookup) if len(self.user_lookup) == m.shape[0] else list(range(m.shape[0]))
        keep = cols >= 0
        full = np.zeros((n, eng.cf_dim), dtype=np.float32)
        full[cols[keep], :f] = np.asarray(self.item_factors, dtype=np.float32)[keep]
        eng.load_catalog(item_factors=full)                       # factor-only reload: embeddings and metadata are kept
        if eng.cf_dim > f:
            self.user_factors = np.ascontiguousarray(np.pad(self.user_factors, ((0, 0), (0, eng.cf_dim - f))))
        self.user_lookup = {u: i for i, u in enumerate(users)}
        self.cf_universe = np.zeros(n, dtype=bool)
        self.cf_universe[cols[keep]] = True
        dense = m.toarray() if hasattr(m, "toarray") else np.asarray(m)
        self.rated = {u: np.sort(cols[keep][np.nonzero(dense[i][keep] > 0)[0]]).astype(np.int32) for i, u in enumerate(users)}
        # the engine's exclusion lists are shared with the constraint filter (owned / wishlisted): append / replace the
        # rated lists through its owner when there is one
        if getattr(self, "constraint_filter", None) is not None:
            self.rated_list_index = self.constraint_filter.replace_rated_lists(self.rated)
        else:
            eng.set_exclusion_lists([self.rated[u] for u in users])
            self.rated_list_index = {u: i for i, u in enumerate(users)}

    def get_recommendations(self, user_id, top_k: int = 10,
                            valid_set_filter: Optional[List[str]] = None) -> List[RecommendationResult]:
        if self.svd_model is None:
            self.train_svd_model()
        if self.svd_model is None:
            logger.info("Using fallback recommendations due to insufficient data")
            return self._cold_start_recommendations(top_k, valid_set_filter)
        if getattr(self, "_engine_mismatch", None):
            raise RuntimeError("trained factors are not in the engine: " + self._engine_mismatch)
        if user_id not in self.user_lookup:
            logger.warning("User %s not found in user lookup", user_id)       # :428-430
            return self._cold_start_recommendations(top_k, valid_set_filter)
        ids, scores = self._search([user_id], top_k, _filter_bitmap(self.catalog, valid_set_filter))
        return [self.catalog.result(int(i), float(s), [self.CF_REASON]) for i, s in zip(ids[0], scores[0]) if i >= 0]

    def _search(self, user_ids: Sequence[Any], top_k: int, bitmap: Optional[np.ndarray]):
        rows = [self.user_lookup[u] for u in user_ids]
        q = self.user_factors[rows]
        n = len(self.catalog)
        allow = None
        if self.cf_universe is not None:
            allow = ScoringEngine.pack_bitmap(self.cf_universe)
        if bitmap is not None:
            if allow is None:
                allow = bitmap
            elif hasattr(bitmap, "is_cuda"):        # device-resident request filter AND the CF universe
                import torch
                u = torch.from_numpy(allow.view(np.int32)).to(bitmap.device)
                allow = torch.bitwise_and(bitmap.reshape(-1), u)
            else:
                allow = allow & bitmap
        cons = None
        if allow is not None:                       # the request's filter travels WITH the call (no handle state)
            cons = L.no_constraint()
            cons["allow_bitmap"] = 0
            allow = allow.reshape(1, -1)
        # the users' already-rated items (:441-451) are per-QUERY exclusion lists, not part of a constraint set
        excl = np.array([[self.rated_list_index.get(u, -1)] for u in user_ids], dtype=np.int32)
        scores, ids = self.engine.search(_clamp_k(top_k, self.engine), q_cf=q, w_content=0.0, w_cf=1.0,
                                         constraints=cons, allow_bitmaps=allow, excl_lists=excl, precision=self.precision)
        if hasattr(scores, "cpu"):
            scores, ids = scores.cpu().numpy(), ids.cpu().numpy()
        del n
        return ids, scores

    # -- cold-start fallbacks: the reference runs SQL aggregates (:485-592); same ordering on host arrays
    def _cold_start_recommendations(self, top_k: int, valid_set_filter=None) -> List[RecommendationResult]:
        c = self.catalog
        if c is None:
            return []
        valid = np.ones(len(c), dtype=bool)
        bm = _filter_bitmap(c, valid_set_filter)
        if bm is not None:
            valid = ScoringEngine.unpack_bitmap(bm, len(c))
        rs = self.ratings_summary
        if rs is not None:
            ok = valid & (rs["count"] >= 3)
            idx = np.nonzero(ok)[0]
            order = idx[np.lexsort((-rs["count"][idx], -rs["avg"][idx]))][:top_k]
            if len(order):
                return [c.result(int(i), float(rs["avg"][i]), [f"Popular set with {int(rs['count'][i])} ratings"])
                        for i in order]
        return self._get_recent_popular_sets(top_k, valid)

    def _get_recent_popular_sets(self, top_k: int, valid: Optional[np.ndarray] = None) -> List[RecommendationResult]:
        c = self.catalog
        ok = (c.year >= 2020) & (c.num_parts >= 100) & (c.num_parts <= 1000)
        if valid is not None:
            ok &= valid
        idx = np.nonzero(ok)[0]
        order = idx[np.lexsort((-c.num_parts[idx], -c.year[idx]))][:top_k]
        out = []
        for i in order:
            r = c.result(int(i), 0.8, ["Popular recent set"])
            r.theme_name = r.theme_name or "Unknown"
            out.append(r)
        return out


class HybridRecommender:
    """Constraints first, then content + CF, then the 0.4 / 0.6 combine -- same flow as the reference."""

    def __init__(self, dbcon=None, engine: Optional[ScoringEngine] = None,
                 content: Optional[ContentBasedRecommender] = None,
                 collaborative: Optional[CollaborativeFilteringRecommender] = None,
                 constraint_filter: Optional[HardConstraintFilter] = None):
        self.dbcon = dbcon
        self.engine = engine
        self.content_recommender = content
        self.collaborative_recommender = collaborative
        self.constraint_filter = constraint_filter
        self.content_weight = 0.4
        self.collaborative_weight = 0.6

    # ------------------------------------------------------------------ construction from arrays
    @classmethod
    def from_arrays(cls, catalog: CatalogTable, feat_matrix: np.ndarray, themes: Dict[int, str],
                    user_ids: Optional[Sequence[Any]] = None, user_factors: Optional[np.ndarray] = None,
                    item_factors: Optional[np.ndarray] = None, rated: Optional[Dict[Any, Sequence[int]]] = None,
                    owned: Optional[Dict[Any, Sequence[int]]] = None,
                    wishlisted: Optional[Dict[Any, Sequence[int]]] = None, cf_universe: Optional[np.ndarray] = None,
                    ratings_summary: Optional[dict] = None, device: int = 0, precision: str = "fp32",
                    k_max: int = 128, cf_dim: Optional[int] = None) -> "HybridRecommender":
        """``cf_dim``: factor width of the engine when no ``item_factors`` are given yet (``train_svd_model`` later
        loads them; the reference trains up to 50 components, recommendation_system.py:368)."""
        from .constraints import theme_bits_for
        n = len(catalog)
        feat = np.ascontiguousarray(feat_matrix, dtype=np.float32)
        f = (cf_dim or 0) if item_factors is None else int(item_factors.shape[1])
        if catalog.theme_id.max(initial=-1) > 32767 or catalog.year.max(initial=0) > 32767 or catalog.year.min(initial=0) < -32768:
            raise ValueError("theme_id / year do not fit the engine's int16 metadata columns")
        eng = ScoringEngine(n_items=n, dim=int(feat.shape[1]), cf_dim=f, device=device, k_max=k_max,
                            theme_bits=theme_bits_for(themes, catalog.theme_id))
        if item_factors is None and f > 0:
            item_factors = np.zeros((n, f), dtype=np.float32)
        theme16 = np.where(catalog.theme_id < 0, -1, catalog.theme_id).astype(np.int16)
        eng.load_catalog(emb=feat, item_factors=None if item_factors is None else np.asarray(item_factors, np.float32),
                         num_parts=np.clip(catalog.num_parts, -2**31, 2**31 - 1).astype(np.int32),
                         year=catalog.year.astype(np.int16), theme_id=theme16)
        filt = HardConstraintFilter(eng, catalog.set_nums, themes, owned=owned, wishlisted=wishlisted, rated=rated)
        rated_idx = {u: filt._excl_index[("rated", u)] for u in (rated or {})}
        content = ContentBasedRecommender(engine=eng, catalog=catalog, precision=precision)
        content.feat_matrix = feat
        cf = CollaborativeFilteringRecommender(engine=eng, catalog=catalog, user_ids=user_ids, user_factors=user_factors,
                                               rated=rated, rated_list_index=rated_idx, cf_universe=cf_universe,
                                               ratings_summary=ratings_summary, precision=precision)
        cf.constraint_filter = filt            # owner of the engine's exclusion lists (re-training replaces the rated ones)
        return cls(engine=eng, content=content, collaborative=cf, constraint_filter=filt)

    # ------------------------------------------------------------------ reference surface
    def get_recommendations(self, user_id: Optional[int] = None, liked_set: Optional[str] = None, top_k: int = 10,
                            constraints: Optional[List[HardConstraint]] = None
                            ) -> Tuple[List[RecommendationResult], ConstraintResult]:
        logger.info("Getting recommendations with %d hard constraints", len(constraints or []))
        constraint_result = None
        valid_set_nums = None
        if constraints:
            constraint_result = self.constraint_filter.apply_constraints(constraints)
            valid_set_nums = constraint_result.valid_set_nums
            if not valid_set_nums:
                logger.warning("No sets meet the specified hard constraints")          # :636-638
                return [], constraint_result
        content_recs: List[RecommendationResult] = []
        collaborative_recs: List[RecommendationResult] = []
        # the validity bitmap the mask kernel just wrote stays on the device for both searches
        valid_filter = valid_set_nums
        if constraint_result is not None and getattr(constraint_result, "device_bitmap", None) is not None:
            valid_filter = DeviceFilter(constraint_result.device_bitmap, len(valid_set_nums))
        if liked_set:
            content_recs = self.content_recommender.get_similar_sets(liked_set, top_k * 2,
                                                                    valid_set_filter=valid_filter)
        if user_id:
            collaborative_recs = self.collaborative_recommender.get_recommendations(user_id, top_k * 2,
                                                                                   valid_set_filter=valid_filter)
        if not content_recs and collaborative_recs:
            final_recs = collaborative_recs[:top_k]
        elif content_recs and not collaborative_recs:
            final_recs = content_recs[:top_k]
        elif not content_recs and not collaborative_recs:
            final_recs = self._get_constrained_popular_sets(top_k, valid_set_nums)
        else:
            final_recs = self._combine_recommendations(content_recs, collaborative_recs, top_k)
        if constraint_result and constraint_result.violations:
            for rec in final_recs:
                rec.constraint_violations = [v.message for v in constraint_result.violations]
        return final_recs, constraint_result

    def get_recommendations_from_request(self, request: RecommendationRequest
                                         ) -> Tuple[List[RecommendationResult], ConstraintResult]:
        try:
            constraints = self.constraint_filter.create_constraint_set(
                price_max=request.price_max, price_min=request.price_min, pieces_max=request.pieces_max,
                pieces_min=request.pieces_min, age_min=request.age_min, age_max=request.age_max,
                year_min=request.year_min, year_max=request.year_max, required_themes=request.required_themes,
                excluded_themes=request.excluded_themes, max_complexity=request.max_complexity,
                min_complexity=request.min_complexity, must_be_available=request.must_be_available,
                exclude_owned=request.exclude_owned, exclude_wishlisted=request.exclude_wishlisted,
                user_id=request.user_id)
        except Exception as e:                                                        # :702-714
            logger.error("Failed to create constraints: %s", e)
            return [], ConstraintResult(valid_set_nums=[], violations=[], applied_constraints=[],
                                        performance_stats={})
        return self.get_recommendations(user_id=request.user_id, liked_set=request.liked_set, top_k=request.top_k,
                                        constraints=constraints)

    def _combine_recommendations(self, content_recs: List[RecommendationResult],
                                 collaborative_recs: List[RecommendationResult],
                                 top_k: int) -> List[RecommendationResult]:
        cat = self.content_recommender.catalog
        c_by = {r.set_num: r for r in content_recs}
        f_by = {r.set_num: r for r in collaborative_recs}
        c_ids = np.array([[cat.index_of[s] for s in c_by]], dtype=np.int32).reshape(1, -1)
        f_ids = np.array([[cat.index_of[s] for s in f_by]], dtype=np.int32).reshape(1, -1)
        c_sc = np.array([[r.score for r in c_by.values()]], dtype=np.float32).reshape(1, -1)
        f_sc = np.array([[r.score for r in f_by.values()]], dtype=np.float32).reshape(1, -1)
        k = min(top_k, c_ids.shape[1] + f_ids.shape[1])
        scores, ids = self.engine.hybrid_combine(c_ids, c_sc, f_ids, f_sc,
                                                 [self.content_weight, self.collaborative_weight], k)
        out = []
        for i, s in zip(ids[0], scores[0]):
            if i < 0:
                break
            sn = cat.set_nums[int(i)]
            c_rec, f_rec = c_by.get(sn), f_by.get(sn)
            base = c_rec if c_rec else f_rec
            reasons = ([f"Content: {r}" for r in c_rec.reasons] if c_rec else []) + \
                      ([f"Community: {r}" for r in f_rec.reasons] if f_rec else [])
            out.append(RecommendationResult(set_num=sn, name=base.name, score=float(s), reasons=reasons,
                                            theme_name=base.theme_name, year=base.year, num_parts=base.num_parts,
                                            img_url=base.img_url))
        return out

    def set_weights(self, content_weight: float, collaborative_weight: float):
        total = content_weight + collaborative_weight                                  # :845-849
        self.content_weight = content_weight / total
        self.collaborative_weight = collaborative_weight / total
        logger.info("Updated weights - Content: %.2f, Collaborative: %.2f", self.content_weight,
                    self.collaborative_weight)

    def _get_constrained_popular_sets(self, top_k: int,
                                      valid_set_nums: Optional[List[str]] = None) -> List[RecommendationResult]:
        """Host restatement of the popular-sets SQL ordering (:723-787): rated >= 3 by average first, then
        rating count, then newest."""
        c = self.content_recommender.catalog
        ok = c.num_parts > 0
        if valid_set_nums:
            m = np.zeros(len(c), dtype=bool)
            m[[c.index_of[s] for s in valid_set_nums if s in c.index_of]] = True
            ok &= m
        idx = np.nonzero(ok)[0]
        rs = self.collaborative_recommender.ratings_summary if self.collaborative_recommender else None
        cnt = rs["count"][idx] if rs is not None else np.zeros(len(idx), dtype=np.int64)
        avg = rs["avg"][idx] if rs is not None else np.zeros(len(idx))
        key1 = np.where(cnt >= 3, avg, 0.0)
        order = np.lexsort((-c.year[idx], -cnt, -key1))[:top_k]
        out = []
        for j in order:
            i = int(idx[j])
            reasons = ["Popular choice"]
            if cnt[j] > 0:
                reasons.append(f"Avg rating: {avg[j]:.1f} ({int(cnt[j])} reviews)")
            else:
                reasons.append("Recent release")
            out.append(c.result(i, float(avg[j]) if cnt[j] > 0 else 0.5, reasons))
        return out

    # ------------------------------------------------------------------ B200-native batched path (north star)
    def recommend_batch(self, user_ids: Optional[Sequence[Any]], liked_sets: Optional[Sequence[str]], top_k: int,
                        constraints: Optional[Sequence[Sequence[HardConstraint]]] = None,
                        precision: str = "bf16_rescore", exclude_rated: bool = True):
        """Fused full-catalog hybrid for a batch: one ``bb_search`` call scores
        w_c*cos + w_f*(u.v) against every row, applies each query's constraint
        descriptor inside the epilogue and returns (scores[B,k], ids[B,k])."""
        cf = self.collaborative_recommender
        cat = self.content_recommender.catalog
        b = len(user_ids) if user_ids is not None else len(liked_sets)
        q_item = self_id = q_cf = None
        if liked_sets is not None:
            q_item = np.array([cat.index_of[s] for s in liked_sets], dtype=np.int32) + np.int32(self.engine.id_offset)
            self_id = q_item
        if user_ids is not None:
            q_cf = cf.user_factors[[cf.user_lookup[u] for u in user_ids]]
        descs = []
        for j in range(b):
            cons_j = constraints[j] if constraints is not None else None
            descs.append(self.constraint_filter.descriptor(cons_j) if cons_j is not None else L.no_constraint())
        # the users' already-rated items are per-QUERY exclusion lists: they do not make the constraint sets distinct
        excl = None
        if exclude_rated and user_ids is not None:
            excl = np.array([[cf.rated_list_index.get(u, -1)] for u in user_ids], dtype=np.int32)
        # queries that share a constraint set (popular filters) share one validity bitmap on the device
        distinct, index = ScoringEngine.dedupe_constraints(np.stack(descs))
        return self.engine.search(top_k, q_item=q_item, q_cf=q_cf,
                                  w_content=self.content_weight if q_item is not None else 0.0,
                                  w_cf=self.collaborative_weight if q_cf is not None else 0.0,
                                  constraints=distinct, constraint_index=index, self_id=self_id, excl_lists=excl,
                                  precision=precision)


def accelerate(reference_hybrid, themes: Dict[int, str], owned=None, wishlisted=None, device: int = 0,
               precision: str = "fp32") -> HybridRecommender:
    """Builds a GPU-backed ``HybridRecommender`` from a LIVE reference
    ``HybridRecommender`` whose ``prepare_features()`` / ``train_svd_model()``
    already ran: adopts ``set_feat`` / ``feat_matrix`` (recommendation_system.py:137-191)
    and the SVD factors / pivot table (:325-337, :402-407).  The returned object
    has the same methods, so callers (recommendation_api.py:456-468, :544) swap it in."""
    content = reference_hybrid.content_recommender
    cfr = reference_hybrid.collaborative_recommender
    catalog = CatalogTable.from_frame(content.set_feat)
    n = len(catalog)
    user_ids = user_factors = item_factors = rated = universe = summary = None
    if getattr(cfr, "svd_model", None) is not None and getattr(cfr, "item_factors", None) is not None:
        pivot = cfr.user_item_matrix
        cols = [catalog.index_of.get(s, -1) for s in pivot.columns]
        keep = np.array([c >= 0 for c in cols])
        col_idx = np.array([c for c in cols if c >= 0], dtype=np.int64)
        item_factors = np.zeros((n, cfr.item_factors.shape[1]), dtype=np.float32)
        item_factors[col_idx] = np.asarray(cfr.item_factors, dtype=np.float32)[keep]
        universe = np.zeros(n, dtype=bool)
        universe[col_idx] = True
        user_ids = list(pivot.index)
        user_factors = np.asarray(cfr.user_factors, dtype=np.float32)
        r = pivot.to_numpy()[:, keep]
        rated = {u: np.sort(col_idx[np.nonzero(r[i] > 0)[0]]).astype(np.int32) for i, u in enumerate(user_ids)}
        cnt = np.zeros(n, dtype=np.int64)
        tot = np.zeros(n)
        cnt[col_idx] = (r > 0).sum(0)
        tot[col_idx] = r.sum(0)
        summary = {"count": cnt, "avg": np.divide(tot, np.maximum(cnt, 1))}
    hyb = HybridRecommender.from_arrays(catalog, np.asarray(content.feat_matrix), themes, user_ids=user_ids,
                                        user_factors=user_factors, item_factors=item_factors, rated=rated, owned=owned,
                                        wishlisted=wishlisted, cf_universe=universe, ratings_summary=summary,
                                        device=device, precision=precision)
    hyb.content_weight = reference_hybrid.content_weight
    hyb.collaborative_weight = reference_hybrid.collaborative_weight
    hyb.dbcon = reference_hybrid.dbcon
    return hyb
