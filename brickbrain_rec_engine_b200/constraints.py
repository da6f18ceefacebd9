"""Hard constraints: the reference's public types and ``HardConstraintFilter``
surface (src/scripts/hard_constraint_filter.py), re-implemented over the GPU mask
kernel instead of SQL.

Same names, argument meaning and error behaviour as the reference:
``ConstraintType`` / ``ConstraintSeverity`` / ``HardConstraint`` /
``ConstraintViolation`` / ``ConstraintResult`` (hard_constraint_filter.py:14-66),
``create_constraint_set(**kwargs)`` (:98-261) and
``apply_constraints(constraints, candidate_set_nums=None)`` (:263-316).  The SQL
fragments of ``_constraint_to_sql`` (:366-480) are folded on the host into one
integer descriptor (``bb_constraint``) and evaluated by ``bb_apply_constraints``
/ inside the scan epilogue; nothing here evaluates a predicate on the CPU.
"""
from __future__ import annotations

import logging
import re
import time
from dataclasses import dataclass, field
from datetime import datetime
from enum import Enum
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L

logger = logging.getLogger(__name__)


class ConstraintType(Enum):
    PRICE_MAX = "price_max"
    PRICE_MIN = "price_min"
    PIECES_MAX = "pieces_max"
    PIECES_MIN = "pieces_min"
    AGE_MIN = "age_min"
    AGE_MAX = "age_max"
    YEAR_MIN = "year_min"
    YEAR_MAX = "year_max"
    THEMES_REQUIRED = "themes_required"
    THEMES_EXCLUDED = "themes_excluded"
    COMPLEXITY_MAX = "complexity_max"
    COMPLEXITY_MIN = "complexity_min"
    AVAILABILITY = "availability"
    EXCLUDE_OWNED = "exclude_owned"
    EXCLUDE_WISHLISTED = "exclude_wishlisted"


class ConstraintSeverity(Enum):
    BLOCKING = "blocking"
    WARNING = "warning"
    INFO = "info"


@dataclass
class HardConstraint:
    constraint_type: ConstraintType
    value: Any
    severity: ConstraintSeverity = ConstraintSeverity.BLOCKING
    description: str = ""

    def __post_init__(self):
        if not self.description:
            self.description = f"{self.constraint_type.value}: {self.value}"


@dataclass
class ConstraintViolation:
    constraint: HardConstraint
    violating_count: int
    total_count: int
    message: str
    suggested_alternatives: List[str] = None


@dataclass
class ConstraintResult:
    valid_set_nums: List[str]
    violations: List[ConstraintViolation]
    applied_constraints: List[HardConstraint]
    performance_stats: Dict[str, Any]
    constraint_sql: str = ""
    # (not in the reference) the validity bitmap on the GPU, reused by the searches of the same request
    device_bitmap: Any = field(default=None, repr=False, compare=False)


# kwarg -> (type, description template); order = emission order of create_constraint_set (:140-259)
_KWARG_TABLE = [
    ("price_max", ConstraintType.PRICE_MAX, "Must cost less than ${v:.2f}"),
    ("price_min", ConstraintType.PRICE_MIN, "Must cost more than ${v:.2f}"),
    ("pieces_max", ConstraintType.PIECES_MAX, "Must have fewer than {v} pieces"),
    ("pieces_min", ConstraintType.PIECES_MIN, "Must have more than {v} pieces"),
    ("age_min", ConstraintType.AGE_MIN, "Must be suitable for ages {v}+"),
    ("age_max", ConstraintType.AGE_MAX, "Must be suitable for ages up to {v}"),
    ("year_min", ConstraintType.YEAR_MIN, "Must be released after {v}"),
    ("year_max", ConstraintType.YEAR_MAX, "Must be released before {v}"),
]
_COMPLEXITY_MAX = {"simple": 200, "moderate": 800, "complex": 999999}     # :453-461
_COMPLEXITY_MIN = {"simple": 0, "moderate": 200, "complex": 800}          # :463-471


def _like_regex(name: str) -> "re.Pattern[str]":
    """``LOWER(name) LIKE LOWER('%<name>%')`` (hard_constraint_filter.py:505-512)."""
    body = "".join(".*" if ch == "%" else "." if ch == "_" else re.escape(ch) for ch in name)
    return re.compile(".*" + body + ".*", re.IGNORECASE | re.DOTALL)


@dataclass
class CompiledConstraint:
    """One folded descriptor + the theme bitmap it needs (if any)."""
    desc: np.ndarray                       # 0-d array of CONSTRAINT_DTYPE (theme_set / excl indices unresolved: -1)
    theme_bitmap: Optional[np.ndarray]     # uint32[theme_words] allowed theme ids, or None
    exclude_owned_user: Optional[Any] = None
    exclude_wishlisted_user: Optional[Any] = None


def theme_bits_for(themes: Dict[int, str], theme_ids=None) -> int:
    """Bitmap width (multiple of 32, >= 1024) that holds every theme id of the ``themes`` table and of the catalog's
    ``theme_id`` column -- pass it as ``ScoringEngine(theme_bits=...)``."""
    top = max([int(t) for t in themes] + [0])
    if theme_ids is not None and len(theme_ids):
        top = max(top, int(np.max(theme_ids)))
    return max(1024, (top + 32) // 32 * 32)


class ConstraintCompiler:
    """Host-side folding of a ``HardConstraint`` list into ``bb_constraint`` fields.

    ``themes``: {theme_id: name} (the reference's ``themes`` table); name lookup is
    the reference's case-insensitive substring match with a per-name-set cache
    (hard_constraint_filter.py:482-532)."""

    def __init__(self, themes: Dict[int, str], theme_bits: int = 1024, now_year: Optional[int] = None):
        self.themes = dict(themes)
        self.theme_bits = theme_bits
        self.theme_words = theme_bits // 32
        self.now_year = now_year
        self._theme_cache: Dict[tuple, List[int]] = {}

    def get_theme_ids(self, theme_names: Sequence[str]) -> List[int]:
        key = tuple(sorted(theme_names))
        if key not in self._theme_cache:
            pats = [_like_regex(n) for n in theme_names]
            self._theme_cache[key] = sorted(t for t, name in self.themes.items() if any(p.fullmatch(name) for p in pats))
        return self._theme_cache[key]

    def _bitmap(self, ids) -> np.ndarray:
        bm = np.zeros(self.theme_words, dtype=np.uint32)
        for t in ids:
            if not 0 <= t < self.theme_bits:
                # the reference's `theme_id = ANY(ids)` would match such a theme; a bitmap that cannot hold it would
                # silently change the result -- size theme_bits from the themes table instead (theme_bits_for)
                raise ValueError(f"theme id {t} does not fit the engine's theme bitmaps ({self.theme_bits} bits): "
                                 "create the engine with theme_bits >= max theme id + 1")
            bm[t >> 5] |= np.uint32(1 << (t & 31))
        return bm

    def compile(self, constraints: Sequence[HardConstraint]) -> CompiledConstraint:
        d = L.no_constraint()
        pmin, pmax = 1, L.INT32_MAX                      # "s.num_parts > 0" is always present (:343)
        ymin, ymax = -32768, 32767
        required: Optional[set] = None
        excluded: set = set()
        theme_active = False
        owned = wish = None
        flags = 0
        for c in constraints:
            t, v = c.constraint_type, c.value
            if t is ConstraintType.PIECES_MAX:
                pmax = min(pmax, int(v))
            elif t is ConstraintType.PIECES_MIN:
                pmin = max(pmin, int(v))
            elif t is ConstraintType.YEAR_MIN:
                ymin = max(ymin, int(v))
            elif t is ConstraintType.YEAR_MAX:
                ymax = min(ymax, int(v))
            elif t is ConstraintType.PRICE_MAX:
                pmax = min(pmax, int(v / 0.10))          # :391-395 (double rounding kept)
            elif t is ConstraintType.PRICE_MIN:
                pmin = max(pmin, int(v / 0.15))          # :397-400
            elif t is ConstraintType.THEMES_REQUIRED:
                ids = self.get_theme_ids(v)
                if ids:
                    required = set(ids) if required is None else required & set(ids)
                    theme_active = True
                else:
                    flags |= L.BB_C_ALWAYS_FALSE         # "1=0" (:408-409)
            elif t is ConstraintType.THEMES_EXCLUDED:
                ids = self.get_theme_ids(v)
                if ids:
                    excluded |= set(ids)
                    theme_active = True
            elif t is ConstraintType.AGE_MIN:
                if v <= 4:
                    pmax = min(pmax, 50)
                elif v <= 8:
                    pmax = min(pmax, 500)
                elif v <= 12:
                    pmax = min(pmax, 1500)
                else:
                    pmin = max(pmin, 500)
            elif t is ConstraintType.AGE_MAX:
                if v <= 8:
                    pmax = min(pmax, 300)
                elif v <= 12:
                    pmax = min(pmax, 800)
            elif t is ConstraintType.EXCLUDE_OWNED:
                owned = v
            elif t is ConstraintType.EXCLUDE_WISHLISTED:
                wish = v
            elif t is ConstraintType.COMPLEXITY_MAX:
                pmax = min(pmax, _COMPLEXITY_MAX.get(v, 999999))
            elif t is ConstraintType.COMPLEXITY_MIN:
                pmin = max(pmin, _COMPLEXITY_MIN.get(v, 0))
            elif t is ConstraintType.AVAILABILITY:
                now = self.now_year if self.now_year is not None else datetime.now().year
                ymin = max(ymin, now - 5)                # :473-477
        d["parts_min"] = max(L.INT32_MIN, min(L.INT32_MAX, pmin))
        d["parts_max"] = max(L.INT32_MIN, min(L.INT32_MAX, pmax))
        d["year_min"] = max(-32768, min(32767, ymin))
        d["year_max"] = max(-32768, min(32767, ymax))
        d["flags"] = flags
        bitmap = None
        if theme_active:
            base = set(range(self.theme_bits)) if required is None else required
            bitmap = self._bitmap(base - excluded)
        return CompiledConstraint(d, bitmap, owned, wish)


class HardConstraintFilter:
    """Drop-in for the reference's ``HardConstraintFilter`` over a ``ScoringEngine``.

    ``catalog`` supplies what the SQL joined against: ``set_nums`` (catalog order =
    ``ORDER BY set_num``), ``themes`` {id: name}, and optional ``owned`` /
    ``wishlisted`` {user_id: item ids} (user_collections / user_wishlists)."""

    def __init__(self, engine, set_nums: Sequence[str], themes: Dict[int, str],
                 owned: Optional[Dict[Any, Sequence[int]]] = None,
                 wishlisted: Optional[Dict[Any, Sequence[int]]] = None, dbcon=None,
                 rated: Optional[Dict[Any, Sequence[int]]] = None):
        self.dbcon = dbcon
        self.engine = engine
        self.set_nums = list(set_nums)
        self.index_of = {s: i for i, s in enumerate(self.set_nums)}
        self.compiler = ConstraintCompiler(themes, theme_bits=engine.theme_words * 32)
        too_big = [t for t in themes if not 0 <= int(t) < engine.theme_words * 32]
        if too_big:
            raise ValueError(f"theme ids {sorted(too_big)[:5]}... do not fit the engine's theme bitmaps "
                             f"({engine.theme_words * 32} bits): create the engine with theme_bits=theme_bits_for(themes)")
        self._theme_cache = self.compiler._theme_cache
        self._constraint_cache: Dict = {}
        self._set_metadata_cache: Dict = {}
        self.performance_stats = {"total_constraints_applied": 0, "total_sets_filtered": 0,
                                  "average_filter_time_ms": 0, "constraint_hit_rates": {}}
        # exclusion lists live on the device: [owned..., wishlisted..., rated...]
        self._excl_index: Dict[tuple, int] = {}
        lists = []
        for kind, table in (("owned", owned), ("wish", wishlisted), ("rated", rated)):
            for user, ids in (table or {}).items():
                self._excl_index[(kind, user)] = len(lists)
                lists.append(np.asarray(ids, dtype=np.int32))
        self._excl_lists = lists
        if lists:
            engine.set_exclusion_lists(lists)
        self._theme_bitmaps: List[np.ndarray] = []
        self._theme_bitmap_index: Dict[bytes, int] = {}

    # -- reference surface ---------------------------------------------------------------------
    def create_constraint_set(self, price_max=None, price_min=None, pieces_max=None, pieces_min=None, age_min=None,
                              age_max=None, year_min=None, year_max=None, required_themes=None, excluded_themes=None,
                              max_complexity=None, min_complexity=None, must_be_available: bool = False,
                              exclude_owned: bool = False, exclude_wishlisted: bool = False,
                              user_id=None) -> List[HardConstraint]:
        kw = dict(price_max=price_max, price_min=price_min, pieces_max=pieces_max, pieces_min=pieces_min,
                  age_min=age_min, age_max=age_max, year_min=year_min, year_max=year_max)
        out = [HardConstraint(t, kw[name], description=fmt.format(v=kw[name]))
               for name, t, fmt in _KWARG_TABLE if kw[name] is not None]
        if required_themes:
            out.append(HardConstraint(ConstraintType.THEMES_REQUIRED, required_themes,
                                      description=f"Must be from themes: {', '.join(required_themes)}"))
        if excluded_themes:
            out.append(HardConstraint(ConstraintType.THEMES_EXCLUDED, excluded_themes,
                                      description=f"Must NOT be from themes: {', '.join(excluded_themes)}"))
        if max_complexity:
            out.append(HardConstraint(ConstraintType.COMPLEXITY_MAX, max_complexity,
                                      description=f"Must be {max_complexity} complexity or simpler"))
        if min_complexity:
            out.append(HardConstraint(ConstraintType.COMPLEXITY_MIN, min_complexity,
                                      description=f"Must be {min_complexity} complexity or more complex"))
        for flag, name, t, text in ((exclude_owned, "exclude_owned", ConstraintType.EXCLUDE_OWNED,
                                     "Must not be in user's collection"),
                                    (exclude_wishlisted, "exclude_wishlisted", ConstraintType.EXCLUDE_WISHLISTED,
                                     "Must not be in user's wishlist")):
            if flag:
                if not user_id:
                    logger.warning("%s requires user_id, constraint will be ignored", name)      # :233-251
                else:
                    out.append(HardConstraint(t, user_id, description=text))
        if must_be_available:
            out.append(HardConstraint(ConstraintType.AVAILABILITY, True,
                                      description="Must be currently available for purchase"))
        return out

    def replace_rated_lists(self, rated: Dict[Any, Sequence[int]]) -> Dict[Any, int]:
        """Re-uploads the exclusion lists with new already-rated lists (after a factor re-training): owned / wishlisted
        lists keep their indices, rated lists follow.  Returns {user: list index}."""
        keep = [(key, idx) for key, idx in self._excl_index.items() if key[0] != "rated"]
        keep.sort(key=lambda kv: kv[1])
        lists = [self._excl_lists[idx] for _, idx in keep]
        self._excl_index = {key: i for i, (key, _) in enumerate(keep)}
        out = {}
        for user, ids in rated.items():
            self._excl_index[("rated", user)] = out[user] = len(lists)
            lists.append(np.asarray(ids, dtype=np.int32))
        self._excl_lists = lists
        self.engine.set_exclusion_lists(lists)
        return out

    def _theme_set_index(self, bitmap: np.ndarray) -> int:
        key = bitmap.tobytes()
        if key not in self._theme_bitmap_index:
            self._theme_bitmap_index[key] = len(self._theme_bitmaps)
            self._theme_bitmaps.append(bitmap)
            self.engine.set_theme_sets(np.stack(self._theme_bitmaps))
        return self._theme_bitmap_index[key]

    def descriptor(self, constraints: Sequence[HardConstraint], rated_user=None) -> np.ndarray:
        """``bb_constraint`` for a constraint list, with device table indices resolved."""
        cc = self.compiler.compile(constraints)
        d = cc.desc.copy()
        if cc.theme_bitmap is not None:
            d["theme_set"] = self._theme_set_index(cc.theme_bitmap)
        slots = []
        for kind, user in (("owned", cc.exclude_owned_user), ("wish", cc.exclude_wishlisted_user),
                           ("rated", rated_user)):
            if user is not None and (kind, user) in self._excl_index:
                slots.append(self._excl_index[(kind, user)])
        ex = d["excl_list"]
        ex[: len(slots)] = slots
        d["excl_list"] = ex
        return d

    def apply_constraints(self, constraints: List[HardConstraint],
                          candidate_set_nums: Optional[List[str]] = None) -> ConstraintResult:
        t0 = time.perf_counter()
        logger.info("Applying %d hard constraints...", len(constraints))
        try:
            descs = [self.descriptor(constraints)] + [self.descriptor([c]) for c in constraints]
            dev_bitmaps, counts = self.engine.apply_constraints(np.stack(descs), device_out=True)
            if hasattr(dev_bitmaps, "cpu"):
                bitmaps = dev_bitmaps.cpu().numpy().view(np.uint32)
            else:                                  # (CPU test doubles return host arrays)
                bitmaps, dev_bitmaps = dev_bitmaps, None
            n = len(self.set_nums)
            cand = None
            if candidate_set_nums:
                cand = np.zeros(n, dtype=bool)
                cand[[self.index_of[s] for s in candidate_set_nums if s in self.index_of]] = True
            valid = self.engine.unpack_bitmap(bitmaps[0], n)
            if cand is not None:
                valid &= cand
            valid_set_nums = [self.set_nums[i] for i in np.nonzero(valid)[0]]
            # violation analysis: each constraint alone vs "num_parts > 0" only (:534-587)
            base_desc = self.descriptor([])
            base_bm, _ = self.engine.apply_constraints(base_desc)
            base = self.engine.unpack_bitmap(base_bm[0], n)
            if cand is not None:
                base &= cand
            total = int(base.sum())
            violations = []
            for j, c in enumerate(constraints):
                alone = self.engine.unpack_bitmap(bitmaps[1 + j], n)
                if cand is not None:
                    alone &= cand
                remaining = int(alone.sum())
                rate = (total - remaining) / total if total else 0.0
                if rate > 0.8:
                    violations.append(ConstraintViolation(
                        constraint=c, violating_count=total - remaining, total_count=total,
                        message=f"Constraint '{c.description}' eliminated {rate:.1%} of sets",
                        suggested_alternatives=_alternatives(c)))
            ms = (time.perf_counter() - t0) * 1000.0
            self._update_performance_stats(len(constraints), len(valid_set_nums), ms)
            res = ConstraintResult(
                valid_set_nums=valid_set_nums, violations=violations, applied_constraints=constraints,
                performance_stats={"filter_time_ms": ms,
                                   "input_sets": len(candidate_set_nums) if candidate_set_nums else "all",
                                   "output_sets": len(valid_set_nums), "constraint_count": len(constraints)},
                constraint_sql="")
            # the validity bitmap stays on the device for the searches that follow (HybridRecommender); with a
            # candidate pre-filter the host-side AND above is the truth, so nothing is attached
            res.device_bitmap = dev_bitmaps[0:1] if (dev_bitmaps is not None and cand is None) else None
            return res
        except Exception as e:  # same contract as the reference: log and re-raise (:314-316)
            logger.error("Error applying hard constraints: %s", e)
            raise

    def _update_performance_stats(self, constraint_count: int, result_count: int, filter_time: float):
        st = self.performance_stats
        st["total_constraints_applied"] += constraint_count
        st["total_sets_filtered"] += result_count
        ops = st.get("operation_count", 0) + 1
        st["average_filter_time_ms"] = (st["average_filter_time_ms"] * (ops - 1) + filter_time) / ops
        st["operation_count"] = ops

    def get_performance_report(self) -> Dict[str, Any]:
        return self.performance_stats.copy()

    def clear_cache(self):
        self._constraint_cache.clear()
        self._theme_cache.clear()
        self._set_metadata_cache.clear()


def _alternatives(c: HardConstraint) -> List[str]:
    """Suggestion texts for an over-restrictive constraint (hard_constraint_filter.py:589-639)."""
    t, v = c.constraint_type, c.value
    if t is ConstraintType.PIECES_MAX:
        return [f"Try increasing to {int(v * 1.5)} pieces", f"Consider {int(v * 2)} pieces for more options",
                "Remove piece count limit and use price instead"]
    if t is ConstraintType.PIECES_MIN:
        return [f"Try decreasing to {int(v * 0.7)} pieces", f"Consider {int(v * 0.5)} pieces for more options",
                "Remove minimum piece requirement"]
    if t is ConstraintType.PRICE_MAX:
        return [f"Try increasing budget to ${v * 1.3:.2f}", f"Consider ${v * 1.5:.2f} for more premium options",
                "Look for sales or discounted sets"]
    if t is ConstraintType.THEMES_REQUIRED:
        return ["Try broader theme categories (e.g., 'space' instead of 'Star Wars')", "Consider related themes",
                "Remove theme restriction and browse by interest category"]
    if t is ConstraintType.AGE_MIN:
        return [f"Try age {v - 2}+ for more options", "Consider that age ratings are conservative",
                "Look at similar complexity levels across age ranges"]
    return []


# module-level helpers of the reference (hard_constraint_filter.py:664-740)
def create_budget_constraints(budget_max: float, budget_min: float = None) -> List[HardConstraint]:
    out = []
    if budget_max is not None:
        out.append(HardConstraint(ConstraintType.PRICE_MAX, budget_max, description=f"Budget limit: ${budget_max:.2f}"))
    if budget_min is not None:
        out.append(HardConstraint(ConstraintType.PRICE_MIN, budget_min, description=f"Budget floor: ${budget_min:.2f}"))
    return out


def create_age_appropriate_constraints(age: int, strict: bool = True) -> List[HardConstraint]:
    slack, label = (2, "strict") if strict else (4, "flexible")
    return [HardConstraint(ConstraintType.AGE_MIN, max(4, age - slack),
                           description=f"Age appropriate for {age} year old ({label})")]


_SIZE_RANGES = {"mini": (1, 50), "small": (51, 200), "medium": (201, 800), "large": (801, 2000), "xl": (2001, 10000)}


def create_size_constraints(size_category: str) -> List[HardConstraint]:
    rng = _SIZE_RANGES.get(size_category.lower())
    if rng is None:
        return []
    return [HardConstraint(ConstraintType.PIECES_MIN, rng[0], description=f"{size_category.title()} size minimum"),
            HardConstraint(ConstraintType.PIECES_MAX, rng[1], description=f"{size_category.title()} size maximum")]
