"""Seeded synthetic inputs for the scoring path.

TEST / BENCH INFRASTRUCTURE.  Pure NumPy, no reference code.  The constraint
scenarios restate the *parameter sets* exercised by the reference's own tests
(tests/unit/test_recommendations.py:393-472, 511-541, 600-653 and
examples/hard_constraint_demo.py:278-292) as plain kwargs / (type, value) pairs.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

THEME_WORDS = [
    "Star Wars", "City", "Creator", "Technic", "Friends", "Duplo", "Architecture", "Space",
    "Creator Expert", "Ninjago", "Harry Potter", "Castle", "Pirates", "Trains", "Baby",
    "Adult Welcome", "Space Police", "Classic Space", "Town", "Racers", "Bionicle", "Minecraft",
    "Super Heroes", "Ideas", "Speed Champions", "Disney", "Jurassic World", "Art", "Icons", "Dots",
]


def themes_table(n_themes: int = 411, max_id: int = 777, seed: int = 7) -> Dict[int, str]:
    """411 theme ids scattered in [1, 777] (ids reach 777 in
    tests/unit/test_recommendations.py:54).  Names reuse the words the reference
    scenarios query by substring, with sub-theme suffixes so LIKE '%City%' hits
    several ids."""
    rng = np.random.default_rng(seed)
    ids = np.sort(rng.choice(np.arange(1, max_id + 1), size=n_themes, replace=False))
    out = {}
    for j, tid in enumerate(ids):
        base = THEME_WORDS[j % len(THEME_WORDS)]
        out[int(tid)] = base if j < len(THEME_WORDS) else f"{base} Series {j // len(THEME_WORDS)}"
    return out


def embeddings(n: int, d: int, seed: int, kind: str = "iso") -> np.ndarray:
    """Unit-norm fp32 rows.  ``iso``: N(0,1) normalized.  ``aniso``:
    0.8*mu + 0.6*g_hat re-normalized (expected pairwise cosine ~0.64, mimicking
    test_embeddings/index.faiss whose mean pairwise cosine is 0.66)."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, d), dtype=np.float32)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    if kind == "iso":
        return g
    mu = np.random.default_rng(10_000 + d).standard_normal(d).astype(np.float32)
    mu /= np.linalg.norm(mu)
    x = 0.8 * mu[None, :] + 0.6 * g
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


def cf_factors(n: int, f: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, f), dtype=np.float32) / np.float32(np.sqrt(f))).astype(np.float32)


def metadata(n: int, seed: int, themes: Dict[int, str], null_theme_frac: float = 0.01
             ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(num_parts int32 log-uniform [1,12000], year int16 1950..2025, theme_id int16 with -1 = NULL)."""
    rng = np.random.default_rng(seed)
    num_parts = np.exp(rng.uniform(0.0, np.log(12000.0), size=n)).astype(np.int32)
    num_parts = np.clip(num_parts, 1, 12000).astype(np.int32)
    year = rng.integers(1950, 2026, size=n).astype(np.int16)
    tids = np.array(sorted(themes), dtype=np.int16)
    theme_id = tids[rng.integers(0, len(tids), size=n)].astype(np.int16)
    theme_id[rng.random(n) < null_theme_frac] = -1
    return num_parts, year, theme_id


def set_nums(n: int) -> List[str]:
    """Zero-padded so lexicographic order == catalog index order (the reference
    orders by set_num, recommendation_system.py:117, hard_constraint_filter.py:359)."""
    return [f"{i:08d}-1" for i in range(n)]


def user_lists(n_users: int, n_items: int, mean_len: int, seed: int) -> Dict[int, np.ndarray]:
    """Sorted unique item-id lists per user (owned / wishlisted / rated)."""
    rng = np.random.default_rng(seed)
    out = {}
    for u in range(n_users):
        m = int(rng.integers(max(1, mean_len // 2), mean_len * 3 // 2 + 1))
        out[u + 1] = np.sort(rng.choice(n_items, size=min(m, n_items), replace=False)).astype(np.int32)
    return out


# (name, kwargs for HardConstraintFilter.create_constraint_set) -- reference scenarios
SCENARIO_KWARGS: List[Tuple[str, dict]] = [
    ("multi", dict(price_max=150.0, pieces_min=200, pieces_max=1000, age_min=8,
                   required_themes=["City", "Creator"], year_min=2018)),
    ("heavy", dict(price_max=200.0, price_min=30.0, pieces_min=200, pieces_max=1500, age_min=8,
                   required_themes=["City", "Creator", "Technic"], excluded_themes=["Duplo", "Baby"],
                   year_min=2018, max_complexity="moderate")),
    ("budget_parent", dict(price_max=50.0, age_min=6, pieces_max=500, excluded_themes=["Adult Welcome"],
                           max_complexity="simple")),
    ("teen_star_wars", dict(required_themes=["Star Wars"], age_min=10, pieces_min=500, max_complexity="complex")),
    ("adult_collector", dict(price_min=100.0, pieces_min=1000, age_min=16, year_min=2020,
                             required_themes=["Creator Expert", "Architecture", "Technic"])),
    ("gift_buyer", dict(price_max=80.0, price_min=25.0, age_min=8, pieces_min=200, pieces_max=800,
                        required_themes=["City", "Friends", "Creator"])),
    ("space_fan", dict(required_themes=["Space", "City"], pieces_min=100, excluded_themes=["Duplo"])),
    ("no_such_theme_required", dict(required_themes=["Nonexistent Theme XYZ"], pieces_max=500)),
    ("no_such_theme_excluded", dict(excluded_themes=["Nonexistent Theme XYZ"], pieces_max=500)),
    ("toddler", dict(age_min=3)),
    ("age_max_kid", dict(age_max=8, year_max=2010)),
    ("age_max_youth", dict(age_max=12, min_complexity="moderate")),
    ("age_max_teen_noop", dict(age_max=15)),
    ("available", dict(must_be_available=True, min_complexity="complex")),
    ("owned_wishlist", dict(exclude_owned=True, exclude_wishlisted=True, user_id=3, pieces_max=2000)),
    ("owned_no_user", dict(exclude_owned=True, pieces_min=50)),
    ("price_rounding", dict(price_max=29.99, price_min=7.49)),
    ("unknown_complexity", dict(max_complexity="galactic", min_complexity="galactic")),
]

# (name, [(ConstraintType value, value), ...]) -- hand-built HardConstraint lists in the reference tests
SCENARIO_LISTS: List[Tuple[str, list]] = [
    ("basic", [("pieces_max", 500), ("year_min", 2015)]),
    ("budget", [("price_max", 100.0), ("price_min", 20.0)]),
    ("age8_strict", [("age_min", 6)]),
    ("size_medium", [("pieces_min", 201), ("pieces_max", 800)]),
    ("popular_themes", [("themes_required", ["Star Wars", "City", "Creator"])]),
    ("restrictive", [("pieces_min", 5000), ("price_max", 5.0)]),
    ("light", [("pieces_max", 1000), ("year_min", 2010)]),
    ("medium", [("pieces_min", 100), ("pieces_max", 800),
                ("themes_required", ["City", "Creator", "Star Wars"]), ("year_min", 2015)]),
    ("demo_theme", [("themes_required", ["City"])]),
    ("demo_complex", [("pieces_min", 200), ("pieces_max", 1000), ("themes_required", ["Star Wars", "City"]),
                      ("year_min", 2015), ("price_max", 150.0)]),
    ("two_required", [("themes_required", ["City"]), ("themes_required", ["Space"])]),
    ("required_and_excluded", [("themes_required", ["City", "Space"]), ("themes_excluded", ["Police"])]),
    ("like_wildcard", [("themes_required", ["S_ace"])]),
    ("empty", []),
]


def rating_matrix(n_users: int, n_items: int, seed: int, lo: int = 10, hi: int = 30) -> np.ndarray:
    """Dense user x item ratings in the style of the reference's synthetic users
    (recommendation_system.py:340-366): 10-30 ratings per user from {3,4,5} with
    p = (.2,.3,.5)."""
    rng = np.random.default_rng(seed)
    m = np.zeros((n_users, n_items), dtype=np.float64)
    for u in range(n_users):
        k = int(rng.integers(lo, hi + 1))
        items = rng.choice(n_items, size=min(k, n_items), replace=False)
        m[u, items] = rng.choice([3, 4, 5], size=len(items), p=[0.2, 0.3, 0.5])
    return m


def zipf_ratings_csr(n_users: int, n_items: int, mean_per_user: int, seed: int):
    """Config-3 style interactions: Zipf item popularity, ~mean_per_user ratings
    per user, values 1..5 (exact in bf16/fp32).  Returns CSR (indptr int64,
    indices int32 sorted per row, data int8->float32)."""
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, n_items + 1) ** 0.8
    w /= w.sum()
    perm = rng.permutation(n_items)
    counts = np.clip(rng.poisson(mean_per_user, size=n_users), 1, None)
    total = int(counts.sum())
    draws = perm[rng.choice(n_items, size=total, p=w)]
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    idx_out = []
    off = 0
    pos = 0
    for u in range(n_users):
        c = int(counts[u])
        row = np.unique(draws[pos:pos + c])
        pos += c
        idx_out.append(row)
        off += row.shape[0]
        indptr[u + 1] = off
    indices = np.concatenate(idx_out).astype(np.int32)
    data = rng.integers(1, 6, size=indices.shape[0]).astype(np.float32)
    return indptr, indices, data
