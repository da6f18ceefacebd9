"""Bulk recommendation generation (BASELINE config 5): every user x the whole
catalog, hybrid top-k, in batches through the resident engine.

The content query of a user is the stored embedding of a liked set (gathered on
the device by ``q_item``), the CF query is the user's factor row.  Batches are
issued back to back on one stream with device-resident inputs and outputs, so the
host only enqueues; results are handed to ``sink`` (e.g. a pinned-memory writer).
"""
from __future__ import annotations

from typing import Callable, Optional


def bulk_recommend(engine, user_factors, liked_items, top_k: int, batch: int = 8192, w_content: float = 0.4,
                   w_cf: float = 0.6, constraints=None, precision: str = "bf16_rescore",
                   sink: Optional[Callable[[int, object, object], None]] = None):
    """user_factors [U, F] and liked_items [U] (global ids) may be NumPy arrays or CUDA tensors.
    Returns the list of (scores, ids) per batch when no sink is given."""
    n_users = int(user_factors.shape[0])
    outs = []
    for lo in range(0, n_users, batch):
        hi = min(n_users, lo + batch)
        items = liked_items[lo:hi] if liked_items is not None else None
        cons = None
        if constraints is not None:
            cons = constraints if len(constraints) == 1 else constraints[lo:hi]
        scores, ids = engine.search(top_k, q_item=items, q_cf=user_factors[lo:hi], w_content=w_content if items is not None else 0.0,
                                    w_cf=w_cf, constraints=cons, self_id=items, precision=precision)
        if sink is not None:
            sink(lo, scores, ids)
        else:
            outs.append((scores, ids))
    return outs
