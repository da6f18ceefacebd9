"""Bulk recommendation generation (BASELINE config 5): every user x the whole catalog, hybrid top-k.

Layout on N GPUs: the users are sharded over the ranks (``user_shard``), the catalog is REPLICATED on every
GPU (1M x 448: 0.9 GB bf16 + 1.8 GB fp32), so there is no collective on the data path.  Each rank walks its users in
batches through the resident engine; the content query of a user is the stored embedding of a liked set (gathered on
the device by ``q_item``), the CF query is the user's factor row.

The step after the path (the reference writes every result row into ``user_recommendations`` with one SQL
INSERT per row, recommendation_api.py:2005-2036) is the ``BulkSink``: two device output buffers and two PINNED host
buffers per array; the device->host copy of batch i runs on its own stream while batch i+1 is being searched, and a
consumer (``RecommendationCache.bulk_upsert``, a file writer, ...) receives NumPy views of the pinned slot.
1M users x top-100 is 0.8 GB of (score, id) pairs: ~15 ms of PCIe time hidden behind ~1 s of scoring.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def user_shard(n_users: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous user range [begin, end) of ``rank``; remainders go to the low ranks."""
    base, rem = divmod(n_users, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class BulkSink:
    """Pinned double-buffered device->host stream of result batches.

    ``consumer(user_lo, scores, ids)`` is called with NumPy views ([n, k] float32 / int32) of a pinned slot once that
    slot's copy has landed; the views are only valid during the call.  ``keep=True`` collects copies instead
    (``results()``)."""

    def __init__(self, k: int, batch: int, device, consumer: Optional[Callable[[int, np.ndarray, np.ndarray], None]] = None,
                 keep: bool = False):
        if torch is None or not torch.cuda.is_available():
            raise RuntimeError("BulkSink needs a CUDA device")
        self.k, self.batch, self.device = k, batch, device
        self.consumer = consumer
        self.keep = keep
        self.kept: List[Tuple[int, np.ndarray, np.ndarray]] = []
        self.copy_stream = torch.cuda.Stream(device=device)
        self.dev = [(torch.empty((batch, k), dtype=torch.float32, device=device),
                     torch.empty((batch, k), dtype=torch.int32, device=device)) for _ in range(2)]
        self.host = [(torch.empty((batch, k), dtype=torch.float32).pin_memory(),
                      torch.empty((batch, k), dtype=torch.int32).pin_memory()) for _ in range(2)]
        self.done = [None, None]          # (event, user_lo, n) of the copy in flight per slot
        self.turn = 0
        self.rows = 0
        self.bytes = 0

    def device_out(self, n: int):
        """Device buffers the next search should write into (the slot's previous copy is awaited on the device)."""
        slot = self.turn & 1
        self._drain(slot)
        s, i = self.dev[slot]
        return s[:n], i[:n]

    def _drain(self, slot: int):
        pending = self.done[slot]
        if pending is None:
            return
        ev, user_lo, n = pending
        ev.synchronize()
        hs, hi = self.host[slot]
        sv, iv = hs[:n].numpy(), hi[:n].numpy()
        if self.consumer is not None:
            self.consumer(user_lo, sv, iv)
        if self.keep:
            self.kept.append((user_lo, sv.copy(), iv.copy()))
        self.done[slot] = None

    def __call__(self, user_lo: int, scores, ids):
        """Queues the device->host copy of the batch that was just searched into ``device_out``."""
        slot = self.turn & 1
        n = int(scores.shape[0])
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            self.host[slot][0][:n].copy_(scores, non_blocking=True)
            self.host[slot][1][:n].copy_(ids, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        # (the next search into this slot goes through device_out(), which awaits this copy first)
        self.done[slot] = (ev, user_lo, n)
        self.turn += 1
        self.rows += n
        self.bytes += n * self.k * 8

    def flush(self):
        for slot in ((self.turn & 1), (self.turn + 1) & 1):
            self._drain(slot)

    def results(self):
        """(user_lo sorted) concatenated (scores, ids) of everything kept."""
        self.flush()
        self.kept.sort(key=lambda t: t[0])
        return (np.concatenate([s for _, s, _ in self.kept]), np.concatenate([i for _, _, i in self.kept]))


def bulk_recommend(engine, user_factors, liked_items, top_k: int, batch: int = 8192, w_content: float = 0.4,
                   w_cf: float = 0.6, constraints=None, precision: str = "bf16_rescore",
                   sink: Optional[Callable[[int, object, object], None]] = None, user_offset: int = 0,
                   excl_lists=None, constraint_index=None):
    """user_factors [U, F] and liked_items [U] (global ids) may be NumPy arrays or CUDA tensors (this rank's users).
    Batches are issued back to back on the current stream; with a ``BulkSink`` the searches write into the sink's
    device buffers and the host copies overlap the next batch.  ``user_offset``: global index of this rank's first
    user (what the sink's consumer sees).  Returns the list of (scores, ids) per batch when no sink is given."""
    n_users = int(user_factors.shape[0])
    outs = []
    for lo in range(0, n_users, batch):
        hi = min(n_users, lo + batch)
        items = liked_items[lo:hi] if liked_items is not None else None
        cons = None
        cidx = None
        if constraints is not None:
            if constraint_index is not None:
                cons, cidx = constraints, constraint_index[lo:hi]
            else:
                cons = constraints if len(constraints) == 1 else constraints[lo:hi]
        out = sink.device_out(hi - lo) if isinstance(sink, BulkSink) else None
        scores, ids = engine.search(top_k, q_item=items, q_cf=user_factors[lo:hi], w_content=w_content if items is not None else 0.0,
                                    w_cf=w_cf, constraints=cons, constraint_index=cidx, self_id=items, precision=precision, out=out,
                                    excl_lists=None if excl_lists is None else excl_lists[lo:hi])
        if sink is not None:
            sink(user_offset + lo, scores, ids)
        else:
            outs.append((scores, ids))
    return outs
