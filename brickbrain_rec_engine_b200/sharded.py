"""Row-sharded catalog across the GPUs of one node.

One process per GPU (``torch.distributed``, NCCL over NVLink / NVSwitch).  Catalog
rows, metadata and item factors are partitioned contiguously; queries and their
constraint descriptors are replicated.  Each rank runs the fused scan over its
shard and emits B x k (score, GLOBAL id) candidates; the only data-path
collective is one all-gather of those candidates (B*k*8 bytes per rank), followed
by a k-way merge kernel on every rank.  Ties stay (-score, +id) because ids are global.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range [begin, end) of ``rank``; remainders go to the low ranks."""
    base, rem = divmod(n_items, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def exchange_threshold_handles(engine, group=None) -> bool:
    """All ranks export the cudaIpc handle of their per-query threshold array and map everyone else's, so the
    scan kernels of all shards raise ONE logical threshold per query through NVLink peer memory (a rank's local
    k-th best is a valid lower bound of the global k-th best).  Returns False when the engine was created without
    max_batch (no exportable array) -- sharding still works, each shard then prunes on its own."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    try:
        mine = engine.export_threshold_handle()
    except Exception:
        mine = None
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    if any(h is None for h in handles):
        return False
    engine.import_peer_thresholds([h for r, h in enumerate(handles) if r != rank])
    return True


class ShardedSearch:
    """Couples a per-rank engine (anything with ``search`` and ``merge_topk``) with the
    candidate all-gather.  ``engine.id_offset`` must equal this rank's ``shard_bounds`` begin."""

    def __init__(self, engine, group=None, share_thresholds: bool = False):
        if dist is None or not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialized (one process per GPU)")
        self.engine = engine
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.shared = False
        if share_thresholds and self.world_size > 1 and hasattr(engine, "export_threshold_handle"):
            self.shared = exchange_threshold_handles(engine, group)

    #: per-query arrays of ``ScoringEngine.search`` and the dtype the engine expects
    _PER_QUERY = {"q_emb": np.float32, "q_cf": np.float32, "q_item": np.int32, "self_id": np.int32, "constraint_index": np.int32}

    def _spread_host_inputs(self, kwargs, b):
        """Host (NumPy) query batches are replicated on every rank.  Instead of every rank pushing the whole batch
        through its own PCIe link, each rank uploads 1/world of every per-query array and the slices are
        all-gathered over NVLink into device tensors (total H2D = one batch, not world batches)."""
        per = b // self.world_size
        dev = f"cuda:{self.engine.device}" if dist.get_backend(self.group) == "nccl" else "cpu"   # (gloo: CPU tests)
        for name, dt in self._PER_QUERY.items():
            v = kwargs.get(name)
            if v is None or torch.is_tensor(v):
                continue
            t = torch.from_numpy(np.ascontiguousarray(v, dtype=dt))
            part = t[self.rank * per:(self.rank + 1) * per].to(dev, non_blocking=True)
            full = torch.empty((b,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            dist.all_gather_into_tensor(full.view(-1), part.reshape(-1), group=self.group)
            kwargs[name] = full

    def search(self, k: int, host_out=None, **kwargs):
        """Same keyword arguments as ``ScoringEngine.search``; returns the GLOBAL top-k on every rank.
        ``host_out``: optional (scores, ids) CPU tensors (pinned) that receive the merged result of a host call."""
        if self.shared:
            kwargs["share_thresholds"] = True
        nccl = self.world_size > 1 and dist.get_backend(self.group) == "nccl"
        host_call = not any(torch.is_tensor(v) and v.is_cuda for v in kwargs.values())
        if self.world_size > 1 and host_call:
            b = next(len(kwargs[n]) for n in ("q_emb", "q_item", "q_cf") if kwargs.get(n) is not None)
            if b % self.world_size == 0 and b >= self.world_size:
                self._spread_host_inputs(kwargs, b)
            # the per-shard result stays on the device for the gather + merge
            if nccl and kwargs.get("out") is None:
                dev = f"cuda:{self.engine.device}"
                kwargs["out"] = (torch.empty((b, k), dtype=torch.float32, device=dev),
                                 torch.empty((b, k), dtype=torch.int32, device=dev))
        scores, ids = self.engine.search(k, **kwargs)
        if self.world_size == 1:
            return scores, ids
        as_torch = torch.is_tensor(scores)
        s_t = scores if as_torch else torch.from_numpy(np.ascontiguousarray(scores))
        i_t = ids if as_torch else torch.from_numpy(np.ascontiguousarray(ids))
        b = s_t.shape[0]
        all_s = torch.empty((self.world_size, b, k), dtype=s_t.dtype, device=s_t.device)
        all_i = torch.empty((self.world_size, b, k), dtype=i_t.dtype, device=i_t.device)
        dist.all_gather_into_tensor(all_s.view(-1), s_t.reshape(-1).contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_i.view(-1), i_t.reshape(-1).contiguous(), group=self.group)
        if as_torch:
            ms, mi = self.engine.merge_topk(all_s, all_i, k)
            if nccl and host_call:
                if host_out is not None:
                    host_out[0].copy_(ms, non_blocking=True)
                    host_out[1].copy_(mi, non_blocking=True)
                    torch.cuda.synchronize(self.engine.device)
                    return host_out[0].numpy(), host_out[1].numpy()
                return ms.cpu().numpy(), mi.cpu().numpy()
            return ms, mi
        return self.engine.merge_topk(all_s.numpy(), all_i.numpy(), k)
