"""Row-sharded catalog across the GPUs of one node.

One process per GPU (``torch.distributed``, NCCL over NVLink / NVSwitch for the plumbing).  Catalog rows, metadata
and item factors are partitioned contiguously; queries and their constraint descriptors are replicated.  Each rank
runs the fused scan over its shard.  Two exchange modes:

* ``exchange="p2p"`` (default on GPUs): the shards prune against the UNION of all ranks' thresholds, read through
  NVLink peer memory by the scan kernel itself, and each rank's merge kernel stores its final candidates of query q
  straight into the inbox of q's OWNER rank (q // ceil(B / world)), which merges the ``world`` lists of its slice
  (``bb_search_sharded``).  No collective sits on the data path; every rank ends up with the global top-k of its
  own slice of the queries (``result="slice"``), or of all queries after one all-gather of the slices (``"full"``).
* ``exchange="nccl"``: per-shard top-k, one NCCL all-gather of the B x k (score, id) lists, k-way merge kernel on every
  rank.  Also the path of the CPU (gloo) tests.

Ties stay (-score, +id) because ids are global.
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


def slice_bounds(n_queries: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Query rows owned by ``rank`` in a collective search: [rank * s, min(B, (rank + 1) * s)), s = ceil(B / world)."""
    s = -(-n_queries // world_size)
    lo = min(n_queries, rank * s)
    return lo, min(n_queries, lo + s)


def exchange_comm_handles(engine, group=None) -> bool:
    """All ranks export the cudaIpc handle of their communication block and map everyone else's.  Returns False when
    some engine was created without ``max_batch`` (no block) -- sharding still works through NCCL, each shard then
    prunes on its own."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    try:
        mine = engine.comm_export()
    except Exception:
        mine = None
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    if any(h is None for h in handles):
        return False
    engine.comm_import(rank, handles)
    return True


class ShardedSearch:
    """Couples a per-rank engine (anything with ``search`` and ``merge_topk``) with the candidate exchange.
    ``engine.id_offset`` must equal this rank's ``shard_bounds`` begin."""

    def __init__(self, engine, group=None, share_thresholds: bool = True, exchange: str = "p2p"):
        if dist is None or not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialized (one process per GPU)")
        self.engine = engine
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.shared = False
        nccl = dist.get_backend(group) == "nccl"
        if self.world_size > 1 and nccl and hasattr(engine, "comm_export") and (share_thresholds or exchange == "p2p"):
            self.shared = exchange_comm_handles(engine, group)
        self.exchange = exchange if (self.shared and exchange == "p2p") else "nccl"
        self.share_thresholds = self.shared and share_thresholds

    #: per-query arrays of ``ScoringEngine.search`` and the dtype the engine expects
    _PER_QUERY = {"q_emb": np.float32, "q_cf": np.float32, "q_item": np.int32, "self_id": np.int32,
                  "constraint_index": np.int32, "excl_lists": np.int32}

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

    def _resolve_q_item(self, kwargs):
        """``q_item`` names catalog rows whose stored embedding is the query.  On a row-sharded catalog only the owner
        holds that row: every rank fetches the rows it owns (zeros elsewhere), one all-reduce assembles the batch, and
        the search runs with ``q_emb`` (already unit norm) + the same ``self_id`` semantics."""
        ids = kwargs.pop("q_item")
        if torch.is_tensor(ids):
            ids_np = ids.detach().cpu().numpy().astype(np.int64)
        else:
            ids_np = np.asarray(ids, dtype=np.int64)
        lo = self.engine.id_offset
        local = (ids_np >= lo) & (ids_np < lo + self.engine.n_items)
        rows = np.zeros((ids_np.shape[0], self.engine.dim), dtype=np.float32)
        if local.any():
            rows[local] = self.engine.fetch_embeddings(ids_np[local].astype(np.int32))
        nccl = dist.get_backend(self.group) == "nccl"
        t = torch.from_numpy(rows)
        if nccl:
            t = t.to(f"cuda:{self.engine.device}")
        dist.all_reduce(t, group=self.group)
        owned = torch.from_numpy(local.astype(np.int32))
        if nccl:
            owned = owned.to(t.device)
        dist.all_reduce(owned, group=self.group)
        if int(owned.min().item()) < 1:
            raise ValueError("q_item holds an id that no shard owns")
        kwargs["q_emb"] = t if nccl else t.numpy()
        kwargs["prenormalized"] = True
        return kwargs

    def search(self, k: int, host_out=None, result: str = "full", **kwargs):
        """Same keyword arguments as ``ScoringEngine.search``.  ``result="full"``: the GLOBAL top-k of every query on
        every rank; ``result="slice"``: only the rows of ``slice_bounds(B, world, rank)`` (what the p2p exchange
        produces natively -- no further collective).  ``host_out``: optional (scores, ids) pinned CPU tensors that
        receive the result of a host call."""
        host_call = not any(torch.is_tensor(v) and v.is_cuda for v in kwargs.values())
        if self.world_size > 1 and kwargs.get("q_item") is not None:
            kwargs = self._resolve_q_item(dict(kwargs))
        nccl = self.world_size > 1 and dist.get_backend(self.group) == "nccl"
        b = next(len(kwargs[n]) for n in ("q_emb", "q_item", "q_cf") if kwargs.get(n) is not None)
        if self.world_size > 1 and host_call and b % self.world_size == 0 and b >= self.world_size:
            self._spread_host_inputs(kwargs, b)          # (arrays that are device tensors already are left alone)
        if self.world_size == 1:
            return self.engine.search(k, **kwargs)
        dev = f"cuda:{self.engine.device}"
        if self.exchange == "p2p":
            lo, hi = slice_bounds(b, self.world_size, self.rank)
            s, i = self.engine.search(k, sharded=True, **kwargs)        # [n_mine, k] on the device
            if not torch.is_tensor(s):                                   # (all inputs were host arrays and not spread)
                if result != "full":
                    return s, i
                s, i = torch.from_numpy(s).to(dev), torch.from_numpy(i).to(dev)
            if result == "full":
                sl = -(-b // self.world_size)
                ps = torch.full((sl, k), float("-inf"), dtype=torch.float32, device=dev)
                pi = torch.full((sl, k), -1, dtype=torch.int32, device=dev)
                ps[:hi - lo], pi[:hi - lo] = s, i
                fs = torch.empty((self.world_size * sl, k), dtype=torch.float32, device=dev)
                fi = torch.empty((self.world_size * sl, k), dtype=torch.int32, device=dev)
                dist.all_gather_into_tensor(fs.view(-1), ps.view(-1), group=self.group)
                dist.all_gather_into_tensor(fi.view(-1), pi.view(-1), group=self.group)
                s, i = fs[:b], fi[:b]
            return self._deliver(s, i, host_call, host_out)
        # ---- NCCL exchange: per-shard top-k, all-gather, merge on every rank
        if self.share_thresholds:
            kwargs["share_thresholds"] = True
        if nccl and host_call and kwargs.get("out") is None:
            kwargs["out"] = (torch.empty((b, k), dtype=torch.float32, device=dev),
                             torch.empty((b, k), dtype=torch.int32, device=dev))
        scores, ids = self.engine.search(k, **kwargs)
        as_torch = torch.is_tensor(scores)
        s_t = scores if as_torch else torch.from_numpy(np.ascontiguousarray(scores))
        i_t = ids if as_torch else torch.from_numpy(np.ascontiguousarray(ids))
        all_s = torch.empty((self.world_size, b, k), dtype=s_t.dtype, device=s_t.device)
        all_i = torch.empty((self.world_size, b, k), dtype=i_t.dtype, device=i_t.device)
        dist.all_gather_into_tensor(all_s.view(-1), s_t.reshape(-1).contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_i.view(-1), i_t.reshape(-1).contiguous(), group=self.group)
        if as_torch:
            ms, mi = self.engine.merge_topk(all_s, all_i, k)
        else:
            ms, mi = self.engine.merge_topk(all_s.numpy(), all_i.numpy(), k)
        if result == "slice":
            lo, hi = slice_bounds(b, self.world_size, self.rank)
            ms, mi = ms[lo:hi], mi[lo:hi]
        if as_torch:
            return self._deliver(ms, mi, nccl and host_call, host_out)
        return ms, mi

    # ---- pipelined collective searches (p2p exchange, device inputs) ------------------------------------------
    def search_pipelined(self, k: int, **kwargs):
        """Back-to-back collective searches with the owner-side merge of search s launched inside search s+1 (after its
        prologue, before its scan: ``bb_sharded_scatter_after``), so that waiting for the slowest peer's keys overlaps
        work that does not depend on the peers.  Takes the keyword arguments of ``ScoringEngine.search`` (CUDA tensors)
        and returns the (scores, ids) CUDA tensors of THIS search's slice of the queries -- their contents are complete
        after the next ``search_pipelined`` call has been enqueued, or after ``flush()``."""
        if self.world_size == 1 or self.exchange != "p2p":
            s, i = self.search(k, result="slice", **kwargs)
            return s, i
        if kwargs.get("q_item") is not None:
            kwargs = self._resolve_q_item(dict(kwargs))
        prev = getattr(self, "_pending", None)
        dev = f"cuda:{self.engine.device}"
        n_mine = self.engine.sharded_scatter(k, collect_previous_into=None if prev is None else (prev[0], prev[1]), **kwargs)
        alloc = max(n_mine, 1)
        out = (torch.empty((alloc, k), dtype=torch.float32, device=dev), torch.empty((alloc, k), dtype=torch.int32, device=dev))
        self._pending = (out[0], out[1], k, n_mine)
        return out[0][:n_mine], out[1][:n_mine]

    def flush(self):
        """Finishes the last pipelined search (its merge is enqueued on the current stream)."""
        prev = getattr(self, "_pending", None)
        if prev is None:
            return
        self._pending = None
        self.engine.sharded_collect(prev[2], prev[3], out=(prev[0], prev[1]))

    def _deliver(self, s, i, host_call: bool, host_out):
        if not host_call:
            return s, i
        if host_out is not None:
            n = s.shape[0]
            host_out[0][:n].copy_(s, non_blocking=True)
            host_out[1][:n].copy_(i, non_blocking=True)
            torch.cuda.synchronize(self.engine.device)
            return host_out[0][:n].numpy(), host_out[1][:n].numpy()
        return s.cpu().numpy(), i.cpu().numpy()
