"""Multi-GPU parity check of the row-sharded search (run under torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/shard_check.py

Every rank holds one shard; ``ShardedSearch`` must return exactly what ONE engine over the whole catalog returns, on
the exact fp32 path bit for bit and with recall >= 0.999 on the bf16 + re-score path, through BOTH exchanges: ``p2p``
(union thresholds over NVLink peer memory + owner inboxes, fused into the kernels) and ``nccl`` (all-gather + merge
kernel), with host and device inputs, plain and pipelined (``search_pipelined``), for ``q_emb`` and for ``q_item`` queries
(resolved by the owner shard).  Prints
how many candidates each shard's scan appended with / without the union thresholds.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import synth  # noqa: E402
from brickbrain_rec_engine_b200 import ScoringEngine, ShardedSearch, shard_bounds, slice_bounds  # noqa: E402
from brickbrain_rec_engine_b200 import _lib as L  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    n, d, f, b, k = int(os.environ.get("SC_N", "400000")), 384, 64, int(os.environ.get("SC_B", "1024")), 100
    themes = synth.themes_table()
    emb = synth.embeddings(n, d, 3, "aniso")
    itf = synth.cf_factors(n, f, 4)
    parts, year, theme = synth.metadata(n, 5, themes)
    q = synth.embeddings(b, d, 6, "aniso") * 2.5
    u = synth.cf_factors(b, f, 7)
    rng = np.random.default_rng(8)
    self_id = rng.integers(0, n, b).astype(np.int32)
    rated = [np.sort(rng.choice(n, 40, replace=False)).astype(np.int32) for _ in range(b)]
    cons = np.zeros(4, dtype=L.CONSTRAINT_DTYPE)
    cons[:] = L.no_constraint()
    cons[1]["parts_min"], cons[1]["parts_max"] = 200, 900
    cons[2]["year_min"] = 2012
    cons[3]["parts_max"], cons[3]["year_max"] = 150, 2005
    cidx = (np.arange(b) % 4).astype(np.int32)

    full = ScoringEngine(n, d, f, device=local, max_batch=b)
    full.load_catalog(emb=emb, item_factors=itf, num_parts=parts, year=year, theme_id=theme)
    full.set_exclusion_lists(rated)
    lo, hi = shard_bounds(n, world, rank)
    part = ScoringEngine(hi - lo, d, f, device=local, id_offset=lo, max_batch=b)
    part.load_catalog(emb=emb[lo:hi], item_factors=itf[lo:hi], num_parts=parts[lo:hi], year=year[lo:hi], theme_id=theme[lo:hi])
    part.set_exclusion_lists(rated)
    base = dict(q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, constraint_index=cidx, self_id=self_id,
                excl_lists=np.arange(b, dtype=np.int32))
    forms = {"q_emb": dict(base, q_emb=q), "q_item": dict(base, q_item=self_id)}

    ok = True
    p2p = ShardedSearch(part, exchange="p2p")
    nccl = ShardedSearch(part, exchange="nccl", share_thresholds=False)
    nccl_shared = ShardedSearch(part, exchange="nccl", share_thresholds=True)
    assert p2p.exchange == "p2p", "communication blocks were not exchanged"
    s_lo, s_hi = slice_bounds(b, world, rank)

    def same(prec, s, i, fs, fi, ref_i, ref_s):
        s, i = (x.cpu().numpy() if torch.is_tensor(x) else x for x in (s, i))
        if prec == "fp32":
            # two exact paths (tensor pipe + re-score on the single engine, FFMA scan on the shards) sum in different
            # orders: scores equal to fp32 rounding, ids identical except where two neighbours tie within that rounding
            if not np.allclose(s, fs, rtol=1e-5, atol=2e-6):
                return False
            diff = i != fi
            if not diff.any():
                return True
            gap_ok = np.abs(np.where(diff, s - fs, 0.0)) <= 4e-6 * np.maximum(1.0, np.abs(fs))
            return bool(gap_ok.all() and diff.mean() < 0.005)
        rec = np.mean([len(set(i[r]) & set(ref_i[r])) / k for r in range(i.shape[0])]) if i.shape[0] else 1.0
        m = i == ref_i
        return rec >= 0.999 and np.allclose(s[m], ref_s[m], rtol=1e-5, atol=2e-6)

    for form, kw in forms.items():
        for prec in ("fp32", "bf16_rescore"):
            fs, fi = full.search(k, precision=prec, **kw)
            ref_s, ref_i = full.search(k, precision="fp32", **kw) if prec != "fp32" else (fs, fi)
            stats = {}
            for name, ss, result in (("nccl", nccl, "full"), ("nccl+union", nccl_shared, "full"), ("p2p", p2p, "full"),
                                     ("p2p-slice", p2p, "slice"), ("p2p-again", p2p, "slice")):
                s, i = ss.search(k, precision=prec, result=result, **kw)
                stats[name] = part.info()["last_appends"]
                sl = slice(None) if result == "full" else slice(s_lo, s_hi)
                if not same(prec, s, i, fs[sl], fi[sl], ref_i[sl], ref_s[sl]):
                    ok = False
                    print(f"[rank {rank}] MISMATCH {form} {prec} {name}", flush=True)
            # device-resident inputs through the p2p exchange
            kw_dev = {kk: (torch.from_numpy(np.ascontiguousarray(vv)).to(dev) if isinstance(vv, np.ndarray) and kk != "constraints" else vv)
                      for kk, vv in kw.items()}
            s, i = p2p.search(k, precision=prec, result="slice", **kw_dev)
            if not same(prec, s, i, fs[s_lo:s_hi], fi[s_lo:s_hi], ref_i[s_lo:s_hi], ref_s[s_lo:s_hi]):
                ok = False
                print(f"[rank {rank}] MISMATCH {form} {prec} p2p device inputs", flush=True)
            # pipelined: three searches back to back, each one's owner-side merge launched inside the next call
            outs = [p2p.search_pipelined(k, precision=prec, **kw_dev) for _ in range(3)]
            p2p.flush()
            torch.cuda.synchronize()
            for x, (s, i) in enumerate(outs):
                if not same(prec, s, i, fs[s_lo:s_hi], fi[s_lo:s_hi], ref_i[s_lo:s_hi], ref_s[s_lo:s_hi]):
                    ok = False
                    print(f"[rank {rank}] MISMATCH {form} {prec} p2p pipelined search {x}", flush=True)
            # an unflagged search in between must not disturb the collective sequence
            if rank == 0:
                part.search(k, precision=prec, **{kk: vv for kk, vv in kw.items() if kk != "q_item"}, **({"q_emb": q} if form == "q_item" else {}))
            print(f"[rank {rank}] {form} {prec}: candidates appended by this shard's scan {stats}", flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("SHARD_CHECK " + ("OK" if flag.item() == 0 else "FAILED"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
