"""Multi-GPU parity check of the row-sharded search (run under torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/shard_check.py

Every rank holds one shard; ``ShardedSearch`` (NCCL all-gather of the per-shard top-k + merge, running thresholds
shared through NVLink peer memory) must return exactly what ONE engine over the whole catalog returns, with and
without threshold sharing, on the exact fp32 path, and the same recall on the bf16 + re-score path.  Also prints
how many candidates each shard's scan inserted with / without sharing (the point of sharing).
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
from brickbrain_rec_engine_b200 import ScoringEngine, ShardedSearch, shard_bounds  # noqa: E402
from brickbrain_rec_engine_b200 import _lib as L  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n, d, f, b, k = int(os.environ.get("SC_N", "400000")), 384, 64, int(os.environ.get("SC_B", "1024")), 100
    themes = synth.themes_table()
    emb = synth.embeddings(n, d, 3, "aniso")
    itf = synth.cf_factors(n, f, 4)
    parts, year, theme = synth.metadata(n, 5, themes)
    q = synth.embeddings(b, d, 6, "aniso") * 2.5
    u = synth.cf_factors(b, f, 7)
    self_id = np.random.default_rng(8).integers(0, n, b).astype(np.int32)
    cons = np.zeros(4, dtype=L.CONSTRAINT_DTYPE)
    cons[:] = L.no_constraint()
    cons[1]["parts_min"], cons[1]["parts_max"] = 200, 900
    cons[2]["year_min"] = 2012
    cons[3]["parts_max"], cons[3]["year_max"] = 150, 2005
    cidx = (np.arange(b) % 4).astype(np.int32)

    full = ScoringEngine(n, d, f, device=local, max_batch=b)
    full.load_catalog(emb=emb, item_factors=itf, num_parts=parts, year=year, theme_id=theme)
    lo, hi = shard_bounds(n, world, rank)
    part = ScoringEngine(hi - lo, d, f, device=local, id_offset=lo, max_batch=b)
    part.load_catalog(emb=emb[lo:hi], item_factors=itf[lo:hi], num_parts=parts[lo:hi], year=year[lo:hi], theme_id=theme[lo:hi])
    kw = dict(q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, constraint_index=cidx, self_id=self_id)

    ok = True
    plain = ShardedSearch(part, share_thresholds=False)
    shared = ShardedSearch(part, share_thresholds=True)
    assert shared.shared, "threshold handles were not exchanged"
    for prec in ("fp32", "bf16_rescore"):
        fs, fi = full.search(k, precision=prec, **kw)
        ref_s, ref_i = full.search(k, precision="fp32", **kw) if prec != "fp32" else (fs, fi)
        stats = {}
        for name, ss in (("plain", plain), ("shared", shared), ("shared-again", shared)):
            s, i = ss.search(k, precision=prec, **kw)
            stats[name] = part.info()["last_candidates"]
            if prec == "fp32":
                same = np.array_equal(i, fi) and np.array_equal(s, fs)
            else:
                rec = np.mean([len(set(i[r]) & set(ref_i[r])) / k for r in range(b)])
                m = i == ref_i
                same = rec >= 0.999 and np.allclose(s[m], ref_s[m], rtol=1e-5, atol=2e-6)
            if not same:
                ok = False
                print(f"[rank {rank}] MISMATCH {prec} {name}", flush=True)
        # an unflagged search in between must not disturb the collective sequence
        if rank == 0:
            part.search(k, precision=prec, **kw)
        print(f"[rank {rank}] {prec}: candidates inserted per search {stats}", flush=True)
    flag = torch.tensor([0 if ok else 1], device=f"cuda:{local}")
    dist.all_reduce(flag)
    if rank == 0:
        print("SHARD_CHECK " + ("OK" if flag.item() == 0 else "FAILED"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
