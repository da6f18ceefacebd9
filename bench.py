#!/usr/bin/env python
"""Benchmark of the candidate-scoring hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, all host threads)
    python bench.py --workload config2|config3|config5 ...   # the other BASELINE configs (lines kept in profiles/)

Default workload (BASELINE config 4): a step = one pass of the hot path over one batch: 8192 hybrid queries
(w_c*cos + w_f*(u.v), per-query hard-constraint set, per-user already-rated exclusion list, self exclusion, top-100)
against the whole 10M x (384+64) catalog.  With N>1 the catalog is row-sharded over the ranks (one process per GPU);
the shards prune against the union of all ranks' thresholds (peer memory) and deliver their candidates to the
queries' owner ranks inside the merge kernel (``--exchange p2p``, default) or through one NCCL all-gather + merge
(``--exchange nccl``).  Total work per step is fixed, so scaling is "strong".

After the timed loops the run CHECKS what it timed: 64 sampled queries are compared with an independent exact
brute force (torch fp32 matmul + torch mask predicates over the regenerated catalog, all shards): recall@k of the
timed precision, mask violations among the returned ids, id mismatches of the exact-fp32 engine path.

One JSON line is printed by rank 0 (see the keys in ``main``).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hybrid top-k queries/sec @10Mx384 catalog"
UNIT = "queries/s"

THEME_WORDS = ["Star Wars", "City", "Creator", "Technic", "Friends", "Duplo", "Architecture", "Space", "Creator Expert",
               "Ninjago", "Harry Potter", "Castle", "Pirates", "Trains", "Baby", "Adult Welcome", "Space Police",
               "Classic Space", "Town", "Racers", "Bionicle", "Minecraft", "Super Heroes", "Ideas", "Speed Champions",
               "Disney", "Jurassic World", "Art", "Icons", "Dots"]


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="config4", choices=["config4", "config2", "config3", "config5"])
    p.add_argument("--n-items", type=int, default=None)
    p.add_argument("--batch", type=int, default=None)
    p.add_argument("--k", type=int, default=None)
    p.add_argument("--dim", type=int, default=384)
    p.add_argument("--cf-dim", type=int, default=64)
    p.add_argument("--precision", default="bf16_rescore", choices=["fp32", "bf16", "bf16_rescore"])
    p.add_argument("--constraint-form", default="indexed", choices=["indexed", "per-query"],
                   help="indexed: distinct constraint sets + per-query index (class bitmaps); per-query: one descriptor per query")
    p.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                   help="N>1: p2p = union thresholds + owner inboxes over NVLink peer memory, fused into the kernels; "
                        "nccl = per-shard top-k, NCCL all-gather, merge kernel")
    p.add_argument("--no-pipeline", action="store_true",
                   help="N > 1, p2p exchange: collect every search before the next one starts (default: the owner-side merge of "
                        "step s runs inside step s+1, after its prologue and before its scan)")
    p.add_argument("--no-rated", action="store_true", help="drop the per-user already-rated exclusion lists")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-check", action="store_true", help="skip the post-run correctness block")
    p.add_argument("--check-queries", type=int, default=64)
    p.add_argument("--cpu-sample-items", type=int, default=1_000_000)
    p.add_argument("--cpu-sample-queries", type=int, default=1024)
    p.add_argument("--users", type=int, default=1_000_000, help="config5 / config3: number of users")
    a = p.parse_args()
    defaults = {"config4": (10_000_000, 8192, 100), "config2": (25_216, 4096, 50), "config3": (25_216, 0, 20),
                "config5": (1_000_000, 8192, 100)}[a.workload]
    a.n_items = a.n_items or defaults[0]
    a.batch = a.batch or defaults[1]
    a.k = a.k or defaults[2]
    return a


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_sustained": d["bf16_tflops_sustained"], "bf16_burst": d["bf16_tflops"], "hbm": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def host_cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        """Start of the timed window: earlier samples (warm-up) are dropped."""
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)

        def collect(t_lo, t_hi):
            sm, mx, reasons = [], None, set()
            for ts, l in self.lines:
                if ts < t_lo or ts > t_hi:
                    continue
                f = [x.strip() for x in l.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons

        sm, mx, reasons = collect(t_mark, float("inf"))
        out = {}
        if not sm:
            # a timed region shorter than the sampling period (config 2 / 3: a few ms): report the last second of the
            # warm-up, which runs the same steps back to back
            sm, mx, reasons = collect(t_mark - 1.0, t_mark)
            out["window"] = "last second of the warm-up (the timed region is shorter than the 100 ms sampling period)"
        out.update({"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)})
        return out


# ------------------------------------------------------------------------------------------ synthetic workload
def themes_table(n_themes: int = 411, max_id: int = 777, seed: int = 7):
    """411 theme ids scattered in [1, 777]; names reuse the words the reference's scenarios query."""
    rng = np.random.default_rng(seed)
    ids = np.sort(rng.choice(np.arange(1, max_id + 1), size=n_themes, replace=False))
    out = {}
    for j, tid in enumerate(ids):
        base = THEME_WORDS[j % len(THEME_WORDS)]
        out[int(tid)] = base if j < len(THEME_WORDS) else f"{base} Series {j // len(THEME_WORDS)}"
    return out


def scenario_mix():
    """The reference's light / medium / heavy / no-constraint scenarios (tests/unit/test_recommendations.py:511-541)
    as (ConstraintType value, value) pairs."""
    return [
        [("pieces_max", 1000), ("year_min", 2010)],
        [("pieces_min", 100), ("pieces_max", 800), ("themes_required", ["City", "Creator", "Star Wars"]), ("year_min", 2015)],
        [("price_max", 200.0), ("price_min", 30.0), ("pieces_min", 200), ("pieces_max", 1500), ("age_min", 8),
         ("themes_required", ["City", "Creator", "Technic"]), ("themes_excluded", ["Duplo", "Baby"]), ("year_min", 2018),
         ("complexity_max", "moderate")],
        [],
    ]


def scenario_descriptors(batch: int, seed: int = 5):
    """Per-query bb_constraint records + theme bitmaps + the per-query scenario index."""
    from brickbrain_rec_engine_b200 import ConstraintCompiler, ConstraintType, HardConstraint, _lib as L
    themes = themes_table()
    comp = ConstraintCompiler(themes, theme_bits=1024, now_year=2026)
    compiled = [comp.compile([HardConstraint(ConstraintType(t), v) for t, v in c]) for c in scenario_mix()]
    bitmaps, index = [], {}
    descs = np.stack([L.no_constraint() for _ in range(batch)])
    pick = np.random.default_rng(seed).integers(0, len(compiled), size=batch)
    for j in range(batch):
        cc = compiled[pick[j]]
        d = cc.desc.copy()
        if cc.theme_bitmap is not None:
            key = cc.theme_bitmap.tobytes()
            if key not in index:
                index[key] = len(bitmaps)
                bitmaps.append(cc.theme_bitmap)
            d["theme_set"] = index[key]
        descs[j] = d
    return descs, np.stack(bitmaps), themes, pick.astype(np.int32)


def gen_catalog_chunk(torch, dev, lo, hi, dim, cf_dim, mu, theme_ids, seed):
    g = torch.Generator(device=dev).manual_seed(seed * 1_000_003 + lo)
    n = hi - lo
    x = torch.randn(n, dim, device=dev, generator=g)
    x = torch.nn.functional.normalize(x, dim=1)
    x = torch.nn.functional.normalize(0.8 * mu[None, :] + 0.6 * x, dim=1)          # anisotropic, mean cosine ~0.64
    v = torch.randn(n, cf_dim, device=dev, generator=g) / (cf_dim ** 0.5)
    parts = torch.exp(torch.rand(n, device=dev, generator=g) * float(np.log(12000.0))).to(torch.int32).clamp_(1, 12000)
    year = torch.randint(1950, 2026, (n,), device=dev, generator=g).to(torch.int16)
    theme = theme_ids[torch.randint(0, theme_ids.numel(), (n,), device=dev, generator=g)].clone()
    theme[torch.rand(n, device=dev, generator=g) < 0.01] = -1
    return x, v, parts, year, theme.to(torch.int16)


class Dist:
    """torch.distributed plumbing (one process per GPU); degenerates to no-ops at world 1."""

    def __init__(self, torch):
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = f"cuda:{self.local}"
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device(self.dev))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, v: float) -> float:
        if self.dist is None:
            return v
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v: float) -> float:
        if self.dist is None:
            return v
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def load_synthetic_catalog(torch, eng, lo, hi, d, f, dev, themes, seed=7):
    theme_ids = torch.tensor(sorted(themes), device=dev, dtype=torch.int32)
    gmu = torch.Generator(device=dev).manual_seed(1234)
    mu = torch.nn.functional.normalize(torch.randn(d, device=dev, generator=gmu), dim=0)
    chunk = 1 << 20
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        x, v, parts, year, theme = gen_catalog_chunk(torch, dev, c0, c1, d, f, mu, theme_ids, seed)
        eng.load_catalog(emb=x, item_factors=v if f else None, num_parts=parts, year=year, theme_id=theme, row_begin=c0 - lo)
    torch.cuda.synchronize()
    return mu, theme_ids


# ------------------------------------------------------------------------------------------ post-run correctness
def exact_bruteforce(torch, D, lo, hi, d, f, mu, theme_ids, qs, us, self_ids, cls, cons, theme_bitmaps, rated, k, returned):
    """Independent exact answer for the sampled queries over THIS rank's rows [lo, hi): torch fp32 matmul (no TF32) over
    the regenerated catalog chunks, the constraint sets as torch predicates on the regenerated metadata, self and
    already-rated exclusion, running top-k.  Also counts how many of the engine's `returned` ids (global, [S, k'])
    that live in this shard violate their query's mask."""
    dev = D.dev
    torch.backends.cuda.matmul.allow_tf32 = False
    S = qs.shape[0]
    qn = torch.nn.functional.normalize(qs, dim=1)
    best_s = torch.full((S, k), float("-inf"), device=dev)
    best_i = torch.full((S, k), -1, dtype=torch.int64, device=dev)
    violations = 0
    tb = torch.from_numpy(theme_bitmaps.astype(np.int64)).to(dev) if theme_bitmaps is not None else None
    chunk = 1 << 20
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        x, v, parts, year, theme = gen_catalog_chunk(torch, dev, c0, c1, d, f, mu, theme_ids, 7)
        sc = 0.4 * (qn @ x.T) + 0.6 * (us @ v.T)
        parts_l, year_l, theme_l = parts.long(), year.long(), theme.long()
        valid_c = []
        for c in cons:
            ok = (parts_l >= int(c["parts_min"])) & (parts_l <= int(c["parts_max"])) & (year_l >= int(c["year_min"])) & (year_l <= int(c["year_max"]))
            if int(c["theme_set"]) >= 0:
                t = theme_l.clamp(min=0)
                bit = (tb[int(c["theme_set"])][t >> 5] >> (t & 31)) & 1
                ok &= (theme_l >= 0) & (bit == 1)
            if int(c["flags"]) & 1:
                ok &= False
            valid_c.append(ok)
        valid = torch.stack(valid_c)[cls.long()]                       # [S, chunk]
        rows = torch.arange(S, device=dev)
        sl = self_ids.long() - c0
        m = (sl >= 0) & (sl < c1 - c0)
        valid[rows[m], sl[m]] = False
        if rated is not None:
            rl = rated.long() - c0                                      # [S, R]
            mr = (rl >= 0) & (rl < c1 - c0)
            rr = rows[:, None].expand_as(rl)
            valid[rr[mr], rl[mr]] = False
        sc = torch.where(valid, sc, torch.full_like(sc, float("-inf")))
        ts, ti = torch.topk(sc, min(k, c1 - c0), dim=1)
        cat_s = torch.cat([best_s, ts], dim=1)
        cat_i = torch.cat([best_i, ti + c0], dim=1)
        # (-score, +id) order: sort by id first, then stably by descending score
        o1 = torch.argsort(cat_i, dim=1, stable=True)
        cat_s, cat_i = torch.gather(cat_s, 1, o1), torch.gather(cat_i, 1, o1)
        o2 = torch.argsort(cat_s, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_i = torch.gather(cat_s, 1, o2), torch.gather(cat_i, 1, o2)
        for ret in returned:
            rl = ret.long() - c0
            mr = (rl >= 0) & (rl < c1 - c0) & (ret >= 0)
            rr = rows[:, None].expand_as(rl)
            violations += int((~valid[rr[mr], rl[mr]]).sum().item())
    best_i = torch.where(torch.isinf(best_s), torch.full_like(best_i, -1), best_i)
    return best_s, best_i, violations


def merge_exact_over_ranks(torch, D, s, i, k):
    if D.dist is None:
        return s, i
    gs = torch.empty((D.world,) + tuple(s.shape), dtype=s.dtype, device=D.dev)
    gi = torch.empty((D.world,) + tuple(i.shape), dtype=i.dtype, device=D.dev)
    D.dist.all_gather_into_tensor(gs, s.contiguous())
    D.dist.all_gather_into_tensor(gi, i.contiguous())
    cs = gs.permute(1, 0, 2).reshape(s.shape[0], -1)
    ci = gi.permute(1, 0, 2).reshape(s.shape[0], -1)
    ci_sort = torch.where(ci < 0, torch.full_like(ci, 1 << 40), ci)
    o1 = torch.argsort(ci_sort, dim=1, stable=True)
    cs, ci = torch.gather(cs, 1, o1), torch.gather(ci, 1, o1)
    o2 = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(cs, 1, o2), torch.gather(ci, 1, o2)


# ------------------------------------------------------------------------------------------ config 4 / config 2
def run_search_workload(args):
    import torch
    from brickbrain_rec_engine_b200 import ScoringEngine, ShardedSearch, shard_bounds, _lib as L
    from brickbrain_rec_engine_b200.sharded import slice_bounds

    D = Dist(torch)
    world, rank, dev = D.world, D.rank, D.dev
    n, b, k, d, f = args.n_items, args.batch, args.k, args.dim, args.cf_dim
    cfg4 = args.workload == "config4"
    flush_l2 = n * (d + f) * 2 // world < 256 * 1024 * 1024           # shard smaller than 2x L2: flush between steps

    # ---------------- resident catalog shard (synthetic, generated on the device)
    lo, hi = shard_bounds(n, world, rank)
    eng = ScoringEngine(hi - lo, d, f, device=D.local, id_offset=lo, k_max=128, max_batch=b)
    descs, theme_bitmaps, themes, pick = scenario_descriptors(b)
    eng.set_theme_sets(theme_bitmaps)
    if args.constraint_form == "indexed":
        distinct, cons_index = ScoringEngine.dedupe_constraints(descs)
    else:
        distinct, cons_index = descs, np.arange(b, dtype=np.int32)
    mu, theme_ids = load_synthetic_catalog(torch, eng, lo, hi, d, f, dev, themes)
    # per-user already-rated lists (recommendation_system.py:441-451): 48 global ids per query row
    n_rated = 0 if args.no_rated else 48
    rated_np = None
    if n_rated:
        rated_np = np.sort(np.random.default_rng(77).integers(0, n, size=(b, n_rated), dtype=np.int64), axis=1).astype(np.int32)
        eng.set_exclusion_lists(list(rated_np))
    ss = ShardedSearch(eng, exchange=args.exchange) if world > 1 else None
    exchange = ss.exchange if ss is not None else "none"

    # ---------------- query batches (two alternating), host-pinned and device-resident copies
    host, devq = [], []
    for s in range(2):
        g = torch.Generator(device="cpu").manual_seed(100 + s)
        q = torch.nn.functional.normalize(torch.randn(b, d, generator=g), dim=1)
        q = torch.nn.functional.normalize(0.8 * mu.cpu()[None, :] + 0.6 * q, dim=1) * 3.0       # un-normalized on purpose
        u = torch.randn(b, f, generator=g) / (f ** 0.5)
        self_id = torch.randint(0, n, (b,), generator=g, dtype=torch.int32)
        cons = torch.from_numpy(distinct.view(np.uint8).reshape(distinct.shape[0], -1).copy())
        cidx = torch.from_numpy(cons_index.copy())
        excl = torch.arange(b, dtype=torch.int32).reshape(b, 1)
        hq = dict(q=q.pin_memory(), u=u.pin_memory(), self_id=self_id.pin_memory(), cons=cons.pin_memory(),
                  cidx=cidx.pin_memory(), excl=excl.pin_memory())
        host.append(hq)
        devq.append({kk: vv.to(dev) for kk, vv in hq.items()})
    indexed = args.constraint_form == "indexed"
    s_lo, s_hi = slice_bounds(b, world, rank)
    n_out = b if world == 1 else max(1, s_hi - s_lo)
    out_s = torch.empty((n_out, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((n_out, k), dtype=torch.int32, device=dev)
    host_out_s = torch.empty((b, k), dtype=torch.float32).pin_memory()
    host_out_i = torch.empty((b, k), dtype=torch.int32).pin_memory()
    l2_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev) if flush_l2 else None

    def kwargs_of(qd, to_np=False):
        cv = (lambda t: t.numpy()) if to_np else (lambda t: t)
        kw = dict(w_content=0.4, w_cf=0.6, self_id=cv(qd["self_id"]), precision=args.precision, q_emb=cv(qd["q"]), q_cf=cv(qd["u"]))
        kw["constraints"] = qd["cons"].numpy().view(L.CONSTRAINT_DTYPE).reshape(-1) if to_np else qd["cons"]
        if indexed:
            kw["constraint_index"] = cv(qd["cidx"])
        if n_rated:
            kw["excl_lists"] = cv(qd["excl"])
        return kw

    pipelined = world > 1 and exchange == "p2p" and not args.no_pipeline and not flush_l2
    pipe_out = [(torch.empty((n_out, k), dtype=torch.float32, device=dev), torch.empty((n_out, k), dtype=torch.int32, device=dev))
                for _ in range(2)] if pipelined else None
    pipe_state = {"prev": None}

    def step_device(i):
        """Inputs resident in HBM; results stay on the device (world > 1: the rank's slice of the queries)."""
        kw = kwargs_of(devq[i & 1])
        if world == 1:
            return eng.search(k, out=(out_s, out_i), **kw)
        if pipelined:
            # back-to-back collective searches: step i's owner-side merge is launched inside step i+1 (after its prologue,
            # before its scan); pipe_flush() finishes the last one -- K steps still run K scatters and K merges
            eng.sharded_scatter(k, collect_previous_into=pipe_state["prev"], **kw)
            pipe_state["prev"] = pipe_out[i & 1]
            return pipe_out[i & 1]
        if exchange == "p2p":
            return eng.search(k, sharded=True, out=(out_s[:s_hi - s_lo], out_i[:s_hi - s_lo]), **kw)
        return ss.search(k, result="slice", **kw)

    def pipe_flush():
        if pipelined and pipe_state["prev"] is not None:
            eng.sharded_collect(k, max(0, s_hi - s_lo), out=pipe_state["prev"])
            pipe_state["prev"] = None

    def step_e2e(i):
        """The user-facing call with HOST buffers: pinned inputs in, results read back into host memory."""
        kw = kwargs_of(host[i & 1], to_np=True)
        if world == 1:
            return eng.search(k, out=(host_out_s.numpy(), host_out_i.numpy()), **kw)
        # N > 1: every rank uploads 1/world of the replicated batch (all-gathered over NVLink), searches its shard and
        # reads back the merged results of ITS slice of the queries
        return ss.search(k, result="slice", host_out=(host_out_s, host_out_i), **kw)

    # ---------------- device-resident timing (`value`)
    sampler = ClockSampler(D.local)
    if rank == 0:
        sampler.start()                 # nvidia-smi needs a moment to start: launch it before the warm-up
    # small catalogs finish a step in well under a millisecond: W such steps do not bring the SM clocks out of idle
    # (measured: the first process of a session ran 15% slower than the next), so they get 600 more untimed steps
    for i in range(args.warmup + (600 if flush_l2 else 0)):
        step_device(i)
    pipe_flush()
    D.barrier()
    eng.set_profiling(True)
    launches0 = eng.info()["launches"]
    D.barrier()
    sampler.mark()
    if flush_l2:
        # small catalog: flush L2 between steps, time every step with its own pair of events
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            l2_buf.fill_(i & 0xff)
            evs[i][0].record()
            step_device(i)
            evs[i][1].record()
        D.barrier()
        ms_total = D.max(sum(a.elapsed_time(bb) for a, bb in evs))
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            last_out = step_device(i)
        pipe_flush()
        e1.record()
        D.barrier()
        ms_total = D.max(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    info = eng.info()
    eng.set_profiling(False)
    launches = info["launches"] - launches0
    scan_ms = info["scan_ms"] / max(1, info["scan_launches"])
    ms_per_step = ms_total / args.steps
    value = b / (ms_per_step / 1e3)
    appends = info["last_appends"]
    pipe_match = None
    if pipelined:
        # the last pipelined step against the same batch searched and collected in one call (fraction of identical rows)
        ref_s, ref_i = eng.search(k, sharded=True, **kwargs_of(devq[(args.steps - 1) & 1]))
        torch.cuda.synchronize()
        rows = max(1, s_hi - s_lo)
        same = (last_out[1][:rows].long().sort(dim=1).values == ref_i[:rows].long().sort(dim=1).values).all(dim=1).float().mean()
        pipe_match = -D.max(-float(same.item()))

    # ---------------- end-to-end timing through the host-buffer API (`e2e`)
    for i in range(min(2, args.warmup)):
        step_e2e(i)
    D.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        if flush_l2:
            l2_buf.fill_(i & 0xff)
        step_e2e(i)
    D.barrier()
    e2e_ms = D.max((time.perf_counter() - t0) * 1e3) / args.steps
    h2d = b * (d + f) * 4 + b * 4 + distinct.shape[0] * L.CONSTRAINT_DTYPE.itemsize + (b * 4 if indexed else 0) + (b * 4 if n_rated else 0)
    d2h = b * k * 8

    # ---------------- the run proves what it timed
    check = None
    if not args.no_check:
        S = min(args.check_queries, b)
        sel = np.sort(np.random.default_rng(11).choice(b, S, replace=False))
        sel_t = torch.from_numpy(sel).to(dev)
        qd = devq[0]
        kw = kwargs_of(qd)
        full = {}
        for prec in (args.precision, "fp32"):
            kw_p = dict(kw, precision=prec)
            if prec == "fp32":                      # exact engine path on the sampled rows only
                kw_p = {kk: (vv[sel_t] if torch.is_tensor(vv) and vv.shape[0] == b and kk != "constraints" else vv) for kk, vv in kw_p.items()}
                if not indexed:
                    kw_p["constraints"] = kw["constraints"][sel_t]
            if world == 1:
                s_, i_ = eng.search(k, **kw_p)
            else:
                s_, i_ = ss.search(k, result="full", **kw_p)
            full[prec] = (s_, i_) if prec == "fp32" else (s_[sel_t], i_[sel_t])
        cls = torch.from_numpy(cons_index[sel]).to(dev)
        rated_t = torch.from_numpy(rated_np[sel]).to(dev) if n_rated else None
        ex_s, ex_i, viol = exact_bruteforce(torch, D, lo, hi, d, f, mu, theme_ids, qd["q"][sel_t], qd["u"][sel_t], qd["self_id"][sel_t],
                                            cls, distinct, theme_bitmaps, rated_t, k, [full[args.precision][1], full["fp32"][1]])
        ex_s, ex_i = merge_exact_over_ranks(torch, D, ex_s, ex_i, k)
        viol = int(D.sum(viol))
        got_i = full[args.precision][1].long()
        exi = ex_i.cpu().numpy()
        gi = got_i.cpu().numpy()
        recall = float(np.mean([len(set(gi[r][gi[r] >= 0]) & set(exi[r][exi[r] >= 0])) / max(1, int((exi[r] >= 0).sum())) for r in range(S)]))
        f_s, f_i = full["fp32"][0].cpu().numpy(), full["fp32"][1].cpu().numpy().astype(np.int64)
        exs = ex_s.cpu().numpy()
        # an id mismatch counts when the two items' exact scores differ by more than fp32 summation noise (not a tie)
        mism = 0
        for r in range(S):
            for c in np.nonzero(f_i[r] != exi[r])[0]:
                if not (np.isfinite(exs[r, c]) and abs(float(f_s[r, c]) - float(exs[r, c])) <= 4e-6 * max(1.0, abs(float(exs[r, c])))):
                    mism += 1
        fin = np.isfinite(exs) & (f_i == exi)
        rel = float(np.max(np.abs(f_s[fin] - exs[fin]) / np.maximum(np.abs(exs[fin]), 1e-3))) if fin.any() else 0.0
        check = {"queries": int(S), "recall_at_k": recall, "mask_violations": viol, "id_mismatch_fp32": int(mism),
                 "max_score_rel_err_fp32": rel,
                 "how": "sampled query rows vs an independent exact brute force (torch fp32 matmul + torch mask predicates over the "
                        "regenerated catalog, all shards; self and rated exclusion applied); id_mismatch_fp32 compares the engine's "
                        "exact-fp32 path, position by position, ignoring swaps of items whose exact scores tie within 4e-6"}

    if rank != 0:
        D.close()
        return
    pk = peaks()
    flops = 2.0 * b * (hi - lo) * (d + f)                      # algorithmic flop of one scan launch on this rank
    achieved = flops / (scan_ms * 1e-3) / 1e12 if scan_ms > 0 else 0.0
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "scan_dram_bytes.json")
    if os.path.exists(tfile) and cfg4:
        try:
            traffic = json.load(open(tfile)).get("per_n_gpus", {}).get(str(world))     # one ncu capture per N; None when not captured
        except Exception:
            traffic = None
    name = "config4" if cfg4 else "config2"
    workload = (f"{name}: {n} items x ({d}+{f}) {('bf16' if args.precision != 'fp32' else 'fp32')} catalog, batch {b} hybrid queries "
                f"(0.4 content cosine + 0.6 CF dot), per-query hard constraints (light/medium/heavy/none mix of the reference's "
                f"scenarios; passed as {distinct.shape[0]} constraint sets" + (" + per-query index" if indexed else "") + "), "
                + (f"per-user already-rated exclusion lists ({n_rated} ids each), " if n_rated else "") + f"self exclusion, top-{k}")
    line = {
        "metric": METRIC if cfg4 else f"hybrid top-{k} queries/sec @{n}x{d} catalog (BASELINE config 2)", "value": value, "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if args.precision != "fp32" else "f32", "data": "synthetic",
        "config": {"workload": workload, "n_items": n, "batch": b, "k": k, "dim": d, "cf_dim": f, "precision": args.precision,
                   "constraint_form": args.constraint_form,
                   "parallelism": f"row-sharded catalog x{world}" + ("" if world == 1 else
                                  (" + union thresholds and owner inboxes over NVLink peer memory (fused into the scan / merge kernels); each rank keeps the results of its query slice"
                                   + ("; back-to-back steps are pipelined (the owner-side merge of step s runs inside step s+1, after its prologue and before its scan)" if pipelined else "")
                                   if exchange == "p2p" else " + NCCL all-gather top-k merge")),
                   "l2_policy": ("L2 flushed between steps (256 MB write), every step timed with its own CUDA events" if flush_l2 else
                                 f"catalog shard ({(hi - lo) * (d + f) * 2 / 1e9:.2f} GB bf16) is larger than L2 (126 MB); query batches alternate")},
        "clocks": clocks,
        "e2e": {"value": b / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "scan_tc_kernel" if args.precision != "fp32" else "scan_fp32_kernel",
                     "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16_sustained"], "frac_of_burst": achieved / pk["bf16_burst"],
                     "peak_source": pk["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                     "kernel_ms": scan_ms, "kernel_share_of_step": scan_ms / ms_per_step,
                     "algorithmic_flop_per_launch": flops, "traffic": traffic,
                     "appends_per_launch": int(appends)},
    }
    if check is not None:
        if pipe_match is not None:
            check["pipelined_rows_equal_to_plain_collective_search"] = pipe_match
        line["check"] = check
        line["recall_at_k"] = check["recall_at_k"]
        line["mask_violations"] = check["mask_violations"]
        line["id_mismatch_fp32"] = check["id_mismatch_fp32"]
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args)
    emit(line)
    D.close()


# ------------------------------------------------------------------------------------------ config 3: item-item CF
def run_item_item(args):
    import torch
    from brickbrain_rec_engine_b200 import ScoringEngine
    D = Dist(torch)
    if D.rank != 0:
        D.close()
        return
    n_items, n_users, k = args.n_items, args.users, args.k
    rng = np.random.default_rng(3)
    w = 1.0 / np.arange(1, n_items + 1) ** 0.8
    w /= w.sum()
    perm = rng.permutation(n_items)
    counts = np.clip(rng.poisson(20, size=n_users), 1, None)
    draws = perm[rng.choice(n_items, size=int(counts.sum()), p=w)].astype(np.int64)
    rows = np.repeat(np.arange(n_users, dtype=np.int64), counts)
    key = np.unique(rows * n_items + draws)                        # duplicate (user, item) draws collapse
    rows, cols = key // n_items, (key % n_items).astype(np.int32)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n_users), out=indptr[1:])
    vals = rng.integers(1, 6, size=cols.shape[0]).astype(np.float32)
    deg = np.diff(indptr)
    macs = float((deg.astype(np.float64) ** 2).sum())
    eng = ScoringEngine(1, 1, 0, device=D.local)
    d_ptr, d_idx, d_val = (torch.from_numpy(x).to(D.dev) for x in (indptr, cols, vals))
    sampler = ClockSampler(D.local)
    sampler.start()
    t_w = time.perf_counter()
    n_w = 0
    while n_w < args.warmup or time.perf_counter() - t_w < 2.0:      # (a 3 ms step: warm up by time -- a fresh process measured 15% slow for its first second)
        eng.item_item_topk(d_ptr, d_idx, d_val, n_items, k)
        torch.cuda.synchronize()
        n_w += 1
    torch.cuda.synchronize()
    launches0 = eng.info()["launches"]
    sampler.mark()
    t = []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.item_item_topk(d_ptr, d_idx, d_val, n_items, k, device_out=True)
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    ms = float(np.mean(t))
    launches = eng.info()["launches"] - launches0
    eng.item_item_topk(indptr, cols, vals, n_items, k)                # (untimed: first use allocates the staging buffers)
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 4)):
        eng.item_item_topk(indptr, cols, vals, n_items, k)            # host CSR in, host lists out
    e2e_ms = (time.perf_counter() - t0) * 1e3 / max(1, args.steps // 4)
    del out
    pk = peaks()
    alg_bytes = cols.shape[0] * 8.0 + indptr.shape[0] * 8.0 + n_items * k * 8.0
    line = {"metric": f"item-item cosine top-{k} @{n_users} users x {n_items} items (BASELINE config 3)", "value": ms, "unit": "ms",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "replicas only",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config3: Zipf(0.8) item popularity, ~20 ratings per user, {cols.shape[0]} ratings, "
                                   f"{macs:.3e} co-rating multiply-adds, top-{k} neighbours per item", "n_items": n_items, "users": n_users, "k": k,
                       "l2_policy": "rating CSR (~0.24 GB) is larger than L2; nothing is kept between steps"},
            "clocks": sampler.stop(),
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(cols.shape[0] * 8 + indptr.shape[0] * 8), "d2h_bytes_per_step": int(n_items * k * 8)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "ii_accumulate_kernel", "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": alg_bytes / (ms * 1e-3) / 1e9 / pk["hbm"], "traffic": None,
                         "algorithmic_bytes_per_launch": alg_bytes, "macs": macs, "gmacs_per_s": macs / (ms * 1e-3) / 1e9,
                         "note": "achieved = (8 B per rating + indptr + N*k*8 B out) / whole item-item pass; the pass is latency / issue bound (row gathers from L2 + two dense selection passes per item in shared memory), not bound by these bytes -- profiles/r2_item_item_ncu.md"}}
    emit(line)
    D.close()


# ------------------------------------------------------------------------------------------ config 5: bulk generation
def run_bulk(args):
    """1M users x 1M items, hybrid top-100: users sharded over the ranks, the catalog REPLICATED on every GPU (0.9 GB
    bf16 + 1.8 GB fp32), no data-path collective; results stream to host through the pinned double-buffered
    sink.  Total work is fixed -> "strong"."""
    import torch
    from brickbrain_rec_engine_b200 import ScoringEngine
    from brickbrain_rec_engine_b200.bulk import BulkSink, bulk_recommend, user_shard
    D = Dist(torch)
    n, b, k, d, f = args.n_items, args.batch, args.k, args.dim, args.cf_dim
    users = args.users
    u_lo, u_hi = user_shard(users, D.world, D.rank)
    eng = ScoringEngine(n, d, f, device=D.local, k_max=128, max_batch=b)
    themes = themes_table()
    load_synthetic_catalog(torch, eng, 0, n, d, f, D.dev, themes)
    g = torch.Generator(device=D.dev).manual_seed(500 + D.rank)
    uf = torch.randn(u_hi - u_lo, f, device=D.dev, generator=g) / (f ** 0.5)
    liked = torch.randint(0, n, (u_hi - u_lo,), device=D.dev, generator=g, dtype=torch.int32)
    sink = BulkSink(k, batch=b, device=D.dev, keep=False)
    sampler = ClockSampler(D.local)
    if D.rank == 0:
        sampler.start()
    warm = min(u_hi - u_lo, b * max(1, args.warmup))
    bulk_recommend(eng, uf[:warm], liked[:warm], k, batch=b, precision=args.precision, sink=sink, user_offset=u_lo)
    sink.flush()
    D.barrier()
    eng.set_profiling(True)
    launches0 = eng.info()["launches"]
    D.barrier()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    bulk_recommend(eng, uf, liked, k, batch=b, precision=args.precision, sink=sink, user_offset=u_lo)
    e1.record()
    sink.flush()
    D.barrier()
    wall_ms = D.max((time.perf_counter() - t0) * 1e3)
    dev_ms = D.max(e0.elapsed_time(e1))
    clocks = sampler.stop() if D.rank == 0 else None
    info = eng.info()
    launches = info["launches"] - launches0
    scan_ms_total = info["scan_ms"]
    if D.rank != 0:
        D.close()
        return
    pk = peaks()
    n_batches = -(-(u_hi - u_lo) // b)
    flops = 2.0 * (u_hi - u_lo) * n * (d + f)
    achieved = flops / (scan_ms_total * 1e-3) / 1e12 if scan_ms_total > 0 else 0.0
    line = {"metric": f"bulk hybrid top-{k} users/sec @{users} users x {n} items (BASELINE config 5)", "value": users / (dev_ms / 1e3), "unit": "users/s",
            "n_gpus": D.world, "steps": n_batches, "warmup": args.warmup, "ms_per_step": dev_ms / n_batches, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": f"config5: {users} users x {n} items x ({d}+{f}), hybrid (0.4 cosine of the liked set + 0.6 CF dot), self exclusion, top-{k}; "
                                   f"users sharded over {D.world} rank(s) in batches of {b}, catalog replicated per GPU, results to host through a pinned double-buffered sink",
                       "n_items": n, "users": users, "batch": b, "k": k, "precision": args.precision,
                       "parallelism": f"user-sharded x{D.world}, replicated catalog, no collective",
                       "l2_policy": f"catalog ({n * (d + f) * 2 / 1e9:.2f} GB bf16) is larger than L2; every batch has new users"},
            "clocks": clocks,
            "e2e": {"value": users / (wall_ms / 1e3), "unit": "users/s", "ms_total": wall_ms, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": int(b * k * 8), "note": "wall clock incl. the device->host stream of every result batch (pinned double buffer); the per-user inputs (factors, liked set) are device-resident state of the bulk job"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "scan_tc_kernel", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["bf16_sustained"], "frac_of_burst": achieved / pk["bf16_burst"], "kernel_ms_total": scan_ms_total,
                         "kernel_share_of_job": scan_ms_total / dev_ms, "algorithmic_flop_per_rank": flops, "traffic": None}}
    emit(line)
    D.close()


# ------------------------------------------------------------------------------------------ CPU arm
def _blas_threads(n):
    """Pin the BLAS / OpenMP pools explicitly: torchrun exports OMP_NUM_THREADS=1, which used to throttle this arm
    silently.  Returns (context manager, description)."""
    from threadpoolctl import threadpool_limits
    return threadpool_limits(limits=n)


def _blas_info():
    from threadpoolctl import threadpool_info
    return [{"api": p.get("user_api"), "lib": p.get("internal_api"), "threads": p.get("num_threads")} for p in threadpool_info()]


def _cpu_sample(args):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import brick_oracle as bo
    import synth
    ns, bs = min(args.cpu_sample_items, args.n_items), min(args.cpu_sample_queries, args.batch)
    d, f = args.dim, args.cf_dim
    emb = synth.embeddings(ns, d, 3, "aniso")
    itf = synth.cf_factors(ns, f, 4)
    e_cat = np.ascontiguousarray(np.concatenate([emb, itf], axis=1))
    q = bo.l2_normalize_rows(synth.embeddings(bs, d, 5, "aniso") * 3.0)
    u = synth.cf_factors(bs, f, 6)
    q_cat = np.ascontiguousarray(np.concatenate([np.float32(0.4) * q, np.float32(0.6) * u], axis=1))
    themes = synth.themes_table()
    parts, year, theme = synth.metadata(ns, 8, themes)
    # the SAME per-query constraint mix as the GPU arm: four sets -> four [N] masks + a per-query index, plus the
    # per-user rated lists and the self exclusion
    class_valid = np.stack([bo.constraint_mask(bo.compile_constraints(c, themes, 2026), parts, year, theme) for c in scenario_mix()])
    rng = np.random.default_rng(5)
    class_index = rng.integers(0, class_valid.shape[0], size=bs).astype(np.int32)
    rated = None if args.no_rated else np.sort(rng.integers(0, ns, size=(bs, 48)), axis=1).astype(np.int64)
    self_ids = rng.integers(0, ns, size=bs).astype(np.int64)
    return bo, q_cat, e_cat, class_valid, class_index, rated, self_ids, ns, bs


def reference_style_config1(cores: int, n_queries: int = 12):
    """BASELINE.md section 4 item 1 at config 1 (25,216 x 384, one query per call, top-10 + filter): the reference's
    per-query execution (sklearn cosine over the whole catalog, full argsort, Python filter walk, np.dot + Python loop
    + list.sort, dict-union blend; oracle restatement ``reference_style_query``), timed with ONE BLAS thread (the
    reference's deployed setting, docker-compose.yml:52-58) and with all cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import brick_oracle as bo
    import synth
    n, d, f = 25_216, 384, 64
    emb = synth.embeddings(n, d, 13, "aniso")
    itf = synth.cf_factors(n, f, 14)
    themes = synth.themes_table()
    parts, year, theme = synth.metadata(n, 15, themes)
    valid = bo.constraint_mask(bo.compile_constraints([("pieces_max", 1000), ("year_min", 2010)], themes, 2026), parts, year, theme)
    valid_ids = set(np.nonzero(valid)[0].tolist())
    qs = synth.embeddings(n_queries, d, 16, "aniso")
    us = synth.cf_factors(n_queries, f, 17)
    out = {}
    for label, nt in (("1_thread", 1), ("all_cores", cores)):
        with _blas_threads(nt):
            bo.reference_style_query(qs[0], us[0], emb, itf, valid_ids, 0.4, 0.6, 10)
            t0 = time.perf_counter()
            for j in range(n_queries):
                bo.reference_style_query(qs[j], us[j], emb, itf, valid_ids, 0.4, 0.6, 10)
            out[f"ms_per_query_{label}"] = (time.perf_counter() - t0) * 1e3 / n_queries
    out["queries"] = n_queries
    out["what"] = "config 1: 25,216 x (384+64), one hybrid query per call, top-10 with a valid-set filter, reference-style per-query execution"
    return out


def cpu_baseline(args, steps: int = 1, warmup: int = 1):
    """Oracle port (blocked sgemm + per-class masks + exclusions + running top-k, all BLAS threads pinned explicitly)
    on a bounded sample of the same workload; queries/s is extrapolated linearly in catalog size to the benchmark's N."""
    bo, q_cat, e_cat, class_valid, class_index, rated, self_ids, ns, bs = _cpu_sample(args)
    cores = host_cores()
    with _blas_threads(cores):
        info = _blas_info()
        for _ in range(max(1, warmup)):
            bo.batched_cpu_topk(q_cat[:64], e_cat[:65536], args.k, class_valid=class_valid[:, :65536], class_index=class_index[:64])
        t0 = time.perf_counter()
        for _ in range(steps):
            bo.batched_cpu_topk(q_cat, e_cat, args.k, class_valid=class_valid, class_index=class_index, exclusions=rated, self_ids=self_ids)
        dt = (time.perf_counter() - t0) / steps
    qps_sample = bs / dt
    threads = max([p["threads"] or 1 for p in info if p["api"] == "blas"] + [1])
    res = {"value": qps_sample * ns / args.n_items, "unit": UNIT, "cores": cores, "blas_threads": threads, "kind": "port",
           "sample": f"{bs} queries x {ns} items x {args.dim + args.cf_dim} fp32 (sgemm + the GPU arm's 4 constraint-set masks by per-query index + "
                     f"rated / self exclusion + running top-{args.k}) took {dt:.2f} s per step; extrapolated linearly in N to {args.n_items} items",
           "sample_seconds": dt, "threadpools": info,
           "env_omp_num_threads": os.environ.get("OMP_NUM_THREADS"), "os_cpu_count": os.cpu_count()}
    try:
        res["reference_style_config1"] = reference_style_config1(cores)
    except Exception as e:      # pragma: no cover
        res["reference_style_config1"] = {"error": str(e)}
    return res


def run_reference(args):
    """`--impl reference`: the CPU arm.  The reference is Python and cannot travel to the GPU box, so this
    times the oracle port of its scoring path (kind "port") with all host threads; each step is a bounded
    sample of the workload (see `sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
    ns, bs = min(args.cpu_sample_items, args.n_items), min(args.cpu_sample_queries, args.batch)
    value = cb["value"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["sample_seconds"] * 1e3 * (args.batch / bs) * (args.n_items / ns),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the GPU arm's workload (same catalog shape, batch, constraint mix, rated lists, k), computed in fp32 on the host
            "config": {"workload": (f"config4: {args.n_items} items x ({args.dim}+{args.cf_dim}) fp32 catalog, batch {args.batch} hybrid queries "
                                    f"(0.4 content cosine + 0.6 CF dot), per-query hard constraints (light/medium/heavy/none mix of the reference's "
                                    f"scenarios; passed as 4 constraint sets + per-query index), per-user already-rated exclusion lists "
                                    f"(48 ids each), self exclusion, top-{args.k}"),
                       "n_items": args.n_items, "batch": args.batch, "k": args.k, "dim": args.dim, "cf_dim": args.cf_dim,
                       "precision": "fp32", "constraint_form": "indexed", "parallelism": f"host threads x{cb.get('cores')}",
                       "l2_policy": "n/a (CPU)"},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_JSON_FD = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything else (library banners such as NCCL's
    version line, warnings) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)          # keep the real stdout for the JSON line ...
    os.dup2(2, 1)                 # ... and send every other write to fd 1 (C libraries included) to stderr
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "config3":
        run_item_item(args)
    elif args.workload == "config5":
        run_bulk(args)
    else:
        run_search_workload(args)


if __name__ == "__main__":
    main()
