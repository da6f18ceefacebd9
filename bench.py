#!/usr/bin/env python
"""Benchmark of the candidate-scoring hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, all host threads)

A step = one pass of the hot path over one batch: 8192 hybrid queries
(w_c*cos + w_f*(u.v), per-query hard-constraint descriptor, self exclusion,
top-100) against the whole 10M x (384+64) catalog (BASELINE config 4).  With
N>1 the catalog is row-sharded over the ranks (one process per GPU) and the
per-rank candidates are merged through one NCCL all-gather; total work per step is
fixed, so scaling is "strong".

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
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "hybrid top-k queries/sec @10Mx384 catalog"
UNIT = "queries/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--n-items", type=int, default=10_000_000)
    p.add_argument("--batch", type=int, default=8192)
    p.add_argument("--k", type=int, default=100)
    p.add_argument("--dim", type=int, default=384)
    p.add_argument("--cf-dim", type=int, default=64)
    p.add_argument("--precision", default="bf16_rescore", choices=["fp32", "bf16", "bf16_rescore"])
    p.add_argument("--constraint-form", default="indexed", choices=["indexed", "per-query"],
                   help="indexed: distinct constraint sets + per-query index (class bitmaps); per-query: one descriptor per query")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--share-thresholds", action="store_true",
                   help="N>1: shards raise each other's per-query running thresholds through NVLink peer memory "
                        "(bb_import_peer_thresholds); measured SLOWER on 8 B200s (813k vs 869k queries/s), off by default")
    p.add_argument("--cpu-sample-items", type=int, default=1_000_000)
    p.add_argument("--cpu-sample-queries", type=int, default=1024)
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_sustained": d["bf16_tflops_sustained"], "bf16_burst": d["bf16_tflops"], "hbm": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


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
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        for ts, l in self.lines:
            if ts < t_mark:
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def scenario_descriptors(batch: int, seed: int = 5):
    """Per-query bb_constraint records + theme bitmaps, drawn from the reference's light / medium / heavy
    scenario mix (tests/unit/test_recommendations.py:511-541)."""
    import synth
    from brickbrain_rec_engine_b200 import ConstraintCompiler, ConstraintType, HardConstraint, _lib as L
    themes = synth.themes_table()
    comp = ConstraintCompiler(themes, theme_bits=1024, now_year=2026)
    T = ConstraintType
    mix = [
        [HardConstraint(T.PIECES_MAX, 1000), HardConstraint(T.YEAR_MIN, 2010)],
        [HardConstraint(T.PIECES_MIN, 100), HardConstraint(T.PIECES_MAX, 800),
         HardConstraint(T.THEMES_REQUIRED, ["City", "Creator", "Star Wars"]), HardConstraint(T.YEAR_MIN, 2015)],
        [HardConstraint(T.PRICE_MAX, 200.0), HardConstraint(T.PRICE_MIN, 30.0), HardConstraint(T.PIECES_MIN, 200),
         HardConstraint(T.PIECES_MAX, 1500), HardConstraint(T.AGE_MIN, 8),
         HardConstraint(T.THEMES_REQUIRED, ["City", "Creator", "Technic"]),
         HardConstraint(T.THEMES_EXCLUDED, ["Duplo", "Baby"]), HardConstraint(T.YEAR_MIN, 2018),
         HardConstraint(T.COMPLEXITY_MAX, "moderate")],
        [],
    ]
    compiled = [comp.compile(c) for c in mix]
    bitmaps, index = [], {}
    descs = np.stack([L.no_constraint() for _ in range(batch)])
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, len(mix), size=batch)
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
    return descs, np.stack(bitmaps), themes


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


def run_ours(args):
    import torch
    import torch.distributed as dist
    from brickbrain_rec_engine_b200 import ScoringEngine, shard_bounds, _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    n, b, k, d, f = args.n_items, args.batch, args.k, args.dim, args.cf_dim

    # ---------------- resident catalog shard (synthetic, generated on the device)
    lo, hi = shard_bounds(n, world, rank)
    eng = ScoringEngine(hi - lo, d, f, device=local, id_offset=lo, k_max=128, max_batch=b)
    # row-sharded catalog: per-shard search + NCCL all-gather of the candidates + merge (ShardedSearch); optionally the
    # shards' per-query running thresholds become one logical array over NVLink peer memory
    ss = None
    if world > 1:
        from brickbrain_rec_engine_b200 import ShardedSearch
        ss = ShardedSearch(eng, share_thresholds=args.share_thresholds)
    share = ss is not None and ss.shared
    descs, theme_bitmaps, themes = scenario_descriptors(b)
    eng.set_theme_sets(theme_bitmaps)
    # queries carry (index into) the batch's distinct constraint sets; --distinct-constraints per-query disables it
    if args.constraint_form == "indexed":
        distinct, cons_index = ScoringEngine.dedupe_constraints(descs)
    else:
        distinct, cons_index = descs, np.arange(b, dtype=np.int32)
    theme_ids = torch.tensor(sorted(themes), device=dev, dtype=torch.int32)
    gmu = torch.Generator(device=dev).manual_seed(1234)
    mu = torch.nn.functional.normalize(torch.randn(d, device=dev, generator=gmu), dim=0)
    chunk = 1 << 20
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        x, v, parts, year, theme = gen_catalog_chunk(torch, dev, c0, c1, d, f, mu, theme_ids, seed=7)
        eng.load_catalog(emb=x, item_factors=v, num_parts=parts, year=year, theme_id=theme, row_begin=c0 - lo)
    torch.cuda.synchronize()
    del x, v, parts, year, theme

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
        hq = dict(q=q.pin_memory(), u=u.pin_memory(), self_id=self_id.pin_memory(), cons=cons.pin_memory(),
                  cidx=cidx.pin_memory())
        host.append(hq)
        devq.append({kk: vv.to(dev) for kk, vv in hq.items()})
    indexed = args.constraint_form == "indexed"
    out_s = torch.empty((b, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((b, k), dtype=torch.int32, device=dev)
    gather_s = torch.empty((world, b, k), dtype=torch.float32, device=dev) if world > 1 else None
    gather_i = torch.empty((world, b, k), dtype=torch.int32, device=dev) if world > 1 else None
    host_out_s = torch.empty((b, k), dtype=torch.float32).pin_memory()
    host_out_i = torch.empty((b, k), dtype=torch.int32).pin_memory()

    def step_device(i):
        qd = devq[i & 1]
        eng.search(k, q_emb=qd["q"], q_cf=qd["u"], w_content=0.4, w_cf=0.6, constraints=qd["cons"],
                   constraint_index=qd["cidx"] if indexed else None, self_id=qd["self_id"], precision=args.precision,
                   out=(out_s, out_i), share_thresholds=share)
        if world > 1:
            dist.all_gather_into_tensor(gather_s.view(-1), out_s.view(-1))
            dist.all_gather_into_tensor(gather_i.view(-1), out_i.view(-1))
            return eng.merge_topk(gather_s, gather_i, k)
        return out_s, out_i

    def step_e2e(i):
        """The user-facing call with HOST buffers: pinned inputs are copied in by bb_search, the result is read back."""
        qh = host[i & 1]
        cons_np = qh["cons"].numpy().view(L.CONSTRAINT_DTYPE).reshape(-1)
        cidx_np = qh["cidx"].numpy() if indexed else None
        if world == 1:
            s, ids = eng.search(k, q_emb=qh["q"].numpy(), q_cf=qh["u"].numpy(), w_content=0.4, w_cf=0.6,
                                constraints=cons_np, constraint_index=cidx_np, self_id=qh["self_id"].numpy(),
                                precision=args.precision, out=(host_out_s.numpy(), host_out_i.numpy()))
            return s, ids
        # N > 1: the user-facing call is ShardedSearch.search with HOST arrays -- every rank uploads 1/world of the
        # replicated batch, the slices are all-gathered over NVLink, then shard search + candidate all-gather + merge;
        # the merged result is read back into pinned host memory
        return ss.search(k, host_out=(host_out_s, host_out_i), q_emb=qh["q"].numpy(), q_cf=qh["u"].numpy(), w_content=0.4,
                         w_cf=0.6, constraints=cons_np, constraint_index=cidx_np, self_id=qh["self_id"].numpy(),
                         precision=args.precision)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident timing (`value`)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # nvidia-smi needs a moment to start: launch it before the warm-up
    for i in range(args.warmup):
        step_device(i)
    barrier()
    eng.set_profiling(True)
    launches0 = eng.info()["launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    e0.record()
    for i in range(args.steps):
        step_device(i)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    info = eng.info()
    eng.set_profiling(False)
    launches = info["launches"] - launches0 + (args.steps if world > 1 else 0)
    scan_ms = info["scan_ms"] / max(1, info["scan_launches"])
    ms_per_step = ms_total / args.steps
    value = b / (ms_per_step / 1e3)

    # ---------------- end-to-end timing through the host-buffer API (`e2e`)
    for i in range(min(2, args.warmup)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    h2d = b * (d + f) * 4 + b * 4 + distinct.shape[0] * L.CONSTRAINT_DTYPE.itemsize + (b * 4 if indexed else 0)
    d2h = b * k * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    flops = 2.0 * b * (hi - lo) * (d + f)                      # algorithmic flop of one scan launch on this rank
    achieved = flops / (scan_ms * 1e-3) / 1e12 if scan_ms > 0 else 0.0
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "scan_dram_bytes.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if args.precision != "fp32" else "f32", "data": "synthetic",
        "config": {"workload": f"config4: {n} items x ({d}+{f}) {('bf16' if args.precision != 'fp32' else 'fp32')} catalog, "
                               f"batch {b} hybrid queries (0.4 content cosine + 0.6 CF dot), per-query hard constraints "
                               f"(light/medium/heavy/none mix of the reference's scenarios; passed as "
                               f"{distinct.shape[0]} constraint sets" + (" + per-query index" if indexed else "") +
                               f"), self exclusion, top-{k}",
                   "n_items": n, "batch": b, "k": k, "dim": d, "cf_dim": f, "precision": args.precision,
                   "constraint_form": args.constraint_form, "parallelism": f"row-sharded catalog x{world}" + (" + NCCL all-gather top-k merge" if world > 1 else "") + (" + NVLink-shared running thresholds" if share else ""),
                   "l2_policy": f"catalog shard ({(hi - lo) * (d + f) * 2 / 1e9:.2f} GB bf16) is larger than L2 (126 MB); "
                                "query batches alternate"},
        "clocks": clocks,
        "e2e": {"value": b / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "scan_tc_kernel" if args.precision != "fp32" else "scan_fp32_kernel",
                     "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16_sustained"], "frac_of_burst": achieved / pk["bf16_burst"],
                     "peak_source": pk["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                     "kernel_ms": scan_ms, "kernel_share_of_step": scan_ms / ms_per_step,
                     "algorithmic_flop_per_launch": flops, "traffic": traffic},
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_sample(args, rng_seed=0):
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
    spec = bo.compile_constraints([("pieces_max", 1000), ("year_min", 2010)], themes, 2026)
    valid = bo.constraint_mask(spec, parts, year, theme)
    return bo, q_cat, e_cat, valid, ns, bs


def cpu_baseline(args):
    """Oracle port (blocked sgemm + mask + argpartition, all BLAS threads) on a bounded sample of the same
    workload; queries/s is extrapolated linearly in catalog size to the benchmark's N."""
    bo, q_cat, e_cat, valid, ns, bs = _cpu_sample(args)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    bo.batched_cpu_topk(q_cat[:64], e_cat[:65536], args.k, valid[:65536])          # warm-up
    t0 = time.perf_counter()
    bo.batched_cpu_topk(q_cat, e_cat, args.k, valid)
    dt = time.perf_counter() - t0
    qps_sample = bs / dt
    return {"value": qps_sample * ns / args.n_items, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{bs} queries x {ns} items x {args.dim + args.cf_dim} fp32 (sgemm + [N] mask + argpartition "
                      f"top-{args.k}) took {dt:.2f} s; extrapolated linearly in N to {args.n_items} items",
            "sample_seconds": dt}


def run_reference(args):
    """`--impl reference`: the CPU arm.  The reference is Python and cannot travel to the GPU box, so this
    times the oracle port of its scoring path (kind "port") with all host threads; each step is a bounded
    sample of the workload (see `sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bo, q_cat, e_cat, valid, ns, bs = _cpu_sample(args)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for _ in range(max(1, min(args.warmup, 1))):
        bo.batched_cpu_topk(q_cat[:128], e_cat, args.k, valid)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bo.batched_cpu_topk(q_cat, e_cat, args.k, valid)
    dt = (time.perf_counter() - t0) / args.steps
    value = (bs / dt) * ns / args.n_items
    sample = (f"each step = {bs} queries x {ns} items x {args.dim + args.cf_dim} fp32 (sgemm + mask + argpartition); "
              f"value extrapolated linearly in N to {args.n_items} items")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 * (args.batch / bs) * (args.n_items / ns),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config4: {args.n_items} items x ({args.dim}+{args.cf_dim}), batch {args.batch}, top-{args.k}",
                       "n_items": args.n_items, "batch": args.batch, "k": args.k},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
