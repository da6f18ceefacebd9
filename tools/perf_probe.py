"""Times the tcgen05 scan on a mid-size problem under the BB_DEBUG_FLAGS triage modes
(each in a subprocess because the flag is read per search from the environment)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(n, b, k, cons):
    import numpy as np
    import torch
    import brickbrain_rec_engine_b200 as bb
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    d, f = 384, 64
    eng = bb.ScoringEngine(n, d, f, max_batch=b)
    for lo in range(0, n, 1 << 20):
        hi = min(n, lo + (1 << 20))
        x = torch.nn.functional.normalize(torch.randn(hi - lo, d, device=dev, generator=g) + 0.9 * torch.ones(1, d, device=dev), dim=1)
        v = torch.randn(hi - lo, f, device=dev, generator=g) / 8
        parts = torch.randint(1, 3000, (hi - lo,), device=dev, dtype=torch.int32, generator=g)
        year = torch.randint(1950, 2026, (hi - lo,), device=dev, generator=g).to(torch.int16)
        theme = torch.randint(1, 778, (hi - lo,), device=dev, generator=g).to(torch.int16)
        eng.load_catalog(emb=x, item_factors=v, num_parts=parts, year=year, theme_id=theme, row_begin=lo)
    q = torch.nn.functional.normalize(torch.randn(b, d, device=dev, generator=g) + 0.9, dim=1)
    u = torch.randn(b, f, device=dev, generator=g) / 8
    c = None
    cidx = None
    if cons == 1:
        cn = bb.no_constraint()
        cn["parts_min"], cn["parts_max"] = 1, 1000
        c = torch.from_numpy(np.atleast_1d(cn).view(np.uint8).reshape(1, -1).copy()).to(dev)
    elif cons in (2, 3):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import bench
        descs, bitmaps, _, _ = bench.scenario_descriptors(b)
        eng.set_theme_sets(bitmaps)
        if cons == 3:
            descs, cidx_np = eng.dedupe_constraints(descs)
            cidx = torch.from_numpy(cidx_np).to(dev)
        c = torch.from_numpy(descs.view(np.uint8).reshape(descs.shape[0], -1).copy()).to(dev)
    for prec in ("bf16",):
        for _ in range(2):
            eng.search(k, q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=c, constraint_index=cidx, precision=prec)
        torch.cuda.synchronize()
        eng.set_profiling(True)
        for _ in range(3):
            eng.search(k, q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=c, constraint_index=cidx, precision=prec)
        torch.cuda.synchronize()
        i = eng.info()
        ms = i["scan_ms"] / i["scan_launches"]
        tf = 2.0 * b * n * 448 / (ms * 1e-3) / 1e12
        print(f"flags={os.environ.get('BB_DEBUG_FLAGS','0')} cg={i['cta_group']} cons={cons} n={n} b={b} k={k}: scan {ms:.2f} ms = {tf:.1f} TFLOP/s "
              f"units={i['last_units']} cands={i['last_candidates']} appends={i['last_appends']} slow_chunks={i['last_slow_chunks']} wasted={i['last_wasted_chunks']} "
              f"compactions={i['last_compactions']} mma_wait_epi={i['mma_wait_epilogue_cycles']/max(1,i['mma_total_cycles']):.3f} "
              f"mma_wait_tma={i['mma_wait_tma_cycles']/max(1,i['mma_total_cycles']):.3f} "
              f"epi[wait,slow,compact,end,-,hold,ld]={[round(x / max(1, i['epilogue_cycles'][4]), 3) for x in i['epilogue_cycles']]} rounds_kcyc={[round(x / max(1, i['sm_count'] // 2) / 1e3) for x in i['round_cycles']]} round_wait_kcyc={[round(x / max(1, i['sm_count'] // 2) / 1e3) for x in i['round_wait_cycles']]} tile_cycles={i['epilogue_cycles'][4] / max(1, i['sm_count'] * 16) / max(1, (n // 256) * ((b + 255) // 256) / (i['sm_count'] // 2)):.0f}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
        sys.exit(0)
    sizes = ["10000000"]
    for n in sizes:
        for cg in ("2",):
            for flags in ("0", "1"):
                for cons in (0, 1, 2, 3):
                    if flags != "0" and cons:
                        continue
                    env = dict(os.environ, BB_DEBUG_FLAGS=flags, BB_CTA_GROUP=cg)
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), n, "8192", "100", str(cons)], env=env,
                                       capture_output=True, text=True, timeout=900)
                    print(r.stdout.strip() or r.stderr[-2000:], flush=True)
