"""First-contact GPU probe: runs each scan pipeline on a tiny problem in its own
subprocess (bounded by a timeout), compares the dense score dump with a NumPy
restatement and saves the arrays under gpurun_out/ for offline diagnosis.

    python tools/gpu_probe.py            # all stages
    python tools/gpu_probe.py tc 2       # one stage (child mode)
"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
OUT = os.path.join(ROOT, "gpurun_out")


def bf16_round(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(torch.bfloat16).to(torch.float32).numpy()


def child(mode, cg, n=1000, b=70, d=384, f=64):
    os.environ["BB_CTA_GROUP"] = str(cg)
    import brick_oracle as bo
    import synth
    from brickbrain_rec_engine_b200 import ScoringEngine
    emb = synth.embeddings(n, d, 1, "aniso")
    itf = synth.cf_factors(n, f, 2)
    q = synth.embeddings(b, d, 3, "aniso") * 3.0
    u = synth.cf_factors(b, f, 4)
    eng = ScoringEngine(n, d, f)
    eng.load_catalog(emb=emb, item_factors=itf)
    wc, wf = 0.4, 0.6
    t0 = time.time()
    got = eng.debug_scores(q_emb=q, q_cf=u, w_content=wc, w_cf=wf, precision=("fp32" if mode == "fp32" else "bf16"))
    dt = time.time() - t0
    e_unit = bo.l2_normalize_rows(emb)
    if mode == "fp32":
        want = bo.fused_hybrid_scores(q, u, e_unit, itf, np.full(b, wc, np.float32), np.full(b, wf, np.float32))
    else:
        qq = np.concatenate([np.float32(wc) * bo.l2_normalize_rows(q), np.float32(wf) * u], axis=1)
        ee = np.concatenate([e_unit, itf], axis=1)
        want = bf16_round(qq) @ bf16_round(ee).T
    err = np.abs(got - want)
    tag = f"{mode}_cg{cg}"
    np.save(os.path.join(OUT, f"probe_{tag}_got.npy"), got)
    np.save(os.path.join(OUT, f"probe_{tag}_want.npy"), want)
    print(f"[probe {tag}] n={n} b={b} max|err|={err.max():.3e} mean|err|={err.mean():.3e} "
          f"max|want|={np.abs(want).max():.3f} time={dt:.3f}s info={eng.info()}", flush=True)
    # search too
    for prec in (["fp32"] if mode == "fp32" else ["bf16", "bf16_rescore"]):
        s, i = eng.search(10, q_emb=q, q_cf=u, w_content=wc, w_cf=wf, precision=prec)
        ref = bo.fused_hybrid_scores(q, u, e_unit, itf, np.full(b, wc, np.float32), np.full(b, wf, np.float32))
        rid, rs = bo.fused_hybrid_topk(ref, 10)
        same = (rid == i).mean()
        print(f"[probe {tag}] search {prec}: id agreement {same:.4f} max score err {np.abs(rs - s).max():.3e}", flush=True)
    return 0 if err.max() < (1e-4 if mode == "fp32" else 2e-3) else 3


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) >= 3:
        sys.exit(child(sys.argv[1], int(sys.argv[2])))
    rc_all = 0
    for mode, cg in (("fp32", 2), ("tc", 1), ("tc", 2)):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), mode, str(cg)], timeout=240,
                               capture_output=True, text=True)
            print(r.stdout[-3000:])
            if r.returncode != 0:
                print(f"[probe] {mode} cg={cg} exit {r.returncode}\n{r.stderr[-3000:]}")
                rc_all = 1
        except subprocess.TimeoutExpired:
            print(f"[probe] {mode} cg={cg} TIMEOUT")
            rc_all = 1
    sys.exit(rc_all)


if __name__ == "__main__":
    main()
