import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np, synth
import brickbrain_rec_engine_b200 as bb
n, d = 25216, 384
emb = synth.embeddings(n, d, 1, "aniso")
eng = bb.ScoringEngine(n, d, 0, device=0)
eng.load_catalog(emb=emb)
mask = np.random.default_rng(0).random(n) < 0.3
bm = eng.pack_bitmap(mask).reshape(1, -1)
cons = bb.no_constraint(); cons["allow_bitmap"] = 0
rows = np.array([77], dtype=np.int32)
q = emb[77:78].copy()
def t(label, **kw):
    for _ in range(5): eng.search(10, precision="bf16_rescore", **kw)
    t0 = time.perf_counter()
    for _ in range(200): eng.search(10, precision="bf16_rescore", **kw)
    print(label, f"{(time.perf_counter()-t0)/200*1e3:.3f} ms", eng.info()["launches"], flush=True)
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
if mode in ("all", "a"): t("q_emb", q_emb=q)
if mode in ("all", "b"): t("q_item+self", q_item=rows, self_id=rows)
if mode in ("all", "c"): t("q_emb+cons+allow", q_emb=q, constraints=cons, allow_bitmaps=bm)
if mode in ("all", "d"): t("q_item+self+cons+allow", q_item=rows, self_id=rows, constraints=cons, allow_bitmaps=bm)
