"""BASELINE config 2 (4096 queries x 25,216 sets, content + CF + mask, top-50) as a small script for ncu launch lists."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth  # noqa: E402
import torch  # noqa: E402
import brickbrain_rec_engine_b200 as bb  # noqa: E402

n, d, f, b = 25216, 384, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
themes = synth.themes_table()
parts, year, theme = synth.metadata(n, 3, themes)
eng = bb.ScoringEngine(n, d, f)
eng.load_catalog(emb=synth.embeddings(n, d, 1, "aniso"), item_factors=synth.cf_factors(n, f, 2), num_parts=parts, year=year, theme_id=theme)
filt = bb.HardConstraintFilter(eng, synth.set_nums(n), themes)
desc = filt.descriptor(filt.create_constraint_set(pieces_max=1000, year_min=2010))
cons = torch.from_numpy(np.atleast_1d(desc).view(np.uint8).reshape(1, -1).copy()).cuda()
q = torch.from_numpy(synth.embeddings(b, d, 6, "aniso")).cuda()
u = torch.from_numpy(synth.cf_factors(b, f, 7)).cuda()
for prec in ("bf16_rescore", "fp32"):
    for _ in range(3):
        eng.search(50, q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, precision=prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.search(50, q_emb=q, q_cf=u, w_content=0.4, w_cf=0.6, constraints=cons, precision=prec)
    e1.record()
    torch.cuda.synchronize()
    i = eng.info()
    print(f"config2 b={b} {prec}: {e0.elapsed_time(e1) / 10:.3f} ms/search units={i['last_units']} appends={i['last_appends']} compactions={i['last_compactions']}")
