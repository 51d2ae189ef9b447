#!/usr/bin/env python
"""Which part of a wrong row is wrong? Compares a launch with the oracle and explains the difference by input features."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np, torch
from idash2019_2_b200 import api
from oracle import pyoracle as po
from helpers import make_case

S, G, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = api.Context(0)
geo, model, cts, var = make_case(S, T=500, G=G, n=5, seed=S + 9)
NR, RS = geo.NR, geo.RS
m = api.Model(ctx, S, NR, RS, model.out_bidx, model.row_ptr, model.col, model.coef)
L = api.compile_layout(S, NR, RS, model.out_bidx, model.row_ptr, model.col, model.coef)
print("tiles f_base/32, K/32:", [(int(t['f_base']) // 32, int(t['K']) // 32) for t in L.tiles])
x = torch.from_numpy(cts.view(np.int32)).cuda()
ref, _ = po.cloud_port(S, NR, RS, np.arange(len(cts), dtype=np.uint32), cts, var, model.row_ptr, model.col, model.coef)
def rot(ct, r):      # X^(-r RS) * ct, both polynomials
    s = r * RS
    o = np.empty(2048, np.uint32)
    for p0 in (0, 1024):
        v = cts[ct, p0:p0 + 1024]
        o[p0:p0 + 1024 - s] = v[s:]
        o[p0 + 1024 - s:p0 + 1024] = (0 - v[:s]).astype(np.uint32)
    return o
for rep in range(reps):
    out = torch.full((model.n_out, 2048), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
    api.cloud_compute_score_device(ctx, m, x, out)
    torch.cuda.synchronize()
    o = out.cpu().numpy().view(np.uint32)
    bad = np.argwhere((o != ref).any(axis=1)).flatten()
    print("rep", rep, "bad rows", bad.tolist())
    for r in bad[:2]:
        sl = [int((o[r, 128 * s:128 * s + 128] != ref[r, 128 * s:128 * s + 128]).any()) for s in range(16)]
        print("  row", r, "bad slices", sl)
        s0 = sl.index(1)
        w = 128 * s0
        diff = (o[r, w:w + 4].astype(np.int64) - ref[r, w:w + 4].astype(np.int64)) % (1 << 32)
        ent = [(int(model.col[e]), int(model.coef[e])) for e in range(model.row_ptr[r], model.row_ptr[r + 1]) if model.col[e] != 0xFFFFFFFF]
        print("  entries (feature, block, coef):", [(f, f // 32, c) for f, c in ent])
        # does dropping a set of features explain the difference?
        for blk in sorted({f // 32 for f, _ in ent}):
            contrib = np.zeros(4, np.int64)
            for f, c in ent:
                if f // 32 == blk:
                    contrib += c * rot(f // NR, f % NR)[w:w + 4].astype(np.int64)
            print("   block", blk, "missing would give diff", ((-contrib) % (1 << 32)).tolist(), "observed", diff.tolist())
