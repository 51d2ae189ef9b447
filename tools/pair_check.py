#!/usr/bin/env python
"""Quick parity check of the CTA-pair decrypt kernel (K4p) against the exact oracle; run under `timeout`."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from idash2019_2_b200 import api
from oracle import pyoracle as po

ctx = api.Context(0)
ctx.set_decrypt_kernel(api.DECRYPT_TENSOR_PAIR)
for n_ct, S in ((32, 1004), (37, 1004), (64, 16), (333, 400), (5000, 1004), (20000, 1024)):
    rng = np.random.default_rng(n_ct)
    key = rng.integers(0, 2, 1024).astype(np.int32)
    ct = rng.integers(0, 2 ** 32, size=(n_ct, 2048), dtype=np.uint32)
    ct[0] = 0xFFFFFFFF
    scores, phase = api.decrypt_predictions(ctx, key, S, ct, want_phase=True)
    ref = po.phase_exact_port(key, ct)
    ok_p = np.array_equal(phase, ref)
    ok_s = np.array_equal(scores, po.decode_port(S, ref))
    bad = np.argwhere(phase != ref)
    print(n_ct, S, "phase", ok_p, "scores", ok_s, "kernel", ctx.last_decrypt_kernel(), "first bad", bad[:3].tolist(), "n bad", len(bad), flush=True)
