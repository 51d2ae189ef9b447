#!/usr/bin/env python
"""Times the batched phase + decode kernel (K4, decrypt_predictions) on the output of one iDASH-scale cloud evaluation:
242 646 ciphertexts, S = 1004. Prints one JSON line (device-resident kernel time, achieved bytes/s, and the reference's
decrypt_predictions on a bounded sample of the same ciphertexts on the host cores)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    from idash2019_2_b200 import api
    from oracle import pyoracle as po
    n, S = int(os.environ.get("N_CT", 242646)), 1004
    ctx = api.Context(0)
    g = torch.Generator(device="cuda").manual_seed(7)
    ct = torch.randint(-2 ** 31, 2 ** 31, (n, 2048), dtype=torch.int32, device="cuda", generator=g)
    key = np.random.default_rng(1).integers(0, 2, 1024).astype(np.int32)
    scores = torch.empty((n, S), dtype=torch.float32, device="cuda")
    for _ in range(3):
        api.decrypt_predictions_device(ctx, key, S, ct, scores)
    torch.cuda.synchronize()
    ctx.timing_enable(10)
    for _ in range(10):
        api.decrypt_predictions_device(ctx, key, S, ct, scores)
    torch.cuda.synchronize()
    ms = ctx.timing_read(10)
    k_ms = float(np.mean(ms))
    # parity of a sample against the exact oracle
    sample = 64
    ref_phase = po.phase_exact_port(key, ct[:sample].cpu().numpy().view(np.uint32))
    ok = bool(np.array_equal(scores[:sample].cpu().numpy(), po.decode_port(S, ref_phase)))
    line = {"kernel": "decrypt_kernel", "ciphertexts": n, "S": S, "kernel_ms": k_ms, "ct_per_s": n / (k_ms * 1e-3),
            "algorithmic_bytes": n * (8192 + 4 * S), "achieved_GBps": n * (8192 + 4 * S) / (k_ms * 1e-3) * 1e-9,
            "sample_matches_exact_oracle": ok}
    if po.have_ref():
        m = min(n, 20001) // 3 * 3
        host = ct[:m].cpu().numpy().view(np.uint32)
        os.environ["OMP_NUM_THREADS"] = str(po.host_threads())
        _, dt = po.decrypt_ref(S, key, host)          # the reference's own "decrypt wall time" 
        line["cpu_reference"] = {"ct_per_s": m / dt, "sample": m, "cores": po.host_threads(), "what": "reference decrypt_predictions (FFT path), its own wall time"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
