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
    line = {"ciphertexts": n, "S": S, "algorithmic_bytes": n * (8192 + 4 * S), "kernels": {}}
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    quick = bool(os.environ.get("DEC_QUICK"))       # tensor kernel only, no CPU leg (knock-out / tuning runs)
    kernels = (("decrypt_pair_kernel", api.DECRYPT_TENSOR_PAIR), ("decrypt_tc_kernel", api.DECRYPT_TENSOR), ("decrypt_kernel", api.DECRYPT_IADD))
    if os.environ.get("DEC_KERNELS"):      # e.g. DEC_KERNELS=decrypt_tc_kernel
        kernels = tuple(k for k in kernels if k[0] in os.environ["DEC_KERNELS"].split(","))
    for name, which in kernels[:1] if quick else kernels:
        ctx.set_decrypt_kernel(which)
        scores.zero_()
        for _ in range(3):
            api.decrypt_predictions_device(ctx, key, S, ct, scores)
        torch.cuda.synchronize()
        ctx.timing_enable(10)
        for _ in range(10):
            api.decrypt_predictions_device(ctx, key, S, ct, scores)
        torch.cuda.synchronize()
        ms = ctx.timing_read(10)
        k_ms = float(np.mean(ms))
        # parity of a sample (first, middle, last ciphertexts) against the exact oracle
        idx = np.unique(np.concatenate([np.arange(min(n, 64)), np.arange(n // 2, min(n, n // 2 + 64)), np.arange(max(0, n - 64), n)]))
        ref_phase = po.phase_exact_port(key, ct[idx].cpu().numpy().view(np.uint32))
        ok = bool(np.array_equal(scores[idx].cpu().numpy(), po.decode_port(S, ref_phase)))
        gbs = n * (8192 + 4 * S) / (k_ms * 1e-3) * 1e-9
        line["kernels"][name] = {"kernel_ms": k_ms, "min_ms": float(np.min(ms)), "ct_per_s": n / (k_ms * 1e-3), "achieved_GBps": gbs,
                                 "frac_of_measured_hbm_peak": gbs / peaks["hbm_gbs"] if "hbm_gbs" in peaks else None,
                                 "int8_mac_per_s": n * 4 * 1024 * 1024 / (k_ms * 1e-3) if which != api.DECRYPT_IADD else None,
                                 "sample_matches_exact_oracle": ok}
    ctx.set_decrypt_kernel(api.DECRYPT_AUTO)
    if po.have_ref() and not quick:
        m = min(n, 20001) // 3 * 3
        host = ct[:m].cpu().numpy().view(np.uint32)
        os.environ["OMP_NUM_THREADS"] = str(po.host_threads())
        _, dt = po.decrypt_ref(S, key, host)          # the reference's own "decrypt wall time" 
        line["cpu_reference"] = {"ct_per_s": m / dt, "sample": m, "cores": po.host_threads(), "what": "reference decrypt_predictions (FFT path), its own wall time"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
