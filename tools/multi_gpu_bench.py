#!/usr/bin/env python
"""One process, N GPUs: idash_b200_cloud_eval_multi_device at iDASH scale -- the job's ciphertexts resident on GPU 0, the target range
cut over the first n GPUs (peer-copied input slabs, rows stored straight into GPU 0's output array over NVLink).
Prints one JSON line per n with the wall time of the synchronous call (median of --reps), checked against the unsharded result.

usage: multi_gpu_bench.py [--neighbors 20] [--samples 1004] [--gpus 1,2,4,8] [--reps 7]"""
import argparse
import json
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from idash2019_2_b200 import api, synth  # noqa: E402

T, G, SEED = 16184, 80882, 1234


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--neighbors", type=int, default=20)
    ap.add_argument("--samples", type=int, default=1004)
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--reps", type=int, default=7)
    a = ap.parse_args()
    have = torch.cuda.device_count()
    S = a.samples
    geo = synth.Geometry(S, T, G)
    tag, tgt = synth.make_positions(T, G, SEED)
    model = synth.make_model(tag, tgt, a.neighbors, SEED)
    n_in, n_rows = geo.n_in_ct_used, 3 * G
    ctxs = [api.Context(g) for g in range(have)]
    m0 = api.Model(ctxs[0], S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef)
    models = [m0] + [m0.clone(ctxs[g]) for g in range(1, have)]
    with torch.cuda.device(0):
        gen = torch.Generator(device="cuda").manual_seed(SEED)
        x = torch.randint(-2 ** 31, 2 ** 31, (n_in, 2048), dtype=torch.int32, device="cuda", generator=gen)
        whole = torch.empty((n_rows, 2048), dtype=torch.int32, device="cuda")
        api.cloud_compute_score_device(ctxs[0], m0, x, whole)
        torch.cuda.synchronize()
        out = torch.empty_like(whole)
        for n in [int(v) for v in a.gpus.split(",")]:
            if n > have:
                continue
            out.zero_()
            torch.cuda.synchronize()
            ts = []
            for _ in range(a.reps + 2):
                t0 = time.perf_counter()
                api.cloud_compute_score_multi_device(ctxs[:n], models[:n], x, out)
                ts.append((time.perf_counter() - t0) * 1e3)
            ms = statistics.median(ts[2:])
            print(json.dumps({"n_gpus": n, "neighbors": a.neighbors, "samples": S, "ms_per_evaluation": ms, "min_ms": min(ts[2:]),
                              "equals_unsharded": bool(torch.equal(out, whole)),
                              "out_ct_per_s": n_rows / (ms * 1e-3), "bytes_over_nvlink": int((n - 1) / n * (n_rows + n_in) * 8192) if n > 1 else 0,
                              "how": "wall time of the synchronous idash_b200_cloud_eval_multi_device call: peer copies of the input slabs, "
                                     "kernels on every GPU storing into GPU 0's output array, all streams synchronised"}), flush=True)
    for m in models:
        m.free()
    for c in ctxs:
        c.close()


if __name__ == "__main__":
    main()
