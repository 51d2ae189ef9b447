#!/usr/bin/env python
"""BASELINE configs[3] at full size: three population-stratified model sets (335 / 334 / 335 samples of the same 16184 tag x 80882 target
SNPs, NUM_REGIONS = 3, neighbors = 20, each population its own coefficients) -- ONE launch of the ring kernel for all three
(idash_b200_cloud_eval_device_multi_model) against three launches back to back on one stream. CUDA events, median of --reps.
Prints one JSON line; the algorithmic bytes are 3 x 8192 x (16184 + 242646)."""
import argparse
import json
import statistics
import sys
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
    ap.add_argument("--sizes", default="335,334,335")
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    sizes = [int(v) for v in a.sizes.split(",")]
    ctx = api.Context(0)
    tag, tgt = synth.make_positions(T, G, SEED)
    models, ins, outs, outs1 = [], [], [], []
    gen = torch.Generator(device="cuda").manual_seed(SEED)
    for k, S in enumerate(sizes):
        geo = synth.Geometry(S, T, G)
        mod = synth.make_model(tag, tgt, a.neighbors, SEED + k)          # same positions, the population's own coefficients
        models.append(api.Model(ctx, S, geo.NR, geo.RS, mod.out_bidx, mod.row_ptr, mod.col, mod.coef))
        ins.append(torch.randint(-2 ** 31, 2 ** 31, (geo.n_in_ct_used, 2048), dtype=torch.int32, device="cuda", generator=gen))
        outs.append(torch.empty((3 * G, 2048), dtype=torch.int32, device="cuda"))
        outs1.append(torch.empty((3 * G, 2048), dtype=torch.int32, device="cuda"))
    n_in = ins[0].shape[0]

    def one():
        api.cloud_compute_score_device_multi_model(ctx, models, ins, outs)

    def three():
        for b in range(len(sizes)):
            api.cloud_compute_score_device(ctx, models[b], ins[b], outs1[b])

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    t3, t1 = timed(three), timed(one)
    same = all(bool(torch.equal(x, y)) for x, y in zip(outs, outs1))
    alg = len(sizes) * 8192 * (n_in + 3 * G)
    peak = 6553.3
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    print(json.dumps({"workload": f"{len(sizes)} populations ({a.sizes} samples) x {T} tag x {G} target SNPs, neighbors={a.neighbors}, NUM_REGIONS={1024 // sizes[0]}",
                      "one_launch_ms": t1, "three_launches_ms": t3, "one_launch_equals_separate": same, "algorithmic_bytes": alg,
                      "one_launch_frac": alg / (t1 * 1e-3) * 1e-9 / peak, "three_launches_frac": alg / (t3 * 1e-3) * 1e-9 / peak, "peak_gbs": peak}))
    for m in models:
        m.free()
    ctx.close()


if __name__ == "__main__":
    main()
