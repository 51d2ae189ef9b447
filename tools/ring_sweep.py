#!/usr/bin/env python
"""Ring-kernel schedule sweep on one GPU (profiling build of the library: IDASH_B200_USE_PROFILE_LIB=1 is set here).

For every workload (samples, neighbors) the model is compiled and uploaded ONCE; every setting (IDASH_B200_TUNE / _RING_SLOTS /
_RING_BCHUNKS / _COEF_PREFETCH, read by the profiling build at each launch) is timed with per-launch CUDA events after warm-up,
and its output is compared word for word with the first setting's output (a schedule switch must not change a single bit).

usage: ring_sweep.py [--workloads 1004:5,1004:20,1004:50,335:20] [--settings "8;40;72;104;104,slots=9"] [--steps 20] [--out FILE]
  a setting is  TUNE[,slots=N][,bchunks=N][,prefetch=N][,ko=MASK (knock-out: results are wrong, same_bits is False)][,slices=N]
"""
import argparse
import json
import os
import statistics
import sys
from pathlib import Path

os.environ["IDASH_B200_USE_PROFILE_LIB"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from idash2019_2_b200 import api, synth  # noqa: E402

T, G, SEED = 16184, 80882, 1234
ENV = {"slots": "IDASH_B200_RING_SLOTS", "bchunks": "IDASH_B200_RING_BCHUNKS", "prefetch": "IDASH_B200_COEF_PREFETCH",
       "ko": "IDASH_B200_KNOCKOUT", "slices": "IDASH_B200_RING_SLICES", "chunks": "IDASH_B200_RING_CHUNKS",
       "extra": "IDASH_B200_RING_EXTRA", "fill": "IDASH_B200_RING_EXTRA_FILL"}


def apply(setting: str):
    parts = setting.split(",")
    for v in ENV.values():
        os.environ.pop(v, None)
    os.environ["IDASH_B200_TUNE"] = parts[0]
    for kv in parts[1:]:
        k, v = kv.split("=")
        os.environ[ENV[k]] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="1004:5,1004:20,1004:50,335:20")
    ap.add_argument("--settings", default="8;40;72;104")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--targets", type=int, default=G)
    ap.add_argument("--out", default=None)
    ap.add_argument("--trace-dir", default=None, help="also record the per-tile timeline of CTA 0 under every setting (tools/trace_ring.py)")
    a = ap.parse_args()
    peak = 6553.3
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    ctx = api.Context(0)
    rows = []
    for wl in a.workloads.split(","):
        S, n = (int(x) for x in wl.split(":"))
        NR = 1024 // S
        RS = 1024 // NR
        tag, tgt = synth.make_positions(T, a.targets, SEED)
        model = synth.make_model(tag, tgt, n, SEED)
        m = api.Model(ctx, S, NR, RS, model.out_bidx, model.row_ptr, model.col, model.coef)
        n_in = (3 * T - 1) // NR + 1
        gen = torch.Generator(device="cuda").manual_seed(SEED)
        x = torch.randint(-2 ** 31, 2 ** 31, (n_in, 2048), dtype=torch.int32, device="cuda", generator=gen)
        ref = None
        bytes_ = 8192 * (n_in + model.n_out)
        for setting in a.settings.split(";"):
            apply(setting)
            out = torch.zeros((model.n_out, 2048), dtype=torch.int32, device="cuda")
            try:
                for _ in range(3):
                    api.cloud_compute_score_device(ctx, m, x, out)
                torch.cuda.synchronize()
                ctx.check_device_status()
                ctx.timing_enable(a.steps)
                for _ in range(a.steps):
                    api.cloud_compute_score_device(ctx, m, x, out)
                torch.cuda.synchronize()
                ms = ctx.timing_read(a.steps)
                ctx.timing_enable(0)
            except Exception as e:  # a setting the plan rejects
                rows.append({"S": S, "n": n, "setting": setting, "error": repr(e)})
                print(rows[-1], flush=True)
                continue
            if ref is None:
                ref = out.clone()
                same = True
            else:
                same = bool(torch.equal(out, ref))
            if a.trace_dir:
                Path(a.trace_dir).mkdir(parents=True, exist_ok=True)
                os.environ["IDASH_B200_TRACE"] = "0"
                os.environ["IDASH_B200_TRACE_FILE"] = str(Path(a.trace_dir) / f"trace_S{S}_n{n}_{setting.replace(',', '_').replace('=', '')}.txt")
                api.cloud_compute_score_device(ctx, m, x, out)
                torch.cuda.synchronize()
                os.environ.pop("IDASH_B200_TRACE")
            k = statistics.mean(ms)
            rows.append({"S": S, "n": n, "setting": setting, "kernel_ms": round(k, 4), "min_ms": round(min(ms), 4),
                         "frac": round(bytes_ / (k * 1e-3) * 1e-9 / peak, 4), "same_bits": same,
                         "kernel": ctx.last_kernel(), "tile_kmax": m.info["tile_kmax"]})
            print(rows[-1], flush=True)
            del out
        m.free()
        del x, ref
        torch.cuda.empty_cache()
    if a.out:
        Path(a.out).write_text("\n".join(json.dumps(r) for r in rows) + "\n")


if __name__ == "__main__":
    main()
