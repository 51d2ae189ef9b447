#!/usr/bin/env python
"""The whole file-level pipeline at iDASH scale with the real binaries, reference vs B200 host layer:

    synthetic tags / targets / .hr model dir  (idash2019_2_b200.synth, seed 1234)
    reference keygen, encrypt                 -> params.bin keys.bin encrypted_data.bin   (oracle/_ref/bin, unmodified)
    reference cloud                           -> ref/encrypted_prediction.bin + its BENCHMARK block
    B200 cloud (idash2019_2_b200/lib/bin)     -> b200/encrypted_prediction.bin + BENCHMARK block
    cmp of the two 2 GB files (byte-identical is the bar), B200 decrypt -> result_bypos.csv timing

Prints one JSON line. Needs a GPU for the B200 stages; --no-gpu stops after the reference stages (sizes the inputs).
usage: pipeline_at_scale.py [--samples 1004 --tags 16184 --targets 80882 --neighbors 5 --workdir DIR --decrypt]"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def bench_block(text):
    out = {}
    for key, pat in (("stage_s", r"(?:fhe|decrypt) wall time \(seconds\)\.*: ([0-9.e+-]+)"), ("serialization_s", r"serialization wall time \(seconds\): ([0-9.e+-]+)"),
                     ("total_s", r"total wall time \(seconds\)\.*: ([0-9.e+-]+)"), ("rss_mb", r"RAM usage \(MB\)\.*: ([0-9.e+-]+)"),
                     ("gpu_call_s", r"gpu call wall time \(seconds\)\.*: ([0-9.e+-]+)")):
        m = re.search(pat, text)
        if m:
            out[key] = float(m.group(1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=1004)
    ap.add_argument("--tags", type=int, default=16184)
    ap.add_argument("--targets", type=int, default=80882)
    ap.add_argument("--neighbors", type=int, default=5)
    ap.add_argument("--workdir", default=None)
    ap.add_argument("--decrypt", action="store_true")
    ap.add_argument("--ref-decrypt", action="store_true")
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--gpus", default="", help='also run the B200 cloud with IDASH_GPUS set to this list, e.g. "0,1"')
    a = ap.parse_args()
    from idash2019_2_b200 import synth
    from oracle import pyoracle as po
    assert po.have_ref(), "oracle/_ref is not built"
    work = Path(a.workdir or tempfile.mkdtemp(prefix="idash_scale_"))
    work.mkdir(parents=True, exist_ok=True)
    res = {"geometry": {"S": a.samples, "T": a.tags, "G": a.targets, "n": a.neighbors}, "host_threads": po.host_threads()}
    t0 = time.perf_counter()
    tag, tgt = synth.make_positions(a.tags, a.targets, 1234)
    geno = synth.make_genotypes(a.tags, a.samples, 1234, na_frac=0.01)
    model = synth.make_model(tag, tgt, a.neighbors, 1234)
    synth.write_tag_file(work / "tags.txt", tag, geno)
    synth.write_target_file(work / "targets.txt", tgt)
    synth.write_hr_dir(work / "model", model, tag, tgt)
    res["generate_inputs_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    po.run_ref_bin("keygen", [work / "targets.txt", work / "tags.txt", 1], work)
    res["ref_encrypt"] = bench_block(po.run_ref_bin("encrypt", [work / "tags.txt"], work))
    res["ref_keygen_encrypt_wall_s"] = time.perf_counter() - t0
    for d in ("ref", "b200"):
        (work / d).mkdir(exist_ok=True)
        for f in ("params.bin", "keys.bin", "encrypted_data.bin"):
            dst = work / d / f
            if dst.exists() or dst.is_symlink():
                dst.unlink()
            os.symlink(work / f, dst)
    t0 = time.perf_counter()
    res["ref_cloud"] = bench_block(po.run_ref_bin("cloud", [work / "model"], work / "ref"))
    res["ref_cloud"]["wall_s"] = time.perf_counter() - t0
    res["bytes"] = {"encrypted_data": (work / "encrypted_data.bin").stat().st_size,
                    "encrypted_prediction": (work / "ref" / "encrypted_prediction.bin").stat().st_size}
    if not a.no_gpu:
        bins = ROOT / "idash2019_2_b200" / "lib" / "bin"
        cache = work / "b200" / "models.bin"
        if cache.exists():
            cache.unlink()
        # run 1 parses the .hr files, compiles and writes models.bin; run 2 takes the compiled model from it; run 3 (with --gpus)
        # shards the target range over several GPUs
        runs = [("b200_cloud", {}), ("b200_cloud_cached_model", {})]
        if a.gpus:
            runs.append(("b200_cloud_gpus_" + a.gpus.replace(",", "_"), {"IDASH_GPUS": a.gpus}))
        for name, extra in runs:
            t0 = time.perf_counter()
            out = subprocess.run([str(bins / "cloud"), str(work / "model")], cwd=work / "b200", capture_output=True, text=True,
                                 env=dict(os.environ, IDASH_HOST_TIMING="1", **extra))
            assert out.returncode == 0, out.stdout + out.stderr
            res[name] = bench_block(out.stdout)
            res[name]["phases"] = [ln for ln in out.stderr.splitlines() if ln.startswith("[idash_host]")]
            res[name]["wall_s"] = time.perf_counter() - t0
            res[name]["byte_identical"] = subprocess.run(["cmp", str(work / "ref" / "encrypted_prediction.bin"),
                                                          str(work / "b200" / "encrypted_prediction.bin")]).returncode == 0
        t0 = time.perf_counter()
        same = subprocess.run(["cmp", str(work / "ref" / "encrypted_prediction.bin"), str(work / "b200" / "encrypted_prediction.bin")]).returncode == 0
        res["encrypted_prediction_byte_identical"] = same
        res["cmp_s"] = time.perf_counter() - t0
        if a.decrypt:
            if a.ref_decrypt:      # ~5 minutes of host time at iDASH scale (one flush per csv row)
                t0 = time.perf_counter()
                rout = po.run_ref_bin("decrypt", ["bypos"], work / "ref")
                res["ref_decrypt"] = bench_block(rout)
                res["ref_decrypt"]["wall_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            out = subprocess.run([str(bins / "decrypt"), "bypos"], cwd=work / "b200", capture_output=True, text=True,
                                 env=dict(os.environ, IDASH_HOST_TIMING="1"))
            assert out.returncode == 0, out.stdout + out.stderr
            res["b200_decrypt"] = bench_block(out.stdout)
            res["b200_decrypt"]["phases"] = [ln for ln in out.stderr.splitlines() if ln.startswith("[idash_host]")]
            res["b200_decrypt"]["wall_s"] = time.perf_counter() - t0
            res["bytes"]["result_bypos_csv"] = (work / "b200" / "result_bypos.csv").stat().st_size
    print(json.dumps(res))
    if not a.workdir:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
