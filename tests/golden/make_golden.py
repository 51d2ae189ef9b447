"""Generates the golden fixtures in this directory by running the UNMODIFIED reference
(oracle/_ref, built from /root/reference by oracle/build_ref.sh) on small seeded synthetic inputs.

    python tests/golden/make_golden.py          (only where /root/reference exists)

Per case <name>/ :
    tags.txt targets.txt model/*.hr             synthetic inputs (idash2019_2_b200.synth, seed in CASES)
    params.bin keys.bin encrypted_data.bin      written by the reference keygen / encrypt
    encrypted_prediction.bin                    written by the reference cloud   (the parity target)
    ref.npz                                     phase_fft / phase_exact (reference tLwePhase and TFHE's exact
                                                Karatsuba product through oracle/ref_shim.cpp), scores
                                                (reference decrypt_predictions), model CSR as the reference's
                                                read_model produced it, constants
The reference never seeds its RNG, so these files are reproducible bit for bit.
"""
from __future__ import annotations

import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from idash2019_2_b200 import formats, synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

# name: (S, T, G, neighbors, seed, na_frac, coef_range, bias_range)
CASES = {
    "s1004_nr1": (1004, 9, 6, 5, 11, 0.02, 200, 500),
    "s400_nr2": (400, 12, 6, 5, 12, 0.0, 8191, 8191),
    "s335_nr3": (335, 12, 6, 4, 13, 0.0, 200, 500),
    "s16_nr64": (16, 70, 8, 5, 14, 0.05, 200, 500),
}


def make_case(name, S, T, G, n, seed, na_frac, coef_range, bias_range):
    out = HERE / name
    if out.exists():
        shutil.rmtree(out)
    out.mkdir()
    tag_pos, tgt_pos = synth.make_positions(T, G, seed)
    geno = synth.make_genotypes(T, S, seed, na_frac=na_frac)
    model = synth.make_model(tag_pos, tgt_pos, n, seed, coef_range=coef_range, bias_range=bias_range)
    synth.write_tag_file(out / "tags.txt", tag_pos, geno)
    synth.write_target_file(out / "targets.txt", tgt_pos)
    synth.write_hr_dir(out / "model", model, tag_pos, tgt_pos)
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        po.run_ref_bin("keygen", [out / "targets.txt", out / "tags.txt", 1], tmp, threads=2)
        po.run_ref_bin("encrypt", [out / "tags.txt"], tmp, threads=2)
        log = po.run_ref_bin("cloud", [out / "model"], tmp, threads=2)
        assert "fhe wall time" in log
        for f in ("params.bin", "keys.bin", "encrypted_data.bin", "encrypted_prediction.bin"):
            shutil.copy(tmp / f, out / f)
    params, key, _ = formats.read_key(out / "keys.bin")
    assert params.NUM_SAMPLES == S
    pred = formats.read_ct_image(out / "encrypted_prediction.bin")
    idx, words, _ = formats.image_views(pred)
    order = np.argsort(idx)
    ct_sorted = np.ascontiguousarray(words[order])
    phase_fft = po.phase_ref(key, ct_sorted, True)
    phase_exact = po.phase_ref(key, ct_sorted, False)
    scores, _ = po.decrypt_ref(S, key, ct_sorted)
    g7, ob, rp, col, coef = po.read_model_ref(out / "params.bin", out / "model")
    c4 = np.zeros(4, np.int32)
    po.ref().ref_constants(c4)
    np.savez_compressed(out / "ref.npz", pred_index_sorted=idx[order].copy(), phase_fft=phase_fft, phase_exact=phase_exact,
                        scores=scores, geometry=g7, model_out_bidx=ob, model_row_ptr=rp, model_col=col, model_coef=coef,
                        constants=c4, genotypes=geno, tag_pos=tag_pos, target_pos=tgt_pos)
    size = sum(f.stat().st_size for f in out.rglob("*") if f.is_file())
    print(f"{name}: S={S} NR={params.NUM_REGIONS} RS={params.REGION_SIZE} in_ct={formats.check_ct_image(formats.read_ct_image(out / 'encrypted_data.bin'))} "
          f"out_ct={len(idx)} bytes={size}")


if __name__ == "__main__":
    assert po.build_ref(), "reference not built (needs /root/reference)"
    for name, args in CASES.items():
        make_case(name, *args)
