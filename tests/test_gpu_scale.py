"""Parity at BASELINE.json's full sizes (1004 / 335 samples x 16184 tag x 80882 target SNPs): EVERY output ciphertext word and
every variance of the CUDA path against the reference's own cloud_compute_score (oracle/_ref, the unmodified reference compiled
by oracle/build_ref.sh; the C restatement where that library is absent) on the same inputs -- eval/idash.cpp:763-848.

The small cases of test_gpu_cloud.py walk at most 17 tiles per chunk; here a persistent CTA walks ~421 tiles (ring-slot / TMEM /
coefficient-ring parities wrap hundreds of times, 20-bit header fields, prefetch bounds), at neighbors = 5, 50 and with
NUM_REGIONS = 3 rotations -- and configs[0] (16 samples x 1k tags x 5k targets) goes through all four stages of the pipeline
with the reference's keygen / encrypt feeding both `cloud` binaries.
"""
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from idash2019_2_b200 import _lib, api, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "idash2019_2_b200" / "lib" / "bin"
T_FULL, G_FULL, SEED = 16184, 80882, 1234
ALPHA2 = 2.0 ** -50


def _reference(S, NR, RS, cts, var, model):
    idx = np.arange(len(cts), dtype=np.uint32)
    if po.have_ref():
        out, ovar, _ = po.cloud_ref(S, NR, RS, idx, cts, var, model.out_bidx, model.row_ptr, model.col, model.coef)
        return out, ovar, "reference"
    out, ovar = po.cloud_port(S, NR, RS, idx, cts, var, model.row_ptr, model.col, model.coef)
    return out, ovar, "port"


def _equal_on_device(out_dev, ref_host, piece=16384):
    """torch.equal of a [n, 2048] int32 CUDA tensor with a host uint32 array, in pieces (the host array is not pinned)."""
    import torch
    bad = 0
    for lo in range(0, len(ref_host), piece):
        r = torch.from_numpy(ref_host[lo:lo + piece].view(np.int32)).to(out_dev.device)
        bad += int((out_dev[lo:lo + piece] != r).any(dim=1).sum())
    return bad


@pytest.mark.parametrize("S,n", [(1004, 5), (1004, 50), (335, 20)])
def test_cloud_full_size_every_word_equals_the_reference(gpu_ctx, S, n):
    import torch
    geo = synth.Geometry(S, T_FULL, G_FULL)
    tag, tgt = synth.make_positions(T_FULL, G_FULL, SEED)
    model = synth.make_model(tag, tgt, n, SEED)
    n_in = geo.n_in_ct_used
    cts = synth.random_ciphertexts(n_in, SEED + n)
    var = np.full(n_in, ALPHA2)
    gpu_ctx.set_kernel(_lib.KERNEL_AUTO)
    m = api.Model(gpu_ctx, S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef)
    assert m.info["ring_ok"] == 1 and m.info["n_rows"] == 3 * G_FULL
    x = torch.from_numpy(cts.view(np.int32)).cuda()
    xv = torch.from_numpy(var).cuda()
    out = torch.zeros((model.n_out, 2048), dtype=torch.int32, device="cuda")
    ovar = torch.zeros(model.n_out, dtype=torch.float64, device="cuda")
    oidx = torch.zeros(model.n_out, dtype=torch.int32, device="cuda")
    api.cloud_compute_score_device(gpu_ctx, m, x, out, in_var=xv, out_index=oidx, out_var=ovar)
    torch.cuda.synchronize()
    gpu_ctx.check_device_status()
    assert gpu_ctx.last_kernel() == _lib.KERNEL_TENSOR_RING
    ref_out, ref_var, kind = _reference(S, geo.NR, geo.RS, cts, var, model)
    bad_rows = _equal_on_device(out, ref_out)
    assert bad_rows == 0, f"{bad_rows} of {model.n_out} output ciphertexts differ from the {kind}"
    assert np.array_equal(ovar.cpu().numpy(), ref_var)
    assert np.array_equal(oidx.cpu().numpy().view(np.uint32), model.out_bidx)
    if S == 1004 and n == 5:
        # the host-buffer entry point (pipelined pieces: copy-in, kernels, copy-out) on the same inputs
        h_out, h_idx, h_var = api.cloud_compute_score(gpu_ctx, m, cts, in_var=var)
        assert np.array_equal(h_out, ref_out) and np.array_equal(h_var, ref_var) and np.array_equal(h_idx, model.out_bidx)
        # eight sample batches of a target range in ONE launch (BASELINE configs[4] per-rank shape): batch 0 = the inputs above
        sub = model.rows(0, 3 * (G_FULL // 8))
        real = sub.col != 0xFFFFFFFF
        n_sub = int(sub.col[real].max()) + 1
        ms = api.Model(gpu_ctx, S, 1, 1024, sub.out_bidx, sub.row_ptr, sub.col, sub.coef)
        g = torch.Generator(device="cuda").manual_seed(7)
        ins = [x[:n_sub].contiguous()] + [torch.randint(-2 ** 31, 2 ** 31, (n_sub, 2048), dtype=torch.int32, device="cuda", generator=g)
                                         for _ in range(7)]
        outs = [torch.zeros((sub.n_out, 2048), dtype=torch.int32, device="cuda") for _ in range(8)]
        api.cloud_compute_score_device_batched(gpu_ctx, ms, ins, outs)
        torch.cuda.synchronize()
        gpu_ctx.check_device_status()
        assert _equal_on_device(outs[0], ref_out[:sub.n_out]) == 0
        single = torch.zeros_like(outs[5])
        api.cloud_compute_score_device(gpu_ctx, ms, ins[5], single)
        torch.cuda.synchronize()
        assert torch.equal(single, outs[5])
        ms.free()
    m.free()


def test_cloud_eight_batches_wide_band_one_launch_full_size(gpu_ctx):
    """Eight sample batches of the full neighbors = 50 model in one call (3 792 x 8 virtual tiles: a 13.5 KB per-CTA header table
    beside 28 KB coefficient images -- the shape whose shared-memory budget used to end in an internal error): the ring plan
    trades input-ring slots for room, the call succeeds, equals single launches, and a sample of rows equals the oracle."""
    import torch
    S, n = 1004, 50
    tag, tgt = synth.make_positions(T_FULL, G_FULL, SEED)
    model = synth.make_model(tag, tgt, n, SEED)
    n_in = synth.Geometry(S, T_FULL, G_FULL).n_in_ct_used
    gpu_ctx.set_kernel(_lib.KERNEL_AUTO)
    m = api.Model(gpu_ctx, S, 1, 1024, model.out_bidx, model.row_ptr, model.col, model.coef)
    g = torch.Generator(device="cuda").manual_seed(3)
    ins = [torch.randint(-2 ** 31, 2 ** 31, (n_in, 2048), dtype=torch.int32, device="cuda", generator=g) for _ in range(8)]
    outs = [torch.zeros((model.n_out, 2048), dtype=torch.int32, device="cuda") for _ in range(8)]
    api.cloud_compute_score_device_batched(gpu_ctx, m, ins, outs)
    torch.cuda.synchronize()
    gpu_ctx.check_device_status()
    assert gpu_ctx.last_kernel() == _lib.KERNEL_TENSOR_RING
    single = torch.zeros_like(outs[0])
    for b in (0, 7):
        api.cloud_compute_score_device(gpu_ctx, m, ins[b], single)
        torch.cuda.synchronize()
        assert torch.equal(single, outs[b]), b
    rows = np.sort(np.random.default_rng(0).choice(model.n_out, 96, replace=False))
    rp, col, coef = [0], [], []
    for r in rows:
        a, b = int(model.row_ptr[r]), int(model.row_ptr[r + 1])
        col.extend(model.col[a:b]); coef.extend(model.coef[a:b]); rp.append(len(col))
    xin = ins[3].cpu().numpy().view(np.uint32)
    ref_out, _ = po.cloud_port(S, 1, 1024, np.arange(n_in, dtype=np.uint32), xin, np.full(n_in, ALPHA2),
                               np.array(rp, np.uint64), np.array(col, np.uint32), np.array(coef, np.int32))
    assert np.array_equal(outs[3][torch.from_numpy(rows).cuda()].cpu().numpy().view(np.uint32), ref_out)
    m.free()
    del ins, outs, single
    torch.cuda.empty_cache()


def _fmt_rows(path):
    rows = Path(path).read_text().splitlines()
    assert rows[0] == "Subject ID,target SNP,0,1,2"
    return rows[1:]


def test_config0_full_pipeline_keygen_encrypt_cloud_decrypt(built_lib, tmp_path):
    """BASELINE configs[0] at its stated size: 16 samples x 1000 tag x 5000 target SNPs, neighbors = 5 (NUM_REGIONS = 64,
    REGION_SIZE = 16). Reference keygen + encrypt write params.bin / keys.bin / encrypted_data.bin; the reference `cloud` and this
    repository's `cloud` read the same files: encrypted_prediction.bin must be byte-identical (all 15 000 ciphertexts, variances,
    record order). `decrypt bypos` of both: same rows, scores equal to 6 significant digits up to the reference FFT's +-1 LSB of
    the phase, and ours equal to the exact integer phase decoded as the reference decodes it."""
    if not po.have_ref() or not (po.REF_BIN / "keygen").exists():
        pytest.skip("oracle/_ref (the compiled reference) is not available on this box")
    subprocess.check_call(["make", "-C", str(ROOT / "idash2019_2_b200" / "host")], stdout=subprocess.DEVNULL)
    S, T, G, n = 16, 1000, 5000, 5
    tag, tgt = synth.make_positions(T, G, SEED)
    geno = synth.make_genotypes(T, S, SEED, na_frac=0.01)
    model = synth.make_model(tag, tgt, n, SEED)
    synth.write_tag_file(tmp_path / "tags.txt", tag, geno)
    synth.write_target_file(tmp_path / "targets.txt", tgt)
    synth.write_hr_dir(tmp_path / "model", model, tag, tgt)
    po.run_ref_bin("keygen", [tmp_path / "targets.txt", tmp_path / "tags.txt", 1], tmp_path)
    po.run_ref_bin("encrypt", [tmp_path / "tags.txt"], tmp_path)
    for d in ("ref", "b200"):
        (tmp_path / d).mkdir()
        for f in ("params.bin", "keys.bin", "encrypted_data.bin"):
            os.symlink(tmp_path / f, tmp_path / d / f)
    log = po.run_ref_bin("cloud", [tmp_path / "model"], tmp_path / "ref")
    assert "fhe wall time" in log
    r = subprocess.run([str(BIN / "cloud"), str(tmp_path / "model")], cwd=tmp_path / "b200", capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ref_img = (tmp_path / "ref" / "encrypted_prediction.bin").read_bytes()
    assert len(ref_img) == 8 + 3 * G * 8208
    assert (tmp_path / "b200" / "encrypted_prediction.bin").read_bytes() == ref_img

    po.run_ref_bin("decrypt", ["bypos"], tmp_path / "ref")
    r = subprocess.run([str(BIN / "decrypt"), "bypos"], cwd=tmp_path / "b200", capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ours, theirs = _fmt_rows(tmp_path / "b200" / "result_bypos.csv"), _fmt_rows(tmp_path / "ref" / "result_bypos.csv")
    assert len(ours) == len(theirs) == S * G
    a = np.array([ln.split(",")[2:] for ln in ours], dtype=np.float64)
    b = np.array([ln.split(",")[2:] for ln in theirs], dtype=np.float64)
    assert [ln.split(",")[:2] for ln in ours] == [ln.split(",")[:2] for ln in theirs]
    # one phase LSB is 2^-32; the csv keeps 6 significant digits (one unit of the last digit is at most 1e-5 of the value)
    tol = 2.0 ** -31 + 1.01e-5 * np.maximum(np.abs(a), np.abs(b))
    assert (np.abs(a - b) <= tol).all()
    assert (a == b).mean() > 0.8
    # ours == exact phase (TFHE's exact product) decoded and printed like the reference prints floats
    from idash2019_2_b200 import formats
    params, key, _ = formats.read_key(tmp_path / "keys.bin")
    idx, words, _ = formats.image_views(np.frombuffer(ref_img, np.uint8))
    order = np.argsort(idx)
    scores = po.decode_port(S, po.phase_exact_port(key, np.ascontiguousarray(words[order])))
    row_of = {int(v): i for i, v in enumerate(idx[order])}
    k = 0
    for s in range(S):
        for pos, bb in zip(params.out_positions, params.out_bidx):
            want = f"{s},{int(pos)}," + ",".join("%g" % float(scores[row_of[int(v)], s]) for v in bb)
            assert ours[k] == want, k
            k += 1
    shutil.rmtree(tmp_path / "model", ignore_errors=True)


@pytest.mark.parametrize("n_gpus,S", [(2, 1004), (2, 335), (4, 1004), (8, 1004)])
def test_cloud_sharded_over_gpus_equals_unsharded_equals_reference(built_lib, n_gpus, S):
    """SURVEY 8e on real GPUs: the target range cut over n GPUs of one process (idash_b200_cloud_eval_multi_device: peer-copied input
    slabs, rows stored straight into GPU 0's output array through the peer mapping) == the unsharded evaluation on GPU 0 == the
    reference, every word, index and variance; full iDASH size, neighbors = 20 (BASELINE configs[4] geometry). 335 samples
    (NUM_REGIONS = 3): the rotated staging and the zero tail (bulk stores through the peer mapping) on the other GPU."""
    import torch
    if torch.cuda.device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    n = 20
    geo = synth.Geometry(S, T_FULL, G_FULL)
    tag, tgt = synth.make_positions(T_FULL, G_FULL, SEED)
    model = synth.make_model(tag, tgt, n, SEED)
    n_in = geo.n_in_ct_used
    cts = synth.random_ciphertexts(n_in, SEED + 7)
    var = np.full(n_in, ALPHA2)
    var[::7] = 2.0 ** -48                               # not uniform: the per-row variance sums run on every GPU
    ctxs = [api.Context(g) for g in range(n_gpus)]
    m0 = api.Model(ctxs[0], S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef)
    models = [m0] + [m0.clone(ctxs[g]) for g in range(1, n_gpus)]
    with torch.cuda.device(0):
        x = torch.from_numpy(cts.view(np.int32)).cuda()
        xv = torch.from_numpy(var).cuda()
        n_rows = 3 * G_FULL
        whole = torch.zeros((n_rows, 2048), dtype=torch.int32, device="cuda")
        wvar = torch.zeros(n_rows, dtype=torch.float64, device="cuda")
        widx = torch.zeros(n_rows, dtype=torch.int32, device="cuda")
        api.cloud_compute_score_device(ctxs[0], m0, x, whole, in_var=xv, out_index=widx, out_var=wvar)
        torch.cuda.synchronize()
        out = torch.zeros_like(whole)
        ovar = torch.zeros_like(wvar)
        oidx = torch.zeros_like(widx)
        api.cloud_compute_score_multi_device(ctxs, models, x, out, in_var=xv, out_index=oidx, out_var=ovar)
        assert torch.equal(out, whole) and torch.equal(ovar, wvar) and torch.equal(oidx, widx)
        ref_out, ref_var, _ = _reference(S, geo.NR, geo.RS, cts, var, model)
        assert _equal_on_device(out, ref_out) == 0
        assert np.array_equal(ovar.cpu().numpy(), ref_var)
        assert np.array_equal(oidx.cpu().numpy().view(np.uint32), model.out_bidx)
    for m in models:
        m.free()
    for c in ctxs:
        c.close()
