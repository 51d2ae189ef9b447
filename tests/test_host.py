"""The C++ host layer (idash2019_2_b200/host): the reference's eval/idash.h API for the cloud / decrypt stages over
the C ABI. CPU part: file formats and container plumbing against the golden fixtures written by the reference
binaries. GPU part: the `cloud` and `decrypt` binaries run in a directory laid out like the reference's eval/run."""
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from idash2019_2_b200 import formats
from oracle import pyoracle as po

from helpers import GOLDEN, GOLDEN_CASES, load_golden

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "idash2019_2_b200" / "lib" / "bin"


@pytest.fixture(scope="session")
def host_bins(built_lib):
    if shutil.which("make") and Path("/usr/bin/g++").exists():
        subprocess.check_call(["make", "-C", str(ROOT / "idash2019_2_b200" / "host")], stdout=subprocess.DEVNULL)
    for b in ("cloud", "decrypt", "host_selftest"):
        assert (BIN / b).exists(), f"{BIN / b} is not built"
    return BIN


def _s32(v: int) -> int:
    return v - (1 << 32) if v >= (1 << 31) else v


def fmt_g(x) -> str:
    return "%g" % float(x)          # what `ostream << float` prints (6 significant digits)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_host_formats_roundtrip(host_bins, tmp_path, name):
    d, params, key, enc, pred, ref = load_golden(name)
    out = subprocess.run([str(host_bins / "host_selftest"), str(d), str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr

    # params.bin: geometry + targets in file order with their bigIndices
    lines = (tmp_path / "params.txt").read_text().splitlines()
    assert [int(x) for x in lines[0].split()] == [params.NUM_SAMPLES, params.NUM_INPUT_POSITIONS, params.NUM_OUTPUT_POSITIONS,
                                                   params.NUM_INPUT_FEATURES, params.NUM_OUTPUT_FEATURES, params.NUM_REGIONS,
                                                   params.REGION_SIZE]
    for ln, pos, nm, b in zip(lines[1:], params.out_positions, params.out_names, params.out_bidx):
        assert ln.split() == [str(int(pos)), nm] + [str(int(x)) for x in b]

    # read_model == the reference's read_model (CSR stored in ref.npz by make_golden.py)
    got = {}
    for ln in (tmp_path / "model.txt").read_text().splitlines():
        o, i, c = ln.split()
        got.setdefault(int(o), {})[int(i)] = int(c)
    want = {}
    for r, o in enumerate(ref["model_out_bidx"]):
        lo, hi = int(ref["model_row_ptr"][r]), int(ref["model_row_ptr"][r + 1])
        want[int(o)] = {int(i): int(c) for i, c in zip(ref["model_col"][lo:hi], ref["model_coef"][lo:hi])}
    assert got == want

    # keys.bin
    kl = (tmp_path / "key.txt").read_text().split()
    assert [int(x) for x in kl[:3]] == [params.NUM_SAMPLES, 1024, 1]
    assert np.array_equal(np.array(kl[3:], dtype=np.int64), key)

    # ciphertext files: every record survives read -> containers -> write (keyed by index: the record order of a
    # re-written file is the hash map's iteration order, in the reference too)
    for src, rt in ((enc, "encrypted_data.rt.bin"), (pred, "encrypted_prediction.rt.bin")):
        img = formats.read_ct_image(tmp_path / rt)
        i0, w0, v0 = formats.image_views(src)
        i1, w1, v1 = formats.image_views(img)
        o0, o1 = np.argsort(i0), np.argsort(i1)
        assert np.array_equal(i0[o0], i1[o1]) and np.array_equal(w0[o0], w1[o1]) and np.array_equal(v0[o0], v1[o1])

    # result csv: sample-major, targets in file order, floats like operator<<(float)
    for fn, by_pos in (("result.csv", False), ("result_bypos.csv", True)):
        rows = (tmp_path / fn).read_text().splitlines()
        assert rows[0] == "Subject ID,target SNP,0,1,2"
        assert len(rows) == 1 + params.NUM_SAMPLES * params.NUM_OUTPUT_POSITIONS
        k = 1
        for s in range(params.NUM_SAMPLES):
            for pos, nm in zip(params.out_positions, params.out_names):
                sc = [np.float32(_s32(((int(pos) & 0xFFFFFFFF) * 2654435761 + v * 40503 + s * 2246822519) & 0xFFFFFFFF) / 2.0 ** 32)
                      for v in range(3)]
                label = str(int(pos)) if by_pos else nm
                assert rows[k] == f"{s},{label}," + ",".join(fmt_g(x) for x in sc), (fn, k)
                k += 1


def test_parse_vw_semantics(host_bins, tmp_path):
    """sscanf("%s %d") semantics of eval/parse_vw.cpp:18-25 through read_model: integer prefix of a float, last
    duplicate wins, blank lines ignored, Constant -> 0xFFFFFFFF."""
    d = tmp_path / "case"
    shutil.copytree(GOLDEN / "s1004_nr1", d)
    params = formats.read_params(d / "params.bin")
    pos = int(params.out_positions[0])
    tag = int(params.in_positions[0])
    tb = params.in_index()[tag]
    (d / "model" / f"{pos}_0.hr").write_text(f"Constant -12.9\n\n{tag}_1 7.99\n  {tag}_2\t-3\n{tag}_1 +5.0\n")
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([str(host_bins / "host_selftest"), str(d), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ob = int(params.out_bidx[0][0])
    got = {int(i): int(c) for o, i, c in (ln.split() for ln in (out / "model.txt").read_text().splitlines()) if int(o) == ob}
    assert got == {0xFFFFFFFF: -12, int(tb[1]): 5, int(tb[2]): -3}


def test_host_errors_die_dramatically(host_bins, tmp_path):
    """eval/idash.h:12-14 convention: "ERROR: ..." on stdout and abort(); a missing .hr file -> stderr + abort()."""
    r = subprocess.run([str(host_bins / "host_selftest"), str(tmp_path), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode < 0 and "ERROR: Cannot open parameters file for read" in r.stdout
    d = tmp_path / "case"
    shutil.copytree(GOLDEN / "s1004_nr1", d)
    next((d / "model").glob("*.hr")).unlink()
    r = subprocess.run([str(host_bins / "host_selftest"), str(d), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode < 0 and "Cannot open file" in r.stderr


@pytest.mark.parametrize("env", [{}, {"IDASH_HOST_NO_MMAP": "1"}, {"IDASH_HOST_THREADS": "1"}, {"IDASH_HOST_THREADS": "3", "IDASH_HOST_NO_WARMUP": "1"}])
def test_host_bench_io_paths(host_bins, tmp_path, env):
    """host_bench on a golden pipeline directory: every file-level phase runs without a GPU, and the slab image written by
    the mapped / pwrite paths (any thread count) reads back identical; the result csv is written by the batch formatter."""
    d = GOLDEN / "s335_nr3"
    r = subprocess.run([str(host_bins / "host_bench"), str(d), str(d / "model"), str(tmp_path), "csv"], capture_output=True, text=True,
                       env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "slab image round trip: identical" in r.stdout
    for phase in ("read_model", "read_encrypted_data", "read_encrypted_predictions", "write (slab image)", "write_decrypted_predictions"):
        assert phase in r.stdout
    params = formats.read_params(d / "params.bin")
    rows = (tmp_path / "result_bypos.csv").read_text().splitlines()
    assert rows[0] == "Subject ID,target SNP,0,1,2" and len(rows) == 1 + params.NUM_SAMPLES * params.NUM_OUTPUT_POSITIONS
    s, pos = params.NUM_SAMPLES - 1, int(params.out_positions[-1])
    sc = [np.float32(_s32(((pos & 0xFFFFFFFF) * 2654435761 + v * 40503 + s * 2246822519) & 0xFFFFFFFF) / 2.0 ** 32) for v in range(3)]
    assert rows[-1] == f"{s},{pos}," + ",".join(fmt_g(x) for x in sc)


def _stage_dir(tmp_path, d):
    for f in ("params.bin", "keys.bin", "encrypted_data.bin"):
        shutil.copy(d / f, tmp_path / f)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cloud_binary_reproduces_reference_file(host_bins, tmp_path, name):
    """`cloud <model dir>` in a reference-style run directory writes the SAME encrypted_prediction.bin as the reference
    binary: every ciphertext word, every variance and the record order (libstdc++ hash-map order)."""
    d = GOLDEN / name
    _stage_dir(tmp_path, d)
    r = subprocess.run([str(host_bins / "cloud"), str(d / "model")], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "fhe wall time (seconds)..........:" in r.stdout and "RAM usage (MB)" in r.stdout
    assert (tmp_path / "encrypted_prediction.bin").read_bytes() == (d / "encrypted_prediction.bin").read_bytes()


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_decrypt_binary_matches_exact_oracle(host_bins, tmp_path, name):
    """`decrypt bypos` on the reference's encrypted_prediction.bin: csv rows equal the exact integer phase decoded and
    printed like the reference prints floats."""
    d, params, key, enc, pred, ref = load_golden(name)
    _stage_dir(tmp_path, d)
    shutil.copy(d / "encrypted_prediction.bin", tmp_path / "encrypted_prediction.bin")
    r = subprocess.run([str(host_bins / "decrypt"), "bypos"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "decrypt wall time (seconds)......:" in r.stdout
    S = params.NUM_SAMPLES
    scores = po.decode_port(S, ref["phase_exact"])                     # rows sorted by output bigIndex
    row_of = {int(b): i for i, b in enumerate(ref["pred_index_sorted"])}
    rows = (tmp_path / "result_bypos.csv").read_text().splitlines()
    assert rows[0] == "Subject ID,target SNP,0,1,2" and len(rows) == 1 + S * params.NUM_OUTPUT_POSITIONS
    k = 1
    for s in range(S):
        for pos, b in zip(params.out_positions, params.out_bidx):
            want = f"{s},{int(pos)}," + ",".join(fmt_g(scores[row_of[int(x)], s]) for x in b)
            assert rows[k] == want, k
            k += 1


@pytest.mark.gpu
def test_cloud_binary_dies_on_missing_input(host_bins, tmp_path):
    """A model that needs a ciphertext the data file lacks: the reference aborts with "shit happens before"."""
    d = GOLDEN / "s1004_nr1"
    _stage_dir(tmp_path, d)
    img = formats.read_ct_image(d / "encrypted_data.bin")
    idx, words, var = formats.image_views(img)
    keep = np.arange(len(idx)) != 0
    formats.build_ct_image(idx[keep], words[keep], var[keep]).tofile(tmp_path / "encrypted_data.bin")
    r = subprocess.run([str(host_bins / "cloud"), str(d / "model")], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode < 0 and "ERROR: shit happens before" in r.stdout


@pytest.mark.gpu
def test_cloud_binary_model_cache(host_bins, tmp_path):
    """models.bin: the first run compiles the .hr files and stores the device layout, the second run takes it from
    the cache -- same bytes out; an edited .hr file changes the fingerprint, so the third run parses again (and sees the edit)."""
    d = GOLDEN / "s1004_nr1"
    _stage_dir(tmp_path, d)
    shutil.copytree(d / "model", tmp_path / "model")
    env = dict(os.environ, IDASH_HOST_TIMING="1")
    want = (d / "encrypted_prediction.bin").read_bytes()
    for k in range(2):
        r = subprocess.run([str(host_bins / "cloud"), "model"], cwd=tmp_path, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        assert ("compiled model taken from models.bin" in r.stderr) == (k == 1), r.stderr
        assert (tmp_path / "encrypted_prediction.bin").read_bytes() == want
    assert (tmp_path / "models.bin").stat().st_size > 0
    # edit one coefficient (the Constant of the first file): the fingerprint covers size + mtime of every .hr file
    f = sorted((tmp_path / "model").glob("*.hr"))[0]
    lines = f.read_text().splitlines()
    k = next(i for i, ln in enumerate(lines) if ln.startswith("Constant "))
    lines[k] = "Constant %.1f" % (float(lines[k].split()[1]) + 3.0)
    f.write_text("\n".join(lines) + "\n")
    r = subprocess.run([str(host_bins / "cloud"), "model"], cwd=tmp_path, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "compiled model taken from" not in r.stderr
    got = (tmp_path / "encrypted_prediction.bin").read_bytes()
    assert len(got) == len(want) and got != want
    # ... and with the cache switched off nothing is written
    (tmp_path / "models.bin").unlink()
    r = subprocess.run([str(host_bins / "cloud"), "model"], cwd=tmp_path, capture_output=True, text=True, env=dict(env, IDASH_MODEL_CACHE="off"))
    assert r.returncode == 0 and not (tmp_path / "models.bin").exists()
    assert (tmp_path / "encrypted_prediction.bin").read_bytes() == got


def _mid_size_run_dir(tmp_path, S=1004, T=600, G=3000, n=5, seed=77):
    """reference keygen + encrypt + cloud on a synthetic case large enough for several tiles per GPU (3 G = 9000 rows = 141 tiles)"""
    from idash2019_2_b200 import synth
    if not po.have_ref() or not (po.REF_BIN / "keygen").exists():
        pytest.skip("oracle/_ref (the compiled reference) is not available on this box")
    tag, tgt = synth.make_positions(T, G, seed)
    geno = synth.make_genotypes(T, S, seed, na_frac=0.01)
    model = synth.make_model(tag, tgt, n, seed)
    synth.write_tag_file(tmp_path / "tags.txt", tag, geno)
    synth.write_target_file(tmp_path / "targets.txt", tgt)
    synth.write_hr_dir(tmp_path / "model", model, tag, tgt)
    po.run_ref_bin("keygen", [tmp_path / "targets.txt", tmp_path / "tags.txt", 1], tmp_path)
    po.run_ref_bin("encrypt", [tmp_path / "tags.txt"], tmp_path)
    for sub in ("ref", "b200"):
        (tmp_path / sub).mkdir()
        for f in ("params.bin", "keys.bin", "encrypted_data.bin"):
            os.symlink(tmp_path / f, tmp_path / sub / f)
    po.run_ref_bin("cloud", [tmp_path / "model"], tmp_path / "ref")
    return (tmp_path / "ref" / "encrypted_prediction.bin").read_bytes()


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", ["0", "0,1", "0,1,2,3"])
def test_cloud_binary_target_ranges_on_several_gpus(host_bins, tmp_path, gpus):
    """IDASH_GPUS: the target range is cut at tile boundaries, every GPU copies only its part of the index-sorted input slab and
    writes its rows of the output slab (eval/idash.cpp:779-790 is the loop being cut). Same file as the reference's
    cloud, byte for byte -- also with one GPU, which takes the same sorted-slab + pipelined path."""
    import torch
    n_gpus = len(gpus.split(","))
    if torch.cuda.device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    want = _mid_size_run_dir(tmp_path)
    r = subprocess.run([str(host_bins / "cloud"), str(tmp_path / "model")], cwd=tmp_path / "b200", capture_output=True, text=True,
                       env=dict(os.environ, IDASH_HOST_TIMING="1", IDASH_GPUS=gpus, IDASH_MODEL_CACHE="off"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"evaluation on {n_gpus} GPU(s)" in r.stderr, r.stderr
    assert (tmp_path / "b200" / "encrypted_prediction.bin").read_bytes() == want
    shutil.rmtree(tmp_path / "model", ignore_errors=True)


def test_reference_glue_compiles_against_reference_headers():
    """INTEGRATION.md section B is code, not prose: host/reference_glue/idash_b200_glue.cpp -- new bodies of cloud_compute_score /
    decrypt_predictions over the C ABI, written against the reference's own eval/idash.h and <tfhe.h> -- must type-check against the
    reference tree (g++ -fsyntax-only; `-include array` is the same work-around the reference build needs), and the
    snippet printed in INTEGRATION.md must be that file."""
    glue = ROOT / "idash2019_2_b200" / "host" / "reference_glue" / "idash_b200_glue.cpp"
    body = glue.read_text()
    doc = (ROOT / "INTEGRATION.md").read_text()
    for line in ("idash_b200_cloud_eval_host(ctx(), m, &in, &out, nullptr)", "idash_b200_decrypt_host(ctx(), key.tlweKey->key[0].coefs, S, &in, scores, nullptr)",
                 "TLweSample *s = enc_preds.createAndGet(out_bidx[r], params.tlweParams);"):
        assert line in body and line in doc
    ref = Path("/root/reference")
    if not (ref / "eval" / "idash.h").exists() or not Path("/usr/bin/g++").exists():
        pytest.skip("the reference tree is not on this box")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-include", "array", f"-I{ref / 'eval'}", f"-I{ref / 'tfhe' / 'src' / 'include'}",
                        f"-I{ROOT / 'include'}", str(glue)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
