"""The oracle (oracle/idash_oracle.c) pinned against the reference: golden vectors produced by the
unmodified reference binaries (tests/golden/make_golden.py), and -- where oracle/_ref is built -- the
reference functions themselves on fresh seeded inputs."""
import numpy as np
import pytest

from idash2019_2_b200 import formats
from oracle import pyoracle as po

from helpers import ALPHA2, GOLDEN_CASES, load_golden, make_case

needs_ref = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_cloud_matches_reference_cloud_binary(name):
    d, params, key, enc, pred, ref = load_golden(name)
    ob, rp, col, coef = formats.read_model(params, d / "model")
    in_idx, in_words, in_var = formats.image_views(enc)
    out_ct, out_var = po.cloud_port(params.NUM_SAMPLES, params.NUM_REGIONS, params.REGION_SIZE, in_idx.copy(),
                                    np.ascontiguousarray(in_words), in_var.copy(), rp, col, coef)
    p_idx, p_words, p_var = formats.image_views(pred)
    order = np.argsort(p_idx)
    assert np.array_equal(p_idx[order], ob)
    assert np.array_equal(out_ct, p_words[order])          # every word of every output ciphertext
    assert np.array_equal(out_var, p_var[order])           # the serialized current_variance, bit for bit


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_phase_and_decode_match_reference(name):
    d, params, key, enc, pred, ref = load_golden(name)
    p_idx, p_words, _ = formats.image_views(pred)
    order = np.argsort(p_idx)
    ct = np.ascontiguousarray(p_words[order])
    phase = po.phase_exact_port(key, ct)
    assert np.array_equal(phase, ref["phase_exact"])       # TFHE's exact Karatsuba product
    diff = (phase.astype(np.int64) - ref["phase_fft"].astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert np.abs(diff).max() <= 1                          # the reference's FFT decrypt: +-1 LSB
    S = params.NUM_SAMPLES
    assert np.array_equal(po.decode_port(S, ref["phase_fft"]), ref["scores"])   # decode rule, bit for bit
    # decoded plaintext: the model applied to the genotypes (semantic sanity of the whole pipeline)
    scores = po.decode_port(S, phase)
    assert np.isfinite(scores).all()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_decrypted_scores_equal_plaintext_model(name):
    """decrypt(cloud(encrypt(x))) ~= model(x) / 16384: ties the ciphertext arithmetic to its meaning."""
    d, params, key, enc, pred, ref = load_golden(name)
    S = params.NUM_SAMPLES
    ob, rp, col, coef = formats.read_model(params, d / "model")
    geno = ref["genotypes"]                                # [T, S], -1 = NA
    onehot = np.zeros((3 * geno.shape[0], S))
    for v in range(3):
        onehot[v::3] = (geno == v)
    na = geno < 0
    for v, w in enumerate((3 / 6, 2 / 6, 1 / 6)):          # eval/idash.cpp:32-37
        onehot[v::3][na] = w
    scores = po.decode_port(S, ref["phase_exact"])
    for r in range(len(ob)):
        acc = np.zeros(S)
        for e in range(int(rp[r]), int(rp[r + 1])):
            acc += coef[e] * (1.0 if col[e] == 0xFFFFFFFF else onehot[col[e]])
        d = (scores[r] - acc / 16384 + 0.5) % 1.0 - 0.5       # the torus wraps: large coefficients overflow [-1/2, 1/2)
        sigma = np.sqrt(sum(float(coef[e]) ** 2 for e in range(int(rp[r]), int(rp[r + 1])) if col[e] != 0xFFFFFFFF)) * 2.0 ** -25
        assert np.abs(d).max() < 6 * sigma + 1e-6             # noise std = alpha * |coefs|_2


@needs_ref
@pytest.mark.parametrize("S,n,cr", [(1004, 5, 200), (1024, 3, 8191), (513, 5, 200), (512, 5, 8191), (400, 7, 200),
                                    (335, 5, 8191), (100, 4, 200), (16, 5, 8191), (1, 2, 200)])
def test_oracle_cloud_matches_reference_function(S, n, cr):
    geo, model, cts, var = make_case(S, T=40, G=64, n=n, seed=S + n, coef_range=cr, bias_range=cr)
    idx = np.arange(len(cts), dtype=np.uint32)
    o1, v1 = po.cloud_port(S, geo.NR, geo.RS, idx, cts, var, model.row_ptr, model.col, model.coef)
    o2, v2, _ = po.cloud_ref(S, geo.NR, geo.RS, idx, cts, var, model.out_bidx, model.row_ptr, model.col, model.coef)
    assert np.array_equal(o1, o2) and np.array_equal(v1, v2)


@needs_ref
def test_oracle_cloud_extreme_coefficients_and_permuted_inputs():
    """int32-range coefficients (wraparound incl. the int32 p*p of the variance) and shuffled input slots."""
    S = 400
    geo, model, cts, var = make_case(S, T=30, G=30, n=5, seed=99)
    rng = np.random.default_rng(1)
    coef = rng.integers(-2 ** 31, 2 ** 31, size=model.nnz).astype(np.int32)
    perm = rng.permutation(len(cts)).astype(np.uint32)          # slot i holds ciphertext index perm[i]
    cts_p = np.ascontiguousarray(cts[perm])
    o1, v1 = po.cloud_port(S, geo.NR, geo.RS, perm, cts_p, var, model.row_ptr, model.col, coef)
    o2, v2, _ = po.cloud_ref(S, geo.NR, geo.RS, perm, cts_p, var, model.out_bidx, model.row_ptr, model.col, coef)
    assert np.array_equal(o1, o2) and np.array_equal(v1, v2)


@needs_ref
def test_oracle_phase_matches_reference_karatsuba_and_fft():
    rng = np.random.default_rng(2)
    key = rng.integers(0, 2, 1024).astype(np.int32)
    ct = rng.integers(0, 2 ** 32, size=(12, 2048), dtype=np.uint32)
    ph = po.phase_exact_port(key, ct)
    assert np.array_equal(ph, po.phase_ref(key, ct, use_fft=False))
    d = (ph.astype(np.int64) - po.phase_ref(key, ct, use_fft=True).astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert np.abs(d).max() <= 1
    sc, _ = po.decrypt_ref(1004, key, ct)
    assert np.array_equal(sc, po.decode_port(1004, po.phase_ref(key, ct, use_fft=True)))


def test_oracle_missing_input_raises():
    geo, model, cts, var = make_case(1004, T=10, G=6, n=3, seed=5)
    idx = np.arange(len(cts), dtype=np.uint32)
    with pytest.raises(KeyError):
        po.cloud_port(1004, 1, 1024, idx[:-3], cts[:-3], var[:-3], model.row_ptr, model.col, model.coef)


def test_oracle_empty():
    out, var = po.cloud_port(1004, 1, 1024, np.zeros(0, np.uint32), np.zeros((0, 2048), np.uint32), np.zeros(0),
                             np.zeros(1, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.int32))
    assert out.shape == (0, 2048) and var.shape == (0,)


def test_constants():
    """ONE / NaN encodings (eval/idash.cpp:29-45) as computed by the reference, stored in the golden files."""
    ref = np.load(load_golden("s1004_nr1")[0] / "ref.npz")
    assert list(ref["constants"]) == [262144, 131072, 87381, 43690]
    assert ALPHA2 == 2.0 ** -50
