"""Parity of the CUDA decrypt/decode kernel with the exact oracle and the reference's golden output."""
import numpy as np
import pytest

from idash2019_2_b200 import api, formats
from oracle import pyoracle as po

from helpers import GOLDEN_CASES, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["pair", "tensor", "iadd"], autouse=True)
def decrypt_kernel(request, gpu_ctx):
    """Every test runs on all decrypt kernels (tcgen05 Toeplitz GEMM on CTA pairs and on single CTAs, and the CUDA-core kernel)."""
    which = {"pair": api.DECRYPT_TENSOR_PAIR, "tensor": api.DECRYPT_TENSOR, "iadd": api.DECRYPT_IADD}[request.param]
    gpu_ctx.set_decrypt_kernel(which)
    yield which
    gpu_ctx.set_decrypt_kernel(api.DECRYPT_AUTO)


@pytest.mark.parametrize("S", [1004, 1024, 400, 335, 16, 1])
def test_decrypt_matches_exact_oracle(gpu_ctx, decrypt_kernel, S):
    rng = np.random.default_rng(S)
    key = rng.integers(0, 2, 1024).astype(np.int32)
    ct = rng.integers(0, 2 ** 32, size=(37, 2048), dtype=np.uint32)
    scores, phase = api.decrypt_predictions(gpu_ctx, key, S, ct, want_phase=True)
    ref_phase = po.phase_exact_port(key, ct)
    assert np.array_equal(phase, ref_phase)                       # bit-exact, every coefficient
    assert np.array_equal(scores, po.decode_port(S, ref_phase))   # decoded floats bit-equal
    assert gpu_ctx.last_decrypt_kernel() == decrypt_kernel


@pytest.mark.parametrize("n_ct,grid,slots", [(32, 1, 12), (64, 1, 8), (333, 2, 12), (1000, 3, 9), (5000, 148, 12)])
def test_decrypt_many_groups_per_cta(gpu_ctx, monkeypatch, n_ct, grid, slots):
    """The persistent tensor-core kernel with several 32-ciphertext groups per CTA: ring slots and TMEM stages wrap
    (IDASH_B200_DECRYPT_GRID / _SLOTS shrink the grid and the ring), records layout with its 8208-byte stride."""
    monkeypatch.setenv("IDASH_B200_DECRYPT_GRID", str(grid))
    monkeypatch.setenv("IDASH_B200_DECRYPT_SLOTS", str(slots))
    rng = np.random.default_rng(n_ct)
    key = rng.integers(0, 2, 1024).astype(np.int32)
    ct = rng.integers(0, 2 ** 32, size=(n_ct, 2048), dtype=np.uint32)
    ct[0] = 0xFFFFFFFF                                            # every byte plane at its maximum
    ct[-1, :1024] = 0x80000000
    ref_phase = po.phase_exact_port(key, ct)
    scores, phase = api.decrypt_predictions(gpu_ctx, key, 1004, ct, want_phase=True)
    assert np.array_equal(phase, ref_phase)
    assert np.array_equal(scores, po.decode_port(1004, ref_phase))
    image = formats.build_ct_image(np.arange(n_ct, dtype=np.uint32), ct, np.zeros(n_ct))
    _, scores_r, phase_r = api.decrypt_predictions_records(gpu_ctx, key, 1004, image, want_phase=True)
    assert np.array_equal(phase_r, ref_phase) and np.array_equal(scores_r, scores)
    scores_only = api.decrypt_predictions(gpu_ctx, key, 1004, ct)
    assert np.array_equal(scores_only, scores)


@pytest.mark.parametrize("key_kind", ["zeros", "ones", "single"])
def test_decrypt_edge_keys(gpu_ctx, key_kind):
    rng = np.random.default_rng(1)
    key = {"zeros": np.zeros(1024, np.int32), "ones": np.ones(1024, np.int32),
           "single": np.eye(1, 1024, 1023, dtype=np.int32)[0]}[key_kind]
    ct = rng.integers(0, 2 ** 32, size=(5, 2048), dtype=np.uint32)
    _, phase = api.decrypt_predictions(gpu_ctx, key, 1004, ct, want_phase=True)
    assert np.array_equal(phase, po.phase_exact_port(key, ct))


def test_decrypt_rejects_non_binary_key(gpu_ctx):
    key = np.zeros(1024, np.int32)
    key[3] = 2
    with pytest.raises(api.IdashB200Error):
        api.decrypt_predictions(gpu_ctx, key, 16, np.zeros((1, 2048), np.uint32))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_decrypt_golden_prediction_file(gpu_ctx, name):
    """encrypted_prediction.bin of the reference -> phases equal TFHE's exact Karatsuba product, within 1 LSB of
    the reference's FFT decrypt; decoded scores equal the reference's except where that LSB shows."""
    d, params, key, enc, pred, ref = load_golden(name)
    S = params.NUM_SAMPLES
    index, scores, phase = api.decrypt_predictions_records(gpu_ctx, key, S, pred, want_phase=True)
    order = np.argsort(index)
    assert np.array_equal(index[order], ref["pred_index_sorted"])
    assert np.array_equal(phase[order], ref["phase_exact"])
    diff = (phase[order].astype(np.int64) - ref["phase_fft"].astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert np.abs(diff).max() <= 1
    assert np.array_equal(scores[order], po.decode_port(S, ref["phase_exact"]))
    # vs the reference binary's floats: identical where the FFT phase is exact, else the float of a phase 1 LSB away
    ref_sc = ref["scores"]
    same_phase = (diff == 0)[:, :S]
    assert np.array_equal(scores[order][same_phase], ref_sc[same_phase])
    lo = po.decode_port(S, (phase[order].astype(np.int64) - 1).astype(np.uint32))
    hi = po.decode_port(S, (phase[order].astype(np.int64) + 1).astype(np.uint32))
    assert ((ref_sc == scores[order]) | (ref_sc == lo) | (ref_sc == hi)).all()


def test_decrypt_device_path_properties_at_scale(gpu_ctx, decrypt_kernel):
    """BASELINE size (242 646 ciphertexts, S = 1004) on device tensors, through properties that need no oracle pass over 2 GB:
    the phase is linear in the ciphertext (mod 2^32), the zero key returns b, scores are the decode of the phases, and a
    strided sample equals the exact oracle."""
    import torch
    if decrypt_kernel == api.DECRYPT_IADD:
        pytest.skip("the CUDA-core kernel takes 17 ms per pass; the property run is for the tensor-core kernel")
    n, S = 242646, 1004
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randint(-2 ** 31, 2 ** 31, (n, 2048), dtype=torch.int32, device="cuda", generator=g)
    y = torch.randint(-2 ** 31, 2 ** 31, (n, 2048), dtype=torch.int32, device="cuda", generator=g)
    key = np.random.default_rng(3).integers(0, 2, 1024).astype(np.int32)
    ph = [torch.empty((n, 1024), dtype=torch.int32, device="cuda") for _ in range(3)]
    sc = torch.empty((n, S), dtype=torch.float32, device="cuda")
    api.decrypt_predictions_device(gpu_ctx, key, S, x, sc, ph[0])
    api.decrypt_predictions_device(gpu_ctx, key, S, y, None, ph[1])
    api.decrypt_predictions_device(gpu_ctx, key, S, x + y, None, ph[2])        # int32 adds wrap: the sum mod 2^32
    torch.cuda.synchronize()
    assert torch.equal(ph[2], ph[0] + ph[1])
    assert torch.equal(sc, (ph[0][:, :S].to(torch.float64) / 2.0 ** 32).to(torch.float32))
    api.decrypt_predictions_device(gpu_ctx, np.zeros(1024, np.int32), S, x, None, ph[1])
    torch.cuda.synchronize()
    assert torch.equal(ph[1], x[:, 1024:])
    idx = torch.arange(0, n, 4099, device="cuda")
    ref = po.phase_exact_port(key, x[idx].cpu().numpy().view(np.uint32))
    assert np.array_equal(ph[0][idx].cpu().numpy().view(np.uint32), ref)


def test_decrypt_empty(gpu_ctx):
    s = api.decrypt_predictions(gpu_ctx, np.zeros(1024, np.int32), 16, np.zeros((0, 2048), np.uint32))
    assert s.shape == (0, 16)
