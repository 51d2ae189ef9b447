"""Parity of the CUDA cloud evaluator (through the C ABI) with the oracle / the reference's golden
output. Bit-exact on every ciphertext word and on the serialized variance."""
import numpy as np
import pytest

from idash2019_2_b200 import _lib, api, formats, synth
from oracle import pyoracle as po

from helpers import ALPHA2, GOLDEN_CASES, load_golden, make_case

pytestmark = pytest.mark.gpu

KERNELS = {"imad": _lib.KERNEL_IMAD, "tile": _lib.KERNEL_TENSOR_TILE, "ring": _lib.KERNEL_TENSOR_RING}


@pytest.fixture(params=["imad", "tile", "ring"])
def kctx(request, gpu_ctx):
    """The shared context with the cloud kernel forced to the IMAD one or to one of the two tcgen05 (tensor-core)
    schedules. Tests call _model(kctx, ...) which skips models the forced kernel cannot take."""
    gpu_ctx.requested = KERNELS[request.param]
    gpu_ctx.set_kernel(gpu_ctx.requested)
    yield gpu_ctx
    gpu_ctx.set_kernel(_lib.KERNEL_AUTO)
    gpu_ctx.requested = None


def _model(ctx, S, NR, RS, out_bidx, row_ptr, col, coef):
    m = api.Model(ctx, S, NR, RS, out_bidx, row_ptr, col, coef)
    req = getattr(ctx, "requested", None)
    if req == _lib.KERNEL_TENSOR_RING and not m.info["ring_ok"]:
        m.free()
        pytest.skip("model not eligible for the persistent ring kernel")
    if req == _lib.KERNEL_TENSOR_TILE and not m.info["n_tiles"]:
        m.free()
        pytest.skip("model not eligible for the tensor-core kernels")
    return m


def _used(ctx):
    req = getattr(ctx, "requested", None)
    if req is not None:
        assert ctx.last_kernel() == req


def _oracle(S, geo, model, cts, var, idx=None):
    idx = np.arange(len(cts), dtype=np.uint32) if idx is None else idx
    return po.cloud_port(S, geo.NR, geo.RS, idx, cts, var, model.row_ptr, model.col, model.coef)


@pytest.mark.parametrize("S,n,cr", [(1004, 5, 200), (1024, 2, 8191), (1004, 50, 200), (513, 5, 200), (512, 5, 8191),
                                    (400, 20, 200), (335, 5, 8191), (335, 20, 200), (100, 4, 200), (16, 5, 8191), (1, 2, 50)])
def test_cloud_packed_matches_oracle(kctx, S, n, cr):
    geo, model, cts, var = make_case(S, T=60, G=101, n=n, seed=S + n, coef_range=cr, bias_range=cr)
    m = _model(kctx, S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef)
    out, idx, ovar = api.cloud_compute_score(kctx, m, cts, in_var=var)
    ref_out, ref_var = _oracle(S, geo, model, cts, var)
    assert np.array_equal(out, ref_out)
    assert np.array_equal(ovar, ref_var)
    assert np.array_equal(idx, model.out_bidx)
    _used(kctx)
    m.free()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cloud_records_reproduces_reference_file(kctx, name):
    """encrypted_data.bin image in, encrypted_prediction.bin image out: byte-identical to the file the
    reference `cloud` binary wrote (same record order via slot_of_row)."""
    d, params, key, enc, pred, ref = load_golden(name)
    ob, rp, col, coef = formats.read_model(params, d / "model")
    m = _model(kctx, params.NUM_SAMPLES, params.NUM_REGIONS, params.REGION_SIZE, ob, rp, col, coef)
    p_idx, _, _ = formats.image_views(pred)
    slot_of_bidx = {int(b): s for s, b in enumerate(p_idx)}
    slot_of_row = np.array([slot_of_bidx[int(b)] for b in ob], np.uint32)
    out_img = api.cloud_compute_score_records(kctx, m, enc, slot_of_row, formats.aligned_image(len(ob)))
    assert out_img.tobytes() == pred.tobytes()
    _used(kctx)
    m.free()


@pytest.mark.parametrize("kind", ["pow2", "zero", "uniform_not_pow2", "one_differs", "packed_no_variance"])
def test_cloud_output_variance_paths(gpu_ctx, kind):
    """The serialized variance (tlwe-functions.cpp:175): one multiply when every input carries the same power-of-two variance
    (the reference's encrypt: 2^-50), the per-entry walk otherwise -- same doubles either way, and as the oracle's."""
    S = 1004
    geo, model, cts, var = make_case(S, T=70, G=120, n=20, seed=77)
    if kind == "zero":
        var = np.zeros_like(var)
    elif kind == "uniform_not_pow2":
        var = var * 3.0
    elif kind == "one_differs":
        var = var.copy(); var[len(var) // 2] *= 2.0
    m = _model(gpu_ctx, S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef)
    out, _, ovar = api.cloud_compute_score(gpu_ctx, m, cts, in_var=None if kind == "packed_no_variance" else var)
    ref_out, ref_var = _oracle(S, geo, model, cts, var)
    assert np.array_equal(out, ref_out)
    assert np.array_equal(ovar, ref_var)
    image = formats.build_ct_image(np.arange(len(cts), dtype=np.uint32), cts, var)
    rec = api.cloud_compute_score_records(gpu_ctx, m, image)
    _, w, v = formats.image_views(rec)
    assert np.array_equal(w, ref_out) and np.array_equal(v, ref_var)
    m.free()


@pytest.mark.parametrize("S,n_batches,G", [(1004, 2, 3000), (1004, 5, 2600), (335, 3, 2600), (1004, 8, 700), (400, 9, 300)])
def test_cloud_batched_launch_matches_single_launches(gpu_ctx, S, n_batches, G):
    """idash_b200_cloud_eval_device_batched: several input sets through the same model in ONE launch of the ring kernel
    (virtual tiles; batch boundaries inside a chunk) == one cloud_eval_device call per set == the oracle on a sample. More
    than 8 sets (or an ineligible model) falls back to one launch per set."""
    import torch
    geo, model, cts, var = make_case(S, T=500, G=G, n=5, seed=S + n_batches)
    m = api.Model(gpu_ctx, S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef)
    gpu_ctx.set_kernel(api.KERNEL_AUTO)
    g = torch.Generator(device="cuda").manual_seed(n_batches)
    ins = [torch.from_numpy(cts.view(np.int32)).cuda()] + \
          [torch.randint(-2 ** 31, 2 ** 31, cts.shape, dtype=torch.int32, device="cuda", generator=g) for _ in range(n_batches - 1)]
    outs_b = [torch.zeros((model.n_out, 2048), dtype=torch.int32, device="cuda") for _ in range(n_batches)]
    outs_s = [torch.zeros((model.n_out, 2048), dtype=torch.int32, device="cuda") for _ in range(n_batches)]
    l0 = gpu_ctx.kernel_launches()
    api.cloud_compute_score_device_batched(gpu_ctx, m, ins, outs_b)
    l1 = gpu_ctx.kernel_launches()
    for b in range(n_batches):
        api.cloud_compute_score_device(gpu_ctx, m, ins[b], outs_s[b])
    torch.cuda.synchronize()
    gpu_ctx.check_device_status()
    if m.info["ring_ok"] and m.info["n_tiles"] * n_batches >= 1 and n_batches <= 8:
        assert gpu_ctx.last_kernel() == _lib.KERNEL_TENSOR_RING
    for b in range(n_batches):
        assert torch.equal(outs_b[b], outs_s[b]), b
    ref_out, _ = _oracle(S, geo, model, cts, var)
    assert np.array_equal(outs_b[0].cpu().numpy().view(np.uint32), ref_out)
    assert l1 > l0
    m.free()


@pytest.mark.parametrize("sizes,n,G", [((335, 334, 335), 5, 2600), ((335, 334, 335), 20, 900), ((1004, 1000, 980, 1004), 5, 1500)])
def test_cloud_population_models_in_one_launch(gpu_ctx, sizes, n, G):
    """BASELINE configs[3]: the population-stratified model sets -- one model and one set of ciphertexts per population, sample counts
    that differ (335 / 334 / 335 columns of the same tag file, so NUM_REGIONS and REGION_SIZE agree) -- evaluated by ONE launch of
    the ring kernel (idash_b200_cloud_eval_device_multi_model) == one launch per population == the oracle, for every population:
    each model's own coefficients, Constants and NUM_SAMPLES (the Constant is added to b[0..S) of ITS population only)."""
    import torch
    cases = [make_case(S, T=500, G=G, n=n, seed=31 * k + S) for k, S in enumerate(sizes)]
    assert len({(c[0].NR, c[0].RS) for c in cases}) == 1
    gpu_ctx.set_kernel(api.KERNEL_AUTO)
    models = [api.Model(gpu_ctx, S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef) for S, (geo, model, _, _) in zip(sizes, cases)]
    ins = [torch.from_numpy(c[2].view(np.int32)).cuda() for c in cases]
    n_out = cases[0][1].n_out
    outs_m = [torch.zeros((n_out, 2048), dtype=torch.int32, device="cuda") for _ in cases]
    outs_s = [torch.zeros((n_out, 2048), dtype=torch.int32, device="cuda") for _ in cases]
    l0 = gpu_ctx.kernel_launches()
    api.cloud_compute_score_device_multi_model(gpu_ctx, models, ins, outs_m)
    torch.cuda.synchronize()
    main_launches = gpu_ctx.kernel_launches() - l0
    assert gpu_ctx.last_kernel() == _lib.KERNEL_TENSOR_RING
    assert main_launches == 1 + len(cases)               # one ring launch + one (tiny) per-row finalize launch per population
    for b, (S, (geo, model, cts, var)) in enumerate(zip(sizes, cases)):
        api.cloud_compute_score_device(gpu_ctx, models[b], ins[b], outs_s[b])
        torch.cuda.synchronize()
        assert torch.equal(outs_m[b], outs_s[b]), b
        ref_out, _ = _oracle(S, geo, model, cts, var)
        assert np.array_equal(outs_m[b].cpu().numpy().view(np.uint32), ref_out), b
    gpu_ctx.check_device_status()
    for m in models:
        m.free()


def test_cloud_permuted_input_slots_and_scattered_outputs(kctx):
    S = 400
    geo, model, cts, var = make_case(S, T=50, G=80, n=5, seed=21)
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(cts)).astype(np.uint32)
    var = var * rng.integers(1, 5, size=len(cts))
    cts_p, var_p = np.ascontiguousarray(cts[perm]), np.ascontiguousarray(var[perm])
    slot_of_row = rng.permutation(model.n_out).astype(np.uint32)
    m = _model(kctx, S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef)
    out, idx, ovar = api.cloud_compute_score(kctx, m, cts_p, in_index=perm, in_var=var_p, slot_of_row=slot_of_row)
    ref_out, ref_var = _oracle(S, geo, model, cts_p, var_p, perm)
    assert np.array_equal(out[slot_of_row], ref_out)
    assert np.array_equal(ovar[slot_of_row], ref_var)
    assert np.array_equal(idx[slot_of_row], model.out_bidx)
    _used(kctx)
    m.free()


def test_cloud_int32_range_coefficients_wrap(gpu_ctx):
    """Coefficients outside int16 are not eligible for the tensor-core kernel: AUTO must pick the IMAD one,
    forcing TENSOR must fail loudly."""
    S = 335
    geo, model, cts, var = make_case(S, T=30, G=40, n=5, seed=31)
    rng = np.random.default_rng(4)
    coef = rng.integers(-2 ** 31, 2 ** 31, size=model.nnz).astype(np.int32)
    m = api.Model(gpu_ctx, S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, coef)
    out, _, ovar = api.cloud_compute_score(gpu_ctx, m, cts, in_var=var)
    ref_out, ref_var = po.cloud_port(S, geo.NR, geo.RS, np.arange(len(cts), dtype=np.uint32), cts, var, model.row_ptr,
                                     model.col, coef)
    assert np.array_equal(out, ref_out) and np.array_equal(ovar, ref_var)
    assert gpu_ctx.last_kernel() == _lib.KERNEL_IMAD and m.info["n_tiles"] == 0
    gpu_ctx.set_kernel(_lib.KERNEL_TENSOR)
    with pytest.raises(api.IdashB200Error):
        api.cloud_compute_score(gpu_ctx, m, cts, in_var=var)
    gpu_ctx.set_kernel(_lib.KERNEL_AUTO)
    m.free()


def test_cloud_outlier_window_does_not_change_the_kernel(gpu_ctx):
    """Per-tile eligibility: one 400-feature-wide window (a tile band holds 224) and one coefficient outside int16 in an otherwise
    neighbors = 5 model. The two rows go to the IMAD kernel (overflow groups), the other 3 790-odd rows stay on the persistent
    ring kernel, and every output word still equals the oracle -- through the device entry point, the host entry point with
    scattered output slots, and a batched call (which falls back to one launch per set when there are overflow rows)."""
    import torch
    S = 1004
    geo, model, cts, var = make_case(S, T=900, G=4000, n=5, seed=23)
    rp, col, coef = model.row_ptr.astype(np.int64), model.col.copy(), model.coef.copy()
    wide, big = 5000, 9001
    a, b = int(rp[wide]), int(rp[wide + 1])
    real = np.nonzero(col[a:b] != 0xFFFFFFFF)[0]
    col[a + real[-1]] = col[a + real[0]] + 399
    a, b = int(rp[big]), int(rp[big + 1])
    real = np.nonzero(col[a:b] != 0xFFFFFFFF)[0]
    coef[a + real[0]] = -70000
    gpu_ctx.set_kernel(_lib.KERNEL_AUTO)
    m = api.Model(gpu_ctx, S, 1, 1024, model.out_bidx, model.row_ptr, col, coef)
    assert m.info["n_overflow_rows"] == 2 and m.info["ring_ok"] == 1 and m.info["n_tiles"] == (model.n_out + 63) // 64
    ref_out, ref_var = po.cloud_port(S, 1, 1024, np.arange(len(cts), dtype=np.uint32), cts, var, model.row_ptr, col, coef)
    out, idx, ovar = api.cloud_compute_score(gpu_ctx, m, cts, in_var=var)
    assert gpu_ctx.last_kernel() == _lib.KERNEL_TENSOR_RING
    assert np.array_equal(out, ref_out) and np.array_equal(ovar, ref_var) and np.array_equal(idx, model.out_bidx)
    slot_of_row = np.random.default_rng(1).permutation(model.n_out).astype(np.uint32)
    out2, _, ovar2 = api.cloud_compute_score(gpu_ctx, m, cts, in_var=var, slot_of_row=slot_of_row)
    assert np.array_equal(out2[slot_of_row], ref_out) and np.array_equal(ovar2[slot_of_row], ref_var)
    x = torch.from_numpy(cts.view(np.int32)).cuda()
    outs = [torch.zeros((model.n_out, 2048), dtype=torch.int32, device="cuda") for _ in range(2)]
    api.cloud_compute_score_device_batched(gpu_ctx, m, [x, x], outs)
    torch.cuda.synchronize()
    gpu_ctx.check_device_status()
    assert np.array_equal(outs[1].cpu().numpy().view(np.uint32), ref_out) and torch.equal(outs[0], outs[1])
    # forcing the IMAD kernel for the whole model builds the full group layout on demand
    gpu_ctx.set_kernel(_lib.KERNEL_IMAD)
    out3, _, _ = api.cloud_compute_score(gpu_ctx, m, cts, in_var=var)
    assert gpu_ctx.last_kernel() == _lib.KERNEL_IMAD and np.array_equal(out3, ref_out)
    gpu_ctx.set_kernel(_lib.KERNEL_AUTO)
    m.free()


def test_cloud_auto_picks_tensor_kernel_for_idash_models(gpu_ctx):
    geo, model, cts, var = make_case(1004, T=60, G=101, n=5, seed=2)
    m = api.Model(gpu_ctx, 1004, 1, 1024, model.out_bidx, model.row_ptr, model.col, model.coef)
    out, _, _ = api.cloud_compute_score(gpu_ctx, m, cts, in_var=var)
    assert gpu_ctx.last_kernel() == _lib.KERNEL_TENSOR_RING and m.info["ring_ok"] == 1
    assert np.array_equal(out, _oracle(1004, geo, model, cts, var)[0])
    m.free()


def test_cloud_tensor_kernel_extreme_int16_coefficients_and_wide_band(kctx):
    """Extreme eligible coefficients (both limbs at +-128/127), all-ones words (max limb products), a 40-neighbour band."""
    S = 1004
    geo, model, cts, var = make_case(S, T=120, G=150, n=40, seed=9, coef_range=200, bias_range=500)
    rng = np.random.default_rng(5)
    coef = rng.choice(np.array([-32896, -32768, -32767, -129, -128, -1, 1, 127, 128, 255, 256, 32639, 32512, -256, -255], np.int32), size=model.nnz)
    cts = cts.copy()
    cts[::3] = 0xFFFFFFFF
    m = _model(kctx, S, 1, 1024, model.out_bidx, model.row_ptr, model.col, coef)
    out, _, ovar = api.cloud_compute_score(kctx, m, cts, in_var=var)
    ref_out, ref_var = po.cloud_port(S, 1, 1024, np.arange(len(cts), dtype=np.uint32), cts, var, model.row_ptr, model.col, coef)
    assert np.array_equal(out, ref_out) and np.array_equal(ovar, ref_var)
    _used(kctx)
    m.free()


def test_cloud_sparse_unbanded_shuffled_rows(kctx):
    rng = np.random.default_rng(8)
    S, NR, RS = 16, 64, 16
    n_ct = 4                        # 256 input features: the widest band a tensor-core tile takes
    cts = synth.random_ciphertexts(n_ct, 8)
    var = np.full(n_ct, ALPHA2)
    out_bidx = rng.permutation(np.array([0, 1, 2, 4, 5, 6, 7, 8, 40, 41, 300, 302, 303], np.uint32))
    row_ptr, col, coef = [0], [], []
    for r in range(len(out_bidx)):
        feats = rng.choice(NR * n_ct, size=int(rng.integers(0, 40)), replace=False)
        if r % 4:
            col.append(0xFFFFFFFF); coef.append(int(rng.integers(-9000, 9000)))
        for f in feats:
            col.append(int(f)); coef.append(int(rng.integers(-300, 300)))
        row_ptr.append(len(col))
    row_ptr, col, coef = np.array(row_ptr, np.uint64), np.array(col, np.uint32), np.array(coef, np.int32)
    m = _model(kctx, S, NR, RS, out_bidx, row_ptr, col, coef)
    out, idx, ovar = api.cloud_compute_score(kctx, m, cts, in_var=var)
    ref_out, ref_var = po.cloud_port(S, NR, RS, np.arange(n_ct, dtype=np.uint32), cts, var, row_ptr, col, coef)
    assert np.array_equal(out, ref_out) and np.array_equal(ovar, ref_var) and np.array_equal(idx, out_bidx)
    _used(kctx)
    m.free()


def test_cloud_missing_input_is_an_error_not_garbage(kctx):
    geo, model, cts, var = make_case(1004, T=20, G=30, n=5, seed=41)
    m = _model(kctx, 1004, 1, 1024, model.out_bidx, model.row_ptr, model.col, model.coef)
    with pytest.raises(api.IdashB200Error) as e:
        api.cloud_compute_score(kctx, m, cts[:-5], in_var=var[:-5])
    assert e.value.code == _lib.ERR_MISSING_INPUT
    # and the context keeps working afterwards
    out, _, _ = api.cloud_compute_score(kctx, m, cts, in_var=var)
    assert np.array_equal(out, _oracle(1004, geo, model, cts, var)[0])
    _used(kctx)
    m.free()


def test_cloud_empty_model_and_row_count_mismatch(gpu_ctx):
    m = api.Model(gpu_ctx, 1004, 1, 1024, np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.uint32),
                  np.zeros(0, np.int32))
    out, idx, var = api.cloud_compute_score(gpu_ctx, m, np.zeros((0, 2048), np.uint32))
    assert out.shape == (0, 2048)
    m.free()
    geo, model, cts, var = make_case(1004, T=10, G=6, n=3, seed=5)
    m = api.Model(gpu_ctx, 1004, 1, 1024, model.out_bidx, model.row_ptr, model.col, model.coef)
    with pytest.raises(api.IdashB200Error):
        api.cloud_compute_score(gpu_ctx, m, cts, out_ct=np.empty((model.n_out, 2048), np.uint32)[:-1].copy())
    m.free()


def test_cloud_default_variance_is_alpha_squared(kctx):
    geo, model, cts, var = make_case(1004, T=10, G=6, n=3, seed=6)
    m = _model(kctx, 1004, 1, 1024, model.out_bidx, model.row_ptr, model.col, model.coef)
    _, _, v0 = api.cloud_compute_score(kctx, m, cts)
    _, _, v1 = api.cloud_compute_score(kctx, m, cts, in_var=np.full(len(cts), ALPHA2))
    assert np.array_equal(v0, v1) and (v0 > 0).all()
    _used(kctx)
    m.free()


def test_cloud_device_path_linearity_at_scale(kctx):
    """Size-independent property at a larger size: eval(x + y) == eval(x) + eval(y) - eval(0) mod 2^32 (the
    map is affine: bias + linear), through the device-resident entry point on torch tensors."""
    import torch
    S, T, G, n = 1004, 400, 2000, 5
    geo = synth.Geometry(S, T, G)
    tag, tgt = synth.make_positions(T, G, 77)
    model = synth.make_model(tag, tgt, n, 77)
    m = _model(kctx, S, 1, 1024, model.out_bidx, model.row_ptr, model.col, model.coef)
    g = torch.Generator(device="cuda").manual_seed(1)
    n_in = geo.n_in_ct_used
    x = torch.randint(-2 ** 31, 2 ** 31, (n_in, 2048), dtype=torch.int32, device="cuda", generator=g)
    y = torch.randint(-2 ** 31, 2 ** 31, (n_in, 2048), dtype=torch.int32, device="cuda", generator=g)
    z = torch.zeros_like(x)
    outs = []
    for inp in (x, y, x + y, z):
        o = torch.empty((model.n_out, 2048), dtype=torch.int32, device="cuda")
        api.cloud_compute_score_device(kctx, m, inp, o)
        outs.append(o)
    torch.cuda.synchronize()
    kctx.check_device_status()
    assert torch.equal(outs[2], outs[0] + outs[1] - outs[3])
    # spot-check 64 rows against the oracle
    rows = np.random.default_rng(0).choice(model.n_out, 64, replace=False)
    sub = model  # full CSR; compute oracle on the selected rows only
    rp = [0]; col = []; coef = []
    for r in rows:
        a, b = int(sub.row_ptr[r]), int(sub.row_ptr[r + 1])
        col.extend(sub.col[a:b]); coef.extend(sub.coef[a:b]); rp.append(len(col))
    xin = x.cpu().numpy().view(np.uint32)
    ref_out, _ = po.cloud_port(S, 1, 1024, np.arange(n_in, dtype=np.uint32), xin, np.full(n_in, ALPHA2),
                               np.array(rp, np.uint64), np.array(col, np.uint32), np.array(coef, np.int32))
    assert np.array_equal(outs[0].cpu().numpy().view(np.uint32)[rows], ref_out)
    _used(kctx)
    m.free()


@pytest.mark.parametrize("n", [5, 20])
def test_cloud_host_path_pipelined_pieces(gpu_ctx, n, monkeypatch):
    """Large-enough models take the pipelined host path (pieces of the target range; copy-in, kernels and copy-out on
    three streams): PACKED with variances and RECORDS outputs must equal the oracle and the unpipelined path, and a
    missing input ciphertext must still be reported."""
    S, T, G = 1004, 800, 3200                      # 9600 rows = 150 tiles -> pipelined (>= 128 tiles)
    geo, model, cts, var = make_case(S, T=T, G=G, n=n, seed=77 + n)
    var = var * (1.0 + np.arange(len(var)) % 3)    # not all equal
    m = api.Model(gpu_ctx, S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef)
    assert m.info["ring_ok"] and m.info["n_tiles"] >= 128
    ref_out, ref_var = _oracle(S, geo, model, cts, var)
    out, idx, ovar = api.cloud_compute_score(gpu_ctx, m, cts, in_var=var)
    assert gpu_ctx.last_kernel() == _lib.KERNEL_TENSOR_RING
    assert np.array_equal(out, ref_out) and np.array_equal(ovar, ref_var) and np.array_equal(idx, model.out_bidx)
    # RECORDS in (hash-like permuted order: whole upload, pipelined download) and RECORDS out
    perm = np.random.default_rng(3).permutation(len(cts))
    img = formats.build_ct_image(perm.astype(np.uint32), cts[perm], var[perm])
    pred = api.cloud_compute_score_records(gpu_ctx, m, img)
    pi, pw, pv = formats.image_views(pred)
    assert np.array_equal(pi, model.out_bidx) and np.array_equal(pw, ref_out) and np.array_equal(pv, ref_var)
    # same call with the pipeline switched off
    monkeypatch.setenv("IDASH_B200_NO_PIPELINE", "1")
    out2, _, ovar2 = api.cloud_compute_score(gpu_ctx, m, cts, in_var=var)
    monkeypatch.delenv("IDASH_B200_NO_PIPELINE")
    assert np.array_equal(out2, out) and np.array_equal(ovar2, ovar)
    # too few input ciphertexts: the model's last tiles reference slots that were not supplied
    with pytest.raises(api.IdashB200Error) as e:
        api.cloud_compute_score(gpu_ctx, m, cts[: len(cts) - 40], in_var=var[: len(cts) - 40])
    assert e.value.code == _lib.ERR_MISSING_INPUT
    out3, _, _ = api.cloud_compute_score(gpu_ctx, m, cts, in_var=var)     # the context stays usable
    assert np.array_equal(out3, ref_out)
    m.free()


@pytest.mark.parametrize("S", [400, 1004])
def test_ring_slot_reuse_with_slow_producers(S):
    """Two MMA warps wait for ring slots by parity; a slot that is reused within two tiles (the band moves by several blocks per tile:
    500 tag x 300 target SNPs) must not be taken for staged while its previous use is still pending. Found under compute-sanitizer,
    whose instrumentation slows the producers; the profiling build reproduces that timing (knock-out 64: producers sleep before every
    block). Runs tools/race_probe2.py in its own process (the profiling library is chosen at import): every launch == the oracle."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    if not (root / "idash2019_2_b200" / "lib" / "libidash_b200_prof.so").exists():
        pytest.skip("profiling build of the library not present")
    env = dict(os.environ, IDASH_B200_USE_PROFILE_LIB="1", IDASH_B200_KNOCKOUT="64")
    r = subprocess.run([sys.executable, str(root / "tools" / "race_probe2.py"), str(S), "300", "3"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("rep")]
    assert len(lines) == 3 and all(ln.endswith("bad rows []") for ln in lines), r.stdout[-2000:]
