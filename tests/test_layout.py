"""Host logic (no GPU): the model -> block-banded layout compiler, checked by interpreting the layout
with numpy (tests/helpers.py) against the oracle; the C ABI library loads and exports every symbol
the headers declare; argument validation that does not need a device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from idash2019_2_b200 import _lib, api, formats, synth
from oracle import pyoracle as po

from helpers import ALPHA2, GOLDEN_CASES, interpret_layout, interpret_tiles, load_golden, make_case

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(built_lib):
    declared = set()
    for h in (ROOT / "include").glob("*.h"):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        declared |= set(re.findall(r"\b(idash_b200_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations found"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(built_lib, name) is not None


@pytest.mark.parametrize("S,n", [(1004, 5), (1024, 1), (512, 5), (400, 6), (335, 5), (64, 3), (16, 5)])
def test_layout_interpreter_matches_oracle(built_lib, S, n):
    geo, model, cts, var = make_case(S, T=25, G=21, n=n, seed=10 * S + n, coef_range=8191, bias_range=8191)
    lay = api.compile_layout(S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef)
    out, ovar = interpret_layout(lay, S, geo.RS, cts, var)
    ref_out, ref_var = po.cloud_port(S, geo.NR, geo.RS, np.arange(len(cts), dtype=np.uint32), cts, var, model.row_ptr,
                                     model.col, model.coef)
    assert np.array_equal(out, ref_out)
    assert np.array_equal(ovar, ref_var)
    info = lay.info
    assert info["n_rows"] == model.n_out and info["nnz"] == model.nnz
    assert info["n_groups"] == (21 + 1) // 2
    assert info["shifts_aligned"] == (1 if geo.RS % 4 == 0 or geo.NR == 1 else 0)
    used = model.col[model.col != 0xFFFFFFFF] // geo.NR
    assert info["ct_min"] == used.min() and info["ct_max"] == used.max()
    # each group lists a (ct, shift) pair at most once and splits it A | AB | B
    for G in lay.groups:
        e0, cnt = int(G["entry_begin"]), int(G["n_a"] + G["n_ab"] + G["n_b"])
        keys = {(int(E["ct"]), int(E["shift"])) for E in lay.entries[e0:e0 + cnt]}
        assert len(keys) == cnt
        assert (lay.entries[e0:e0 + int(G["n_a"])]["coef"][:, 3:] == 0).all()
        assert (lay.entries[e0 + int(G["n_a"] + G["n_ab"]):e0 + cnt]["coef"][:, :3] == 0).all()


@pytest.mark.parametrize("S,n,cr", [(1004, 5, 200), (1024, 1, 32639), (512, 5, 8191), (400, 20, 200), (335, 5, 8191), (64, 3, 200),
                                    (16, 5, 8191)])
def test_tile_interpreter_matches_oracle(built_lib, S, n, cr):
    """The tensor-core operand layout (band tiles, limb-split coefficients) reproduces the oracle."""
    geo, model, cts, var = make_case(S, T=45, G=70, n=n, seed=7 * S + n, coef_range=cr, bias_range=cr)
    lay = api.compile_layout(S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef)
    assert lay.info["n_tiles"] == (model.n_out + 63) // 64 and lay.info["tile_kmax"] % 32 == 0
    out = interpret_tiles(lay, S, geo.NR, geo.RS, cts)
    ref_out, _ = po.cloud_port(S, geo.NR, geo.RS, np.arange(len(cts), dtype=np.uint32), cts, var, model.row_ptr,
                               model.col, model.coef)
    assert np.array_equal(out, ref_out)
    for T in lay.tiles:
        assert T["K"] % 32 == 0 and 32 <= T["K"] <= _lib.TILE_KMAX and T["b_off"] % 16 == 0


def _with_outliers(model, wide_row, big_row, wide_span=400):
    """The CSR of `model` with row `wide_row` stretched to a window of `wide_span` features and one coefficient of row `big_row`
    outside the limb range of the tensor-core kernels."""
    rp, col, coef = model.row_ptr.astype(np.int64), model.col.copy(), model.coef.copy()
    a, b = int(rp[wide_row]), int(rp[wide_row + 1])
    real = np.nonzero(col[a:b] != 0xFFFFFFFF)[0]
    col[a + real[-1]] = col[a + real[0]] + wide_span - 1          # entries stay distinct: the last one moves far to the right
    coef[a + real[-1]] = 77
    a, b = int(rp[big_row]), int(rp[big_row + 1])
    real = np.nonzero(col[a:b] != 0xFFFFFFFF)[0]
    coef[a + real[1]] = 40000
    return col, coef


def test_outlier_rows_are_evicted_and_the_rest_keeps_its_tiles(built_lib):
    """Per-tile eligibility: one window of 400 features (a tile band holds at most 224) and one coefficient outside int16 send TWO
    rows to the overflow groups (IMAD kernel); every other row stays on its band tile and the model stays eligible for the
    persistent ring kernel. Tiles + overflow groups, interpreted, cover every row exactly once and equal the oracle."""
    S = 1004
    geo, model, cts, var = make_case(S, T=300, G=400, n=5, seed=19)
    col, coef = _with_outliers(model, wide_row=301, big_row=700)
    cts = synth.random_ciphertexts(int(col[col != 0xFFFFFFFF].max()) + 1, 19)
    var = np.full(len(cts), ALPHA2)
    lay = api.compile_layout(S, 1, 1024, model.out_bidx, model.row_ptr, col, coef, flags=_lib.COMPILE_DEFAULT)
    info = lay.info
    assert info["n_overflow_rows"] == 2 and info["ring_ok"] == 1 and info["n_tiles"] == (model.n_out + 63) // 64
    assert info["tile_kmax"] <= _lib.RING_KMAX and info["groups_all"] == 0 and len(lay.groups) == 0
    assert sorted(int(r) for G in lay.overflow_groups for r in G["row"] if r != _lib.NO_ROW) == [301, 700]
    ref_out, ref_var = po.cloud_port(S, 1, 1024, np.arange(len(cts), dtype=np.uint32), cts, var, model.row_ptr, col, coef)
    t_out, t_seen = interpret_tiles(lay, S, 1, 1024, cts, partial=True)
    o_out, o_seen = interpret_layout(lay, S, 1024, cts, var, overflow=True)
    assert not (t_seen & o_seen).any() and (t_seen | o_seen).all() and o_seen.sum() == 2
    assert np.array_equal(np.where(t_seen[:, None], t_out, o_out), ref_out)
    # tiles with a hole are not "fast" (consecutive caller rows)
    holes = {301 // 64, 700 // 64}
    for t, T in enumerate(lay.tiles):
        assert (int(T["flags"]) & 1) == (0 if t in holes or t == len(lay.tiles) - 1 and model.n_out % 64 else 1)
    # the full IMAD layout on request still covers everything
    full = api.compile_layout(S, 1, 1024, model.out_bidx, model.row_ptr, col, coef)
    out, ovar = interpret_layout(full, S, 1024, cts, var)
    assert np.array_equal(out, ref_out) and np.array_equal(ovar, ref_var) and full.info["groups_all"] == 1


def test_tiles_absent_only_when_no_row_is_eligible(built_lib):
    geo, model, cts, var = make_case(1004, T=20, G=30, n=5, seed=3)
    coef = model.coef.copy()
    coef[5] = 40000                                                   # outside int16: that row only
    lay = api.compile_layout(1004, 1, 1024, model.out_bidx, model.row_ptr, model.col, coef)
    assert lay.info["n_tiles"] == 2 and lay.info["n_overflow_rows"] == 1
    big = np.where(model.col == 0xFFFFFFFF, model.coef, 40000).astype(np.int32)          # every row
    lay = api.compile_layout(1004, 1, 1024, model.out_bidx, model.row_ptr, model.col, big)
    assert lay.info["n_tiles"] == 0 and len(lay.tiles) == 0 and lay.info["n_groups"] > 0 and lay.info["n_overflow_rows"] == model.n_out
    # a window wider than a tile band
    ob = np.array([0], np.uint32)
    lay = api.compile_layout(1004, 1, 1024, ob, np.array([0, 2], np.uint64), np.array([0, 300], np.uint32),
                             np.array([1, 1], np.int32))
    assert lay.info["n_tiles"] == 0 and lay.info["n_overflow_rows"] == 1
    lay = api.compile_layout(1004, 1, 1024, ob, np.array([0, 2], np.uint64), np.array([0, 223], np.uint32),
                             np.array([1, -1], np.int32))
    assert lay.info["n_tiles"] == 1 and lay.info["tile_kmax"] == 224 and lay.info["n_overflow_rows"] == 0


def test_layout_cache_roundtrip_and_thread_count_independence(built_lib, tmp_path, monkeypatch):
    """models.bin: save + load reproduces every array; a wrong key, a truncated file or a missing file is refused. The compiler's
    parallel passes give the same layout whatever the thread count."""
    S = 335
    geo, model, cts, var = make_case(S, T=200, G=700, n=20, seed=5, coef_range=8191)
    f = tmp_path / "models.bin"
    monkeypatch.setenv("IDASH_B200_THREADS", "1")
    a = api.compile_layout(S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef, save_to=f, key=0xABCDEF)
    monkeypatch.setenv("IDASH_B200_THREADS", "7")
    b = api.compile_layout(S, geo.NR, geo.RS, model.out_bidx, model.row_ptr, model.col, model.coef)
    c = api.load_layout(f, 0xABCDEF)
    for other in (b, c):
        assert other.info == a.info
        for name in ("groups", "entries", "var_ptr", "var_ct", "var_w", "out_bidx", "tiles", "tile_rows", "tile_bias", "tile_coef",
                     "tile_used", "overflow_groups", "overflow_entries", "feat_ptr", "feat_bidx", "feat_coef", "bias"):
            x, y = getattr(a, name), getattr(other, name)
            assert x.dtype == y.dtype and x.tobytes() == y.tobytes(), name
    with pytest.raises(api.IdashB200Error):
        api.load_layout(f, 0xABCDEE)
    with pytest.raises(api.IdashB200Error):
        api.load_layout(tmp_path / "nope.bin", 0xABCDEF)
    raw = f.read_bytes()
    (tmp_path / "short.bin").write_bytes(raw[:-100])
    with pytest.raises(api.IdashB200Error):
        api.load_layout(tmp_path / "short.bin", 0xABCDEF)


def test_variance_terms_keep_the_callers_entry_order(built_lib):
    """tLweAddMulTo adds coef^2 * var_in in the order the row's map is walked (eval/idash.cpp:800-817): the variance CSR keeps the
    caller's order, so non-uniform input variances give the same double as that walk."""
    ob = np.array([0], np.uint32)
    rp = np.array([0, 4], np.uint64)
    col = np.array([9, 0xFFFFFFFF, 3, 6], np.uint32)
    coef = np.array([3, 5, -7, 2], np.int32)
    lay = api.compile_layout(1004, 1, 1024, ob, rp, col, coef)
    assert lay.var_ct.tolist() == [9, 3, 6] and lay.var_w.tolist() == [9.0, 49.0, 4.0]
    assert lay.feat_bidx.tolist() == [3, 6, 9] and lay.feat_coef.tolist() == [-7, 2, 3] and lay.bias.tolist() == [5]


def test_layout_band_is_compact_at_idash_shape(built_lib):
    """~5 targets per tag, n=5: a 2-target group touches about n+1 tags, i.e. ~18 entries instead of 30."""
    geo, model, cts, var = make_case(1004, T=200, G=1000, n=5, seed=1)
    lay = api.compile_layout(1004, 1, 1024, model.out_bidx, model.row_ptr, model.col, model.coef)
    per_group = lay.info["n_entries"] / lay.info["n_groups"]
    assert 15 <= per_group <= 21
    assert lay.info["max_entries_per_group"] <= 30


def test_layout_arbitrary_sparse_rows_missing_variants_and_shuffled_rows(built_lib):
    """Not banded, rows in random caller order, some variants absent, zero coefficients, no Constant."""
    rng = np.random.default_rng(7)
    S, NR, RS = 400, 2, 512
    n_ct = 40
    cts = synth.random_ciphertexts(n_ct, 7)
    var = np.full(n_ct, ALPHA2)
    out_bidx = rng.permutation(np.array([0, 1, 2, 3, 5, 9, 10, 11, 30, 31, 32, 34, 100], np.uint32))
    row_ptr, col, coef = [0], [], []
    for r in range(len(out_bidx)):
        k = int(rng.integers(0, 12))
        feats = rng.choice(NR * n_ct, size=k, replace=False)
        if r % 3:
            col.append(0xFFFFFFFF)
            coef.append(int(rng.integers(-9000, 9000)))
        for f in feats:
            col.append(int(f))
            coef.append(int(rng.integers(-3, 4)))          # includes zeros
        row_ptr.append(len(col))
    row_ptr, col, coef = np.array(row_ptr, np.uint64), np.array(col, np.uint32), np.array(coef, np.int32)
    lay = api.compile_layout(S, NR, RS, out_bidx, row_ptr, col, coef)
    out, ovar = interpret_layout(lay, S, RS, cts, var)
    ref_out, ref_var = po.cloud_port(S, NR, RS, np.arange(n_ct, dtype=np.uint32), cts, var, row_ptr, col, coef)
    assert np.array_equal(out, ref_out) and np.array_equal(ovar, ref_var)
    assert np.array_equal(lay.out_bidx, out_bidx)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_layout_on_golden_model_files(built_lib, name):
    """.hr files -> read_model mirror -> layout -> interpreter == the reference cloud binary's output."""
    d, params, key, enc, pred, ref = load_golden(name)
    ob, rp, col, coef = formats.read_model(params, d / "model")
    # our loader == the reference's read_model (exported through the shim when the golden files were made)
    assert np.array_equal(ob, ref["model_out_bidx"]) and np.array_equal(rp, ref["model_row_ptr"])
    for r in range(len(ob)):
        a, b = int(rp[r]), int(rp[r + 1])
        mine = dict(zip(col[a:b].tolist(), coef[a:b].tolist()))
        theirs = dict(zip(ref["model_col"][a:b].tolist(), ref["model_coef"][a:b].tolist()))
        assert mine == theirs
    S, NR, RS = params.NUM_SAMPLES, params.NUM_REGIONS, params.REGION_SIZE
    lay = api.compile_layout(S, NR, RS, ob, rp, col, coef)
    in_idx, in_words, in_var = formats.image_views(enc)
    slot_of_ct = {int(i): s for s, i in enumerate(in_idx)}
    out, ovar = interpret_layout(lay, S, RS, in_words, in_var, slot_of_ct)
    p_idx, p_words, p_var = formats.image_views(pred)
    order = np.argsort(p_idx)
    assert np.array_equal(out, p_words[order]) and np.array_equal(ovar, p_var[order])
    assert np.array_equal(interpret_tiles(lay, S, NR, RS, in_words, slot_of_ct), p_words[order])


def test_layout_rejects_bad_models(built_lib):
    ob = np.array([0, 1], np.uint32)
    rp = np.array([0, 1, 2], np.uint64)
    col = np.array([3, 4], np.uint32)
    coef = np.array([1, 1], np.int32)
    with pytest.raises(api.IdashB200Error) as e:
        api.compile_layout(1004, 0, 1024, ob, rp, col, coef)
    assert e.value.code == _lib.ERR_INVALID
    with pytest.raises(api.IdashB200Error):
        api.compile_layout(1004, 2, 1024, ob, rp, col, coef)                      # NR * RS > 1024
    with pytest.raises(api.IdashB200Error):
        api.compile_layout(2000, 1, 1024, ob, rp, col, coef)                      # S > 1024
    with pytest.raises(api.IdashB200Error):
        api.compile_layout(1004, 1, 1024, np.array([5, 5], np.uint32), rp, col, coef)   # duplicate output
    with pytest.raises(api.IdashB200Error):
        api.compile_layout(1004, 1, 1024, ob, np.array([0, 2, 2], np.uint64), np.array([7, 7], np.uint32), coef)
    with pytest.raises(api.IdashB200Error):
        api.compile_layout(1004, 1, 1024, ob, np.array([0, 2, 2], np.uint64),
                           np.array([0xFFFFFFFF, 0xFFFFFFFF], np.uint32), coef)   # two Constants


def test_layout_empty_model(built_lib):
    lay = api.compile_layout(1004, 1, 1024, np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.uint32),
                             np.zeros(0, np.int32))
    assert lay.info["n_rows"] == 0 and lay.info["n_groups"] == 0 and len(lay.entries) == 0


def test_init_without_gpu_fails_loudly(built_lib):
    """No CPU fallback: without a CUDA device the context cannot be created."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.IdashB200Error) as e:
        api.Context(0)
    assert e.value.code == _lib.ERR_CUDA


def test_formats_roundtrip():
    d, params, key, enc, pred, ref = load_golden("s16_nr64")
    assert formats.serialize_params(params) == (d / "params.bin").read_bytes()
    idx, words, var = formats.image_views(enc)
    img = formats.build_ct_image(idx.copy(), words.copy(), var.copy())
    assert img.tobytes() == (d / "encrypted_data.bin").read_bytes()
    assert (img.ctypes.data + 16) % 16 == 0
    assert set(np.unique(key)) <= {0, 1} and len(key) == 1024
    assert list(ref["geometry"]) == [params.NUM_SAMPLES, params.NUM_INPUT_POSITIONS, params.NUM_OUTPUT_POSITIONS,
                                     params.NUM_INPUT_FEATURES, params.NUM_OUTPUT_FEATURES, params.NUM_REGIONS,
                                     params.REGION_SIZE]


def test_read_hr_sscanf_semantics(tmp_path):
    f = tmp_path / "1_0.hr"
    f.write_text("Constant -107.0\n16050075_1 23.9\n16050075_2   -0.7\n16050115_0 12\n16050075_1 5.0\n")
    assert formats.read_hr(f) == {"Constant": -107, "16050075_1": 5, "16050075_2": 0, "16050115_0": 12}


def test_ring_kernels_do_not_spill(built_lib):
    """The persistent ring kernel runs at the 128-register cap (512 threads, one CTA per SM). A spill there is not a small cost:
    local-memory traffic queues behind the epilogue's store backlog and the kernel gets 60 % slower (measured).
    ptxas' log of the in-tree build must therefore show no spill for any instantiation of cloud_ring_kernel."""
    log = (ROOT / "idash2019_2_b200" / "lib" / "ptxas.log")
    if not log.exists():
        pytest.skip("the library was not built here (no ptxas log)")
    text = log.read_text()
    entries = re.split(r"ptxas info\s+: Compiling entry function '", text)[1:]
    ring = [e for e in entries if e.startswith("_Z17cloud_ring_kernel")]
    assert ring, "no cloud_ring_kernel instantiation in the ptxas log"
    for e in ring:
        m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", e)
        assert m and m.group(1) == "0" and m.group(2) == "0", e.splitlines()[0] + ": " + (m.group(0) if m else "no spill line")
