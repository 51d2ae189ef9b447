"""Test helpers: seeded cases, golden loading, and a numpy INTERPRETER of the device layout.

The interpreter executes idash_b200_group / idash_b200_entry arrays exactly as cloud_eval_kernel is
specified to (include/idash_b200_layout.h), so the host-side model compiler can be checked against
the oracle on a CPU-only box.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from idash2019_2_b200 import formats, synth
from idash2019_2_b200._lib import NO_ROW, ONE_IN_T32

GOLDEN = Path(__file__).resolve().parent / "golden"
GOLDEN_CASES = ["s1004_nr1", "s400_nr2", "s335_nr3", "s16_nr64"]
N = 1024
ALPHA2 = 2.0 ** -50


def rot(poly: np.ndarray, shift: int) -> np.ndarray:
    """X^(-shift) * poly mod X^1024 + 1 on uint32 words."""
    if shift == 0:
        return poly
    out = np.empty_like(poly)
    out[: N - shift] = poly[shift:]
    out[N - shift:] = (0 - poly[:shift].astype(np.int64)).astype(np.uint32)
    return out


def interpret_layout(layout, S, RS, in_ct, in_var, slot_of_ct=None, overflow=False):
    """-> (out_ct [n_rows, 2048] in caller row order, out_var [n_rows]). overflow=True: the groups of the rows no tile holds
    only; returns (out_ct, seen [n_rows] bool) instead."""
    n_rows = layout.info["n_rows"]
    out = np.zeros((n_rows, 2048), np.uint32)
    seen = np.zeros(n_rows, bool)
    entries = layout.overflow_entries if overflow else layout.entries
    for G in (layout.overflow_groups if overflow else layout.groups):
        acc = np.zeros((6, 2048), np.uint32)
        e = int(G["entry_begin"])
        for cls, cnt in ((0, int(G["n_a"])), (1, int(G["n_ab"])), (2, int(G["n_b"]))):
            rows = {0: (0, 1, 2), 1: (0, 1, 2, 3, 4, 5), 2: (3, 4, 5)}[cls]
            for _ in range(cnt):
                E = entries[e]
                e += 1
                ct = int(E["ct"])
                slot = ct if slot_of_ct is None else slot_of_ct[ct]
                x = np.concatenate([rot(in_ct[slot, :N], int(E["shift"])), rot(in_ct[slot, N:], int(E["shift"]))])
                for r in range(6):
                    c = np.uint32(int(E["coef"][r]) & 0xFFFFFFFF)
                    if r in rows:
                        acc[r] += c * x
                    else:
                        assert E["coef"][r] == 0
        for r in range(6):
            row = int(G["row"][r])
            if row == NO_ROW:
                continue
            v = acc[r].copy()
            sb = np.uint32((int(G["bias"][r]) * ONE_IN_T32) & 0xFFFFFFFF)
            v[N:N + S] += sb
            v[N + RS:] = 0
            assert not seen[row]
            seen[row] = True
            out[row] = v
    if overflow:
        return out, seen
    assert seen.all()
    var = np.zeros(n_rows)
    for r in range(n_rows):
        for e in range(int(layout.var_ptr[r]), int(layout.var_ptr[r + 1])):
            ct = int(layout.var_ct[e])
            slot = ct if slot_of_ct is None else slot_of_ct[ct]
            var[r] += layout.var_w[e] * in_var[slot]
    return out, var


def interpret_tiles(layout, S, NR, RS, in_ct, slot_of_ct=None, partial=False):
    """Executes the band-tile layout the way cloud_tc_kernel is specified to (include/idash_b200_layout.h):
    u8 limb planes of the rotated inputs times the (c_lo u8, c_hi s8) coefficient images, four int32
    accumulators P_w recombined with shifts. -> out_ct [n_rows, 2048] in caller row order."""
    from idash2019_2_b200._lib import TILE_ROWS
    TN = TILE_ROWS
    n_rows = layout.info["n_rows"]
    out = np.zeros((n_rows, 2048), np.uint32)
    seen = np.zeros(n_rows, bool)
    n_slots = len(in_ct)
    for t, T in enumerate(layout.tiles):
        K, f_base, b_off = int(T["K"]), int(T["f_base"]), int(T["b_off"])
        img = layout.tile_coef[b_off:b_off + 2 * K * TN].view(np.int8).reshape(K // 32, 2, 2, TN, 16)   # [chunk][k16][limb][row][k%16]
        c_lo = img[:, :, 0].transpose(2, 0, 1, 3).reshape(TN, K).astype(np.int64)           # [row][k], balanced signed limbs
        c_hi = img[:, :, 1].transpose(2, 0, 1, 3).reshape(TN, K).astype(np.int64)
        used = layout.tile_used[int(T["used_off"]):int(T["used_off"]) + K // 32]
        X = np.zeros((K, 2048), np.uint32)
        for k in range(K):
            f = f_base + k
            ct, shift = f // NR, (f % NR) * RS
            slot = (ct if ct < n_slots else None) if slot_of_ct is None else slot_of_ct.get(ct)
            if slot is None:
                assert not (int(used[k // 32]) >> (k % 32)) & 1, "tile uses a missing ciphertext"
                continue
            X[k] = np.concatenate([rot(in_ct[slot, :N], shift), rot(in_ct[slot, N:], shift)])
        limbs = [((X >> (8 * j)) & 0xFF).astype(np.int64) for j in range(4)]               # [j][k][word]
        acc = np.zeros((TN, 2048), np.int64)
        for w in range(4):
            P = c_lo @ limbs[w]
            if w:
                P = P + c_hi @ limbs[w - 1]
            assert np.abs(P).max() < 2 ** 31                                                # int32 accumulators never overflow
            acc += P << (8 * w)
        acc = (acc & 0xFFFFFFFF).astype(np.uint32)
        for n in range(TN):
            row = int(layout.tile_rows[t * TN + n])
            if row == NO_ROW:
                assert partial or n >= int(T["n_valid"])
                assert not img[:, :, :, n, :].any(), "a hole of a tile carries coefficients"
                continue
            v = acc[n].copy()
            v[N:N + S] += np.uint32((int(layout.tile_bias[t * TN + n]) * ONE_IN_T32) & 0xFFFFFFFF)
            v[N + RS:] = 0
            assert not seen[row]
            seen[row] = True
            out[row] = v
    if partial:
        return out, seen
    assert seen.all()
    return out


def make_case(S, T, G, n, seed, coef_range=200, bias_range=500):
    """Seeded synthetic problem: geometry, CSR model, random ciphertext words, 2^-50 variances."""
    geo = synth.Geometry(S, T, G)
    tag, tgt = synth.make_positions(T, G, seed)
    model = synth.make_model(tag, tgt, n, seed, coef_range=coef_range, bias_range=bias_range)
    n_in = geo.n_in_ct_used
    cts = synth.random_ciphertexts(n_in, seed)
    var = np.full(n_in, ALPHA2)
    return geo, model, cts, var


def load_golden(name):
    d = GOLDEN / name
    params, key, props = formats.read_key(d / "keys.bin")
    enc = formats.read_ct_image(d / "encrypted_data.bin")
    pred = formats.read_ct_image(d / "encrypted_prediction.bin")
    ref = np.load(d / "ref.npz")
    return d, params, key, enc, pred, ref
