"""Python host-side mirror of the reference interface for the hot path, on top of the C ABI.

    Context                       one per GPU (one process per GPU)
    Model                         Model of eval/idash.h:129-134 compiled + uploaded (idash_b200_model_upload)
    cloud_compute_score(...)      eval/idash.cpp:763-848, host numpy buffers  (idash_b200_cloud_eval_host)
    cloud_compute_score_device    same, torch CUDA tensors on the current stream (idash_b200_cloud_eval_device)
    decrypt_predictions(...)      eval/idash.cpp:681-761, host buffers        (idash_b200_decrypt_host)
    decrypt_predictions_device    same, torch CUDA tensors
    compile_layout(...)           the block-banded layout as numpy arrays (no GPU needed)

PyTorch is only plumbing here (device memory, streams); all compute is in libidash_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from ._lib import (COMPILE_DEFAULT, COMPILE_GROUPS_ALL, CONSTANT_BIDX, CT_BYTES, CT_WORDS, DECRYPT_AUTO, DECRYPT_IADD, DECRYPT_TENSOR, DECRYPT_TENSOR_PAIR,  # noqa: F401
                   KERNEL_AUTO, KERNEL_IMAD, KERNEL_TENSOR, KERNEL_TENSOR_RING, KERNEL_TENSOR_TILE, LAYOUT_PACKED, LAYOUT_RECORDS, N,
                   RECORD_BYTES, IdashB200Error)


class Context:
    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        self.device = device
        L.check(L.lib().idash_b200_init(C.byref(self._h), device))

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("Context is closed")
        return self._h

    def kernel_launches(self) -> int:
        return int(L.lib().idash_b200_kernel_launches(self.handle))

    def set_kernel(self, which: int) -> None:
        """KERNEL_AUTO / KERNEL_IMAD / KERNEL_TENSOR (idash_b200_set_kernel)."""
        L.check(L.lib().idash_b200_set_kernel(self.handle, int(which)))

    def last_kernel(self) -> int:
        return int(L.lib().idash_b200_last_kernel(self.handle))

    def set_decrypt_kernel(self, which: int) -> None:
        """DECRYPT_AUTO / DECRYPT_IADD / DECRYPT_TENSOR (idash_b200_set_decrypt_kernel)."""
        L.check(L.lib().idash_b200_set_decrypt_kernel(self.handle, int(which)))

    def last_decrypt_kernel(self) -> int:
        return int(L.lib().idash_b200_last_decrypt_kernel(self.handle))

    def timing_enable(self, max_launches: int) -> None:
        L.check(L.lib().idash_b200_timing_enable(self.handle, int(max_launches)))

    def timing_read(self, capacity: int = 4096):
        """Per-launch durations (ms) of the dominant kernels recorded since the last read."""
        buf = (C.c_float * capacity)()
        n = C.c_int(capacity)
        L.check(L.lib().idash_b200_timing_read(self.handle, buf, C.byref(n)))
        return [float(buf[i]) for i in range(n.value)]

    def check_device_status(self) -> None:
        L.check(L.lib().idash_b200_check_device_status(self.handle))

    def close(self):
        if self._h:
            L.lib().idash_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Model:
    """Device-resident block-banded model."""

    def __init__(self, ctx: Context, S: int, NR: int, RS: int, out_bidx, row_ptr, col, coef):
        self.ctx = ctx
        self.S, self.NR, self.RS = int(S), int(NR), int(RS)
        desc, keep = L.make_desc(S, NR, RS, out_bidx, row_ptr, col, coef)
        self.out_bidx = keep[0]
        self._h = C.c_void_p()
        L.check(L.lib().idash_b200_model_upload(ctx.handle, C.byref(desc), C.byref(self._h)))
        self._read_info()

    @classmethod
    def from_cache(cls, ctx: Context, path, key: int) -> "Model":
        """The cached packed model (idash_b200_layout_load + idash_b200_model_upload_layout); raises if the file does not match."""
        self = cls.__new__(cls)
        self.ctx = ctx
        lay = C.c_void_p()
        L.check(L.lib().idash_b200_layout_load(str(path).encode(), int(key), C.byref(lay)))
        self._h = C.c_void_p()
        L.check(L.lib().idash_b200_model_upload_layout(ctx.handle, lay, C.byref(self._h)))
        self._read_info()
        self.out_bidx = None
        return self

    def clone(self, ctx: Context) -> "Model":
        """The same model on another GPU (idash_b200_model_clone): the compiled layout is shared, the device arrays are uploaded again."""
        other = Model.__new__(Model)
        other.ctx = ctx
        other.S, other.NR, other.RS = getattr(self, "S", None), getattr(self, "NR", None), getattr(self, "RS", None)
        other.out_bidx = self.out_bidx
        other._h = C.c_void_p()
        L.check(L.lib().idash_b200_model_clone(ctx.handle, self.handle, C.byref(other._h)))
        other._read_info()
        return other

    def _read_info(self):
        info = L.ModelInfo()
        L.check(L.lib().idash_b200_model_get_info(self._h, C.byref(info)))
        self.info = info.as_dict()
        self.n_rows = self.info["n_rows"]

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("Model is freed")
        return self._h

    def free(self):
        if self._h:
            L.lib().idash_b200_model_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _ptr(a):
    return None if a is None else a.ctypes.data


def cloud_compute_score(ctx: Context, model: Model, in_ct: np.ndarray, in_index=None, in_var=None, slot_of_row=None,
                        out_ct: np.ndarray | None = None):
    """PACKED host path. in_ct [n_in, 2048] uint32. Returns (out_ct [n_rows, 2048], out_bidx, out_var)."""
    in_ct = np.ascontiguousarray(in_ct, np.uint32).reshape(-1, CT_WORDS)
    n_in = len(in_ct)
    idx = None if in_index is None else np.ascontiguousarray(in_index, np.uint32)
    var = None if in_var is None else np.ascontiguousarray(in_var, np.float64)
    if idx is not None and len(idx) != n_in:
        raise ValueError("in_index length mismatch")
    if var is not None and len(var) != n_in:
        raise ValueError("in_var length mismatch")
    n_out = model.n_rows
    if out_ct is None:
        out_ct = np.empty((n_out, CT_WORDS), np.uint32)
    if out_ct.dtype != np.uint32 or not out_ct.flags.c_contiguous or out_ct.size % CT_WORDS:
        raise ValueError("out_ct must be a C-contiguous uint32 array of whole ciphertexts")
    n_buf = out_ct.size // CT_WORDS            # the library checks it against the model's row count
    out_idx = np.zeros(n_buf, np.uint32)
    out_var = np.zeros(n_buf, np.float64)
    sor = None if slot_of_row is None else np.ascontiguousarray(slot_of_row, np.uint32)
    cin = L.Cts(LAYOUT_PACKED, _ptr(in_ct) if n_in else None, n_in, _ptr(idx), _ptr(var))
    cout = L.Cts(LAYOUT_PACKED, _ptr(out_ct) if n_buf else None, n_buf, _ptr(out_idx), _ptr(out_var))
    L.check(L.lib().idash_b200_cloud_eval_host(ctx.handle, model.handle, C.byref(cin), C.byref(cout), _ptr(sor)))
    return out_ct, out_idx, out_var


def cloud_compute_score_records(ctx: Context, model: Model, in_image: np.ndarray, slot_of_row=None,
                                out_image: np.ndarray | None = None) -> np.ndarray:
    """RECORDS host path: in_image = the bytes of encrypted_data.bin (uint8, including the 8-byte count);
    returns the bytes of encrypted_prediction.bin."""
    in_image = np.ascontiguousarray(in_image, np.uint8)
    n_in = int(in_image[:8].view("<u8")[0])
    if in_image.size != 8 + n_in * RECORD_BYTES:
        raise ValueError("encrypted data image has the wrong size")
    n_out = model.n_rows
    if out_image is None:
        out_image = np.empty(8 + n_out * RECORD_BYTES, np.uint8)
    out_image[:8].view("<u8")[0] = n_out
    sor = None if slot_of_row is None else np.ascontiguousarray(slot_of_row, np.uint32)
    cin = L.Cts(LAYOUT_RECORDS, in_image.ctypes.data + 8 if n_in else None, n_in, None, None)
    cout = L.Cts(LAYOUT_RECORDS, out_image.ctypes.data + 8 if n_out else None, n_out, None, None)
    L.check(L.lib().idash_b200_cloud_eval_host(ctx.handle, model.handle, C.byref(cin), C.byref(cout), _ptr(sor)))
    return out_image


def cloud_compute_score_device(ctx: Context, model: Model, in_ct, out_ct, in_index=None, in_var=None, out_index=None,
                               out_var=None, slot_of_row=None, stream=None) -> None:
    """PACKED device path on torch CUDA tensors (int32/uint32 words); enqueues on `stream` (default: torch's
    current stream) and returns without synchronising."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    n_in = in_ct.numel() // CT_WORDS
    n_out = out_ct.numel() // CT_WORDS
    dp = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    cin = L.Cts(LAYOUT_PACKED, dp(in_ct) if n_in else None, n_in, dp(in_index), dp(in_var))
    cout = L.Cts(LAYOUT_PACKED, dp(out_ct) if n_out else None, n_out, dp(out_index), dp(out_var))
    L.check(L.lib().idash_b200_cloud_eval_device(ctx.handle, model.handle, C.byref(cin), C.byref(cout), dp(slot_of_row),
                                                 C.c_void_p(stream)))


def cloud_compute_score_device_multi_model(ctx: Context, models, in_cts, out_cts, stream=None) -> None:
    """models[b] on (in_cts[b], out_cts[b]) for every b in one call (idash_b200_cloud_eval_device_multi_model): one launch of the ring
    kernel when the models have the same shape -- the population-stratified model sets of BASELINE configs[3]."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    n = len(models)
    if n != len(in_cts) or n != len(out_cts):
        raise ValueError("models / in_cts / out_cts length mismatch")
    cin = (L.Cts * n)()
    cout = (L.Cts * n)()
    for b in range(n):
        cin[b] = L.Cts(LAYOUT_PACKED, in_cts[b].data_ptr(), in_cts[b].numel() // CT_WORDS, None, None)
        cout[b] = L.Cts(LAYOUT_PACKED, out_cts[b].data_ptr(), out_cts[b].numel() // CT_WORDS, None, None)
    hm = (C.c_void_p * n)(*[m.handle.value for m in models])
    L.check(L.lib().idash_b200_cloud_eval_device_multi_model(ctx.handle, n, hm, cin, cout, C.c_void_p(stream)))


def cloud_compute_score_multi_device(ctxs, models, in_ct, out_ct, in_var=None, out_index=None, out_var=None) -> None:
    """One evaluation sharded over several GPUs of this process by contiguous target ranges (idash_b200_cloud_eval_multi_device):
    in_ct / out_ct (and the optional arrays) are torch CUDA tensors on ctxs[0]'s GPU, PACKED, inputs in identity order. models[g] is
    models[0].clone(ctxs[g]). Synchronous; the caller's work on in_ct must be complete (torch.cuda.synchronize())."""
    n = len(ctxs)
    if n != len(models) or n == 0:
        raise ValueError("one model per context")
    n_in = in_ct.numel() // CT_WORDS
    n_out = out_ct.numel() // CT_WORDS
    dp = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    cin = L.Cts(LAYOUT_PACKED, dp(in_ct) if n_in else None, n_in, None, dp(in_var))
    cout = L.Cts(LAYOUT_PACKED, dp(out_ct) if n_out else None, n_out, dp(out_index), dp(out_var))
    hc = (C.c_void_p * n)(*[c.handle.value for c in ctxs])
    hm = (C.c_void_p * n)(*[m.handle.value for m in models])
    L.check(L.lib().idash_b200_cloud_eval_multi_device(n, hc, hm, C.byref(cin), C.byref(cout)))


def cloud_compute_score_device_batched(ctx: Context, model: Model, in_cts, out_cts, stream=None) -> None:
    """The same model on several (input, output) sets of torch CUDA tensors [n, 2048] int32, PACKED, identity order, in one
    call (idash_b200_cloud_eval_device_batched): one launch of the ring kernel when it takes the model."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    n = len(in_cts)
    if n != len(out_cts):
        raise ValueError("in_cts / out_cts length mismatch")
    cin = (L.Cts * n)()
    cout = (L.Cts * n)()
    for b in range(n):
        cin[b] = L.Cts(LAYOUT_PACKED, in_cts[b].data_ptr(), in_cts[b].numel() // CT_WORDS, None, None)
        cout[b] = L.Cts(LAYOUT_PACKED, out_cts[b].data_ptr(), out_cts[b].numel() // CT_WORDS, None, None)
    L.check(L.lib().idash_b200_cloud_eval_device_batched(ctx.handle, model.handle, n, cin, cout, C.c_void_p(stream)))


def decrypt_predictions(ctx: Context, key, S: int, ct: np.ndarray, want_phase: bool = False, out_scores: np.ndarray | None = None):
    """PACKED host path. key [1024] in {0,1}; ct [n, 2048]. Returns scores [n, S] float32 (and phase [n, 1024]). out_scores: a
    caller-owned (e.g. pinned) C-contiguous float32 [n, S] array to receive the scores."""
    key = np.ascontiguousarray(key, np.int32)
    ct = np.ascontiguousarray(ct, np.uint32).reshape(-1, CT_WORDS)
    n = len(ct)
    if out_scores is not None:
        if out_scores.dtype != np.float32 or not out_scores.flags.c_contiguous or out_scores.shape != (n, S):
            raise ValueError("out_scores must be a C-contiguous float32 [n, S] array")
        scores = out_scores
    else:
        scores = np.empty((n, S), np.float32)
    phase = np.empty((n, N), np.uint32) if want_phase else None
    cin = L.Cts(LAYOUT_PACKED, _ptr(ct) if n else None, n, None, None)
    L.check(L.lib().idash_b200_decrypt_host(ctx.handle, _ptr(key), S, C.byref(cin), _ptr(scores), _ptr(phase)))
    return (scores, phase) if want_phase else scores


def decrypt_predictions_records(ctx: Context, key, S: int, image: np.ndarray, want_phase: bool = False):
    """RECORDS host path on the bytes of encrypted_prediction.bin. Returns (index [n], scores [n, S][, phase])."""
    key = np.ascontiguousarray(key, np.int32)
    image = np.ascontiguousarray(image, np.uint8)
    n = int(image[:8].view("<u8")[0])
    if image.size != 8 + n * RECORD_BYTES:
        raise ValueError("encrypted prediction image has the wrong size")
    scores = np.empty((n, S), np.float32)
    phase = np.empty((n, N), np.uint32) if want_phase else None
    cin = L.Cts(LAYOUT_RECORDS, image.ctypes.data + 8 if n else None, n, None, None)
    L.check(L.lib().idash_b200_decrypt_host(ctx.handle, _ptr(key), S, C.byref(cin), _ptr(scores), _ptr(phase)))
    index = np.ndarray((n,), "<u4", image, offset=8, strides=(RECORD_BYTES,)).copy() if n else np.zeros(0, np.uint32)
    return (index, scores, phase) if want_phase else (index, scores)


def decrypt_predictions_device(ctx: Context, key, S: int, ct, scores, phase=None, stream=None) -> None:
    import torch
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    key = np.ascontiguousarray(key, np.int32)
    n = ct.numel() // CT_WORDS
    cin = L.Cts(LAYOUT_PACKED, ct.data_ptr() if n else None, n, None, None)
    L.check(L.lib().idash_b200_decrypt_device(ctx.handle, _ptr(key), S, C.byref(cin),
                                              None if scores is None else scores.data_ptr(),
                                              None if phase is None else phase.data_ptr(), C.c_void_p(stream)))


@dataclass
class Layout:
    info: dict
    groups: np.ndarray    # GROUP_DTYPE
    entries: np.ndarray   # ENTRY_DTYPE
    var_ptr: np.ndarray
    var_ct: np.ndarray
    var_w: np.ndarray
    out_bidx: np.ndarray
    tiles: np.ndarray       # TILE_DTYPE; empty when the model is not eligible for the tensor-core kernel
    tile_rows: np.ndarray   # [n_tiles * 64] caller row or NO_ROW
    tile_bias: np.ndarray   # [n_tiles * 64]
    tile_coef: np.ndarray   # uint8 coefficient images
    tile_used: np.ndarray   # uint32 masks
    overflow_groups: np.ndarray = None    # IMAD groups of the rows no tile holds
    overflow_entries: np.ndarray = None
    feat_ptr: np.ndarray = None           # the model as the layout keeps it: per caller row, entries sorted by input bigIndex
    feat_bidx: np.ndarray = None
    feat_coef: np.ndarray = None
    bias: np.ndarray = None


def compile_layout(S, NR, RS, out_bidx, row_ptr, col, coef, flags: int = COMPILE_GROUPS_ALL, save_to=None, key: int = 0) -> Layout:
    """Runs the host-side model compiler only (no GPU) and copies the layout out as numpy arrays."""
    lib = L.lib()
    desc, keep = L.make_desc(S, NR, RS, out_bidx, row_ptr, col, coef)
    h = C.c_void_p()
    L.check(lib.idash_b200_layout_compile_ex(C.byref(desc), int(flags), C.byref(h)))
    if save_to is not None:
        try:
            L.check(lib.idash_b200_layout_save(h, str(save_to).encode(), int(key)))
        except Exception:
            lib.idash_b200_layout_free(h)
            raise
    return _export_layout(h)


def load_layout(path, key: int) -> Layout:
    """idash_b200_layout_load: the cached packed model as numpy arrays (raises IdashB200Error if it does not match)."""
    h = C.c_void_p()
    L.check(L.lib().idash_b200_layout_load(str(path).encode(), int(key), C.byref(h)))
    return _export_layout(h)


def _export_layout(h) -> Layout:
    lib = L.lib()
    try:
        info = L.ModelInfo()
        L.check(lib.idash_b200_layout_get_info(h, C.byref(info)))
        n = C.c_uint64()
        gp = lib.idash_b200_layout_groups(h, C.byref(n))
        groups = np.ctypeslib.as_array(C.cast(gp, C.POINTER(C.c_uint8)), (n.value * 64,)).view(L.GROUP_DTYPE).copy() \
            if n.value else np.zeros(0, L.GROUP_DTYPE)
        ep = lib.idash_b200_layout_entries(h, C.byref(n))
        entries = np.ctypeslib.as_array(C.cast(ep, C.POINTER(C.c_uint8)), (n.value * 32,)).view(L.ENTRY_DTYPE).copy() \
            if n.value else np.zeros(0, L.ENTRY_DTYPE)
        n_rows = info.n_rows
        vp = np.ctypeslib.as_array(C.cast(lib.idash_b200_layout_var_ptr(h), C.POINTER(C.c_uint64)), (n_rows + 1,)).copy()
        nv = C.c_uint64()
        vc_p = lib.idash_b200_layout_var_ct(h, C.byref(nv))
        if nv.value:
            vc = np.ctypeslib.as_array(C.cast(vc_p, C.POINTER(C.c_uint32)), (nv.value,)).copy()
            vw = np.ctypeslib.as_array(C.cast(lib.idash_b200_layout_var_w(h), C.POINTER(C.c_double)), (nv.value,)).copy()
        else:
            vc, vw = np.zeros(0, np.uint32), np.zeros(0, np.float64)
        ob = np.ctypeslib.as_array(C.cast(lib.idash_b200_layout_out_bidx(h), C.POINTER(C.c_uint32)), (n_rows,)).copy() \
            if n_rows else np.zeros(0, np.uint32)
        def arr(ptr, ctype, count, dtype):
            if not count:
                return np.zeros(0, dtype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), (count,)).view(dtype).copy()
        nt = C.c_uint64()
        tiles = arr(lib.idash_b200_layout_tiles(h, C.byref(nt)), C.c_uint8, nt.value * 32, L.TILE_DTYPE)
        t_rows = arr(lib.idash_b200_layout_tile_rows(h), C.c_uint32, nt.value * L.TILE_ROWS, np.uint32)
        t_bias = arr(lib.idash_b200_layout_tile_bias(h), C.c_int32, nt.value * L.TILE_ROWS, np.int32)
        nb = C.c_uint64()
        cp = lib.idash_b200_layout_tile_coef(h, C.byref(nb))
        t_coef = arr(cp, C.c_uint8, nb.value, np.uint8)
        nu = C.c_uint64()
        up = lib.idash_b200_layout_tile_used(h, C.byref(nu))
        t_used = arr(up, C.c_uint32, nu.value, np.uint32)
        n = C.c_uint64()
        ogp = lib.idash_b200_layout_overflow_groups(h, C.byref(n))
        o_groups = arr(ogp, C.c_uint8, n.value * 64, L.GROUP_DTYPE)
        oep = lib.idash_b200_layout_overflow_entries(h, C.byref(n))
        o_entries = arr(oep, C.c_uint8, n.value * 32, L.ENTRY_DTYPE)
        f_ptr = arr(lib.idash_b200_layout_feat_ptr(h), C.c_uint64, n_rows + 1, np.uint64)
        fb_p = lib.idash_b200_layout_feat_bidx(h, C.byref(n))
        f_bidx = arr(fb_p, C.c_uint32, n.value, np.uint32)
        f_coef = arr(lib.idash_b200_layout_feat_coef(h), C.c_int32, n.value, np.int32)
        bias = arr(lib.idash_b200_layout_bias(h), C.c_int32, n_rows, np.int32)
        return Layout(info.as_dict(), groups, entries, vp, vc, vw, ob, tiles, t_rows, t_bias, t_coef, t_used, o_groups, o_entries,
                      f_ptr, f_bidx, f_coef, bias)
    finally:
        lib.idash_b200_layout_free(h)
