"""ctypes binding of libidash_b200.so (the C ABI declared in include/idash_b200.h).

The library is the product: there is no Python or CPU fallback. If it has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C idash2019_2_b200/csrc`) the
import of this module raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

import os

PKG = Path(__file__).resolve().parent
# tools/ (tuning sweeps, knock-out and trace runs) load the profiling build, which honours the IDASH_B200_* debug variables
LIB_PATH = PKG / "lib" / ("libidash_b200_prof.so" if os.environ.get("IDASH_B200_USE_PROFILE_LIB") == "1" else "libidash_b200.so")
if os.environ.get("IDASH_B200_LIB"):          # tools/: A/B timing of two builds of the library in one GPU session
    LIB_PATH = Path(os.environ["IDASH_B200_LIB"]).resolve()

N = 1024
CT_WORDS = 2048
CT_BYTES = 8192
RECORD_BYTES = 8208
CONSTANT_BIDX = 0xFFFFFFFF
ONE_IN_T32 = 262144
NO_ROW = 0xFFFFFFFF

OK, ERR_INVALID, ERR_CUDA, ERR_MISSING_INPUT, ERR_NOMEM = 0, -1, -2, -3, -4
LAYOUT_PACKED, LAYOUT_RECORDS = 0, 1


class IdashB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"idash_b200 error {code}: {msg}")
        self.code = code


class Cts(C.Structure):
    _fields_ = [("layout", C.c_int32), ("data", C.c_void_p), ("count", C.c_uint64), ("index", C.c_void_p),
                ("variance", C.c_void_p)]


class ModelDesc(C.Structure):
    _fields_ = [("num_samples", C.c_uint32), ("num_regions", C.c_uint32), ("region_size", C.c_uint32),
                ("n_rows", C.c_uint64), ("out_bidx", C.c_void_p), ("row_ptr", C.c_void_p), ("col", C.c_void_p),
                ("coef", C.c_void_p)]


class ModelInfo(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("nnz", C.c_uint64), ("n_groups", C.c_uint64), ("n_entries", C.c_uint64),
                ("ct_min", C.c_uint32), ("ct_max", C.c_uint32), ("max_entries_per_group", C.c_uint32),
                ("shifts_aligned", C.c_uint32), ("device_bytes", C.c_uint64), ("n_tiles", C.c_uint64),
                ("tile_kmax", C.c_uint32), ("ring_ok", C.c_uint32), ("n_overflow_rows", C.c_uint64), ("groups_all", C.c_uint32),
                ("pad", C.c_uint32)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


ENTRY_DTYPE = np.dtype([("ct", "<u4"), ("shift", "<u4"), ("coef", "<i4", (6,))])
GROUP_DTYPE = np.dtype([("entry_begin", "<u4"), ("n_a", "<u4"), ("n_ab", "<u4"), ("n_b", "<u4"), ("row", "<u4", (6,)),
                        ("bias", "<i4", (6,))])
TILE_DTYPE = np.dtype([("f_base", "<u4"), ("K", "<u4"), ("b_off", "<u8"), ("used_off", "<u4"), ("n_valid", "<u4"),
                       ("flags", "<u4"), ("pad", "<u4")])
assert ENTRY_DTYPE.itemsize == 32 and GROUP_DTYPE.itemsize == 64 and TILE_DTYPE.itemsize == 32
TILE_ROWS, TILE_KMAX, RING_KMAX = 64, 256, 224
COMPILE_DEFAULT, COMPILE_GROUPS_ALL = 0, 1
KERNEL_AUTO, KERNEL_IMAD, KERNEL_TENSOR, KERNEL_TENSOR_TILE, KERNEL_TENSOR_RING = 0, 1, 2, 3, 4
DECRYPT_AUTO, DECRYPT_IADD, DECRYPT_TENSOR, DECRYPT_TENSOR_PAIR = 0, 1, 2, 3

# every symbol include/idash_b200.h and include/idash_b200_layout.h declare
EXPORTS = {
    "idash_b200_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "idash_b200_destroy": (C.c_int, [C.c_void_p]),
    "idash_b200_last_error": (C.c_char_p, []),
    "idash_b200_kernel_launches": (C.c_uint64, [C.c_void_p]),
    "idash_b200_set_kernel": (C.c_int, [C.c_void_p, C.c_int]),
    "idash_b200_last_kernel": (C.c_int, [C.c_void_p]),
    "idash_b200_set_decrypt_kernel": (C.c_int, [C.c_void_p, C.c_int]),
    "idash_b200_last_decrypt_kernel": (C.c_int, [C.c_void_p]),
    "idash_b200_timing_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "idash_b200_timing_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "idash_b200_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "idash_b200_host_free": (C.c_int, [C.c_void_p]),
    "idash_b200_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "idash_b200_host_unregister": (C.c_int, [C.c_void_p]),
    "idash_b200_model_upload": (C.c_int, [C.c_void_p, C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]),
    "idash_b200_model_upload_layout": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "idash_b200_model_clone": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "idash_b200_model_free": (C.c_int, [C.c_void_p]),
    "idash_b200_model_get_info": (C.c_int, [C.c_void_p, C.POINTER(ModelInfo)]),
    "idash_b200_cloud_eval_host": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Cts), C.POINTER(Cts), C.c_void_p]),
    "idash_b200_cloud_eval_device": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Cts), C.POINTER(Cts), C.c_void_p,
                                               C.c_void_p]),
    "idash_b200_model_input_range": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "idash_b200_cloud_eval_device_multi_model": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(Cts), C.POINTER(Cts), C.c_void_p]),
    "idash_b200_cloud_eval_multi_device": (C.c_int, [C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(Cts), C.POINTER(Cts)]),
    "idash_b200_cloud_eval_host_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Cts), C.POINTER(Cts), C.c_uint64, C.c_uint64]),
    "idash_b200_cloud_eval_device_batched": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(Cts), C.POINTER(Cts), C.c_void_p]),
    "idash_b200_check_device_status": (C.c_int, [C.c_void_p]),
    "idash_b200_decrypt_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(Cts), C.c_void_p, C.c_void_p]),
    "idash_b200_decrypt_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(Cts), C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "idash_b200_layout_compile": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]),
    "idash_b200_layout_compile_ex": (C.c_int, [C.POINTER(ModelDesc), C.c_uint32, C.POINTER(C.c_void_p)]),
    "idash_b200_layout_ensure_groups_all": (C.c_int, [C.c_void_p]),
    "idash_b200_layout_save": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64]),
    "idash_b200_layout_load": (C.c_int, [C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "idash_b200_layout_feat_ptr": (C.c_void_p, [C.c_void_p]),
    "idash_b200_layout_feat_bidx": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "idash_b200_layout_feat_coef": (C.c_void_p, [C.c_void_p]),
    "idash_b200_layout_bias": (C.c_void_p, [C.c_void_p]),
    "idash_b200_layout_overflow_groups": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "idash_b200_layout_overflow_entries": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "idash_b200_layout_free": (C.c_int, [C.c_void_p]),
    "idash_b200_layout_get_info": (C.c_int, [C.c_void_p, C.POINTER(ModelInfo)]),
    "idash_b200_layout_groups": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "idash_b200_layout_entries": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "idash_b200_layout_var_ptr": (C.c_void_p, [C.c_void_p]),
    "idash_b200_layout_var_ct": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "idash_b200_layout_var_w": (C.c_void_p, [C.c_void_p]),
    "idash_b200_layout_out_bidx": (C.c_void_p, [C.c_void_p]),
    "idash_b200_layout_tiles": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "idash_b200_layout_tile_rows": (C.c_void_p, [C.c_void_p]),
    "idash_b200_layout_tile_bias": (C.c_void_p, [C.c_void_p]),
    "idash_b200_layout_tile_coef": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "idash_b200_layout_tile_used": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "idash_b200_layout_feat_used": (C.c_void_p, [C.c_void_p, C.POINTER(C.c_uint64)]),
}

_lib = None


def lib():
    """Loads libidash_b200.so; raises if it is missing (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is not built: run `make -C {PKG / 'csrc'}` (needs nvcc); "
                              "idash2019_2_b200 has no CPU fallback")
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in EXPORTS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise IdashB200Error(rc, lib().idash_b200_last_error().decode(errors="replace"))


def make_desc(S, NR, RS, out_bidx, row_ptr, col, coef):
    """ModelDesc + the numpy arrays that back it (keep them alive while the desc is in use)."""
    out_bidx = np.ascontiguousarray(out_bidx, np.uint32)
    row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
    col = np.ascontiguousarray(col, np.uint32)
    coef = np.ascontiguousarray(coef, np.int32)
    if len(row_ptr) != len(out_bidx) + 1:
        raise ValueError("row_ptr must have n_rows + 1 entries")
    if len(col) != len(coef) or (len(row_ptr) and int(row_ptr[-1]) != len(col)):
        raise ValueError("col / coef / row_ptr are inconsistent")
    d = ModelDesc(S, NR, RS, len(out_bidx), out_bidx.ctypes.data, row_ptr.ctypes.data,
                  col.ctypes.data if len(col) else None, coef.ctypes.data if len(coef) else None)
    return d, (out_bidx, row_ptr, col, coef)
