"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes front-end to
  * oracle/idash_oracle.c  -- our plain-C restatement of the reference algorithm ("port"), and
  * oracle/_ref/libidash_ref.so -- the UNMODIFIED reference compiled by oracle/build_ref.sh
    ("reference"), when it has been built (this container; it travels prebuilt to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module. Nothing in idash2019_2_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
BUILD = HERE / "_build"
LIB_PORT = BUILD / "libidash_oracle.so"
LIB_REF = HERE / "_ref" / "libidash_ref.so"
REF_BIN = HERE / "_ref" / "bin"

N = 1024
CT_WORDS = 2048
CONSTANT_BIDX = 0xFFFFFFFF

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build_port(force: bool = False) -> Path:
    """gcc -O3 -fopenmp oracle/idash_oracle.c -> oracle/_build/libidash_oracle.so"""
    src = HERE / "idash_oracle.c"
    if force or not LIB_PORT.exists() or LIB_PORT.stat().st_mtime < src.stat().st_mtime:
        BUILD.mkdir(exist_ok=True)
        subprocess.check_call(["/usr/bin/gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared",
                               "-std=c11", "-Wall", str(src), "-o", str(LIB_PORT)])
    return LIB_PORT


def build_ref() -> bool:
    """Runs oracle/build_ref.sh (no-op when /root/reference is absent). True if the .so exists."""
    subprocess.check_call(["bash", str(HERE / "build_ref.sh")])
    return LIB_REF.exists()


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        build_port()
        lib = C.CDLL(str(LIB_PORT))
        lib.oracle_cloud_compute_score.restype = C.c_int
        lib.oracle_cloud_compute_score.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, _u32p, _u32p,
                                                   _f64p, C.c_uint64, _u64p, _u32p, _i32p, _u32p, _f64p, C.c_int]
        lib.oracle_tlwe_phase_exact.restype = None
        lib.oracle_tlwe_phase_exact.argtypes = [_i32p, C.c_uint64, _u32p, _u32p, C.c_int]
        lib.oracle_decode_scores.restype = None
        lib.oracle_decode_scores.argtypes = [C.c_uint32, C.c_uint64, _u32p, _f32p]
        lib.oracle_max_threads.restype = C.c_int
        _port = lib
    return _port


def have_ref() -> bool:
    return LIB_REF.exists()


def ref():
    global _ref
    if _ref is None:
        if not LIB_REF.exists():
            raise FileNotFoundError(f"{LIB_REF} not built (run oracle/build_ref.sh where /root/reference exists)")
        lib = C.CDLL(str(LIB_REF))
        lib.ref_cloud_compute_score.restype = C.c_double
        lib.ref_cloud_compute_score.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, _u32p, _u32p, _f64p,
                                                C.c_uint64, _u32p, _u64p, _u32p, _i32p, C.c_void_p, C.c_void_p]
        lib.ref_decrypt_predictions.restype = C.c_double
        lib.ref_decrypt_predictions.argtypes = [C.c_uint32, _i32p, C.c_uint64, _u32p, _f32p]
        lib.ref_tlwe_phase.restype = None
        lib.ref_tlwe_phase.argtypes = [_i32p, C.c_uint64, _u32p, C.c_int, _u32p]
        lib.ref_model_load.restype = C.c_void_p
        lib.ref_model_load.argtypes = [C.c_char_p, C.c_char_p]
        lib.ref_model_rows.restype = C.c_uint64
        lib.ref_model_rows.argtypes = [C.c_void_p]
        lib.ref_model_nnz.restype = C.c_uint64
        lib.ref_model_nnz.argtypes = [C.c_void_p]
        lib.ref_model_geometry.restype = None
        lib.ref_model_geometry.argtypes = [C.c_void_p, _u32p]
        lib.ref_model_export.restype = None
        lib.ref_model_export.argtypes = [C.c_void_p, _u32p, _u64p, _u32p, _i32p]
        lib.ref_model_free.restype = None
        lib.ref_model_free.argtypes = [C.c_void_p]
        lib.ref_constants.restype = None
        lib.ref_constants.argtypes = [_i32p]
        lib.ref_max_threads.restype = C.c_int
        _ref = lib
    return _ref


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------
def cloud_port(S, NR, RS, in_idx, in_ct, in_var, row_ptr, col, coef, threads=None):
    """Restated cloud_compute_score -> (out_ct [n_out,2048] u32, out_var [n_out] f64)."""
    in_idx, in_ct = _c(in_idx, np.uint32), _c(in_ct, np.uint32).reshape(-1, CT_WORDS)
    in_var = _c(in_var, np.float64)
    row_ptr, col, coef = _c(row_ptr, np.uint64), _c(col, np.uint32), _c(coef, np.int32)
    n_out = len(row_ptr) - 1
    out_ct = np.empty((n_out, CT_WORDS), np.uint32)
    out_var = np.empty(n_out, np.float64)
    if len(col) == 0:
        col, coef = np.zeros(1, np.uint32), np.zeros(1, np.int32)
    err = port().oracle_cloud_compute_score(S, NR, RS, len(in_idx), in_idx, in_ct if len(in_ct) else
                                            np.zeros((1, CT_WORDS), np.uint32), in_var if len(in_var) else
                                            np.zeros(1), n_out, row_ptr, col, coef,
                                            out_ct if n_out else np.zeros((1, CT_WORDS), np.uint32),
                                            out_var if n_out else np.zeros(1), threads or host_threads())
    if err:
        raise KeyError("model references an input ciphertext that is not present")
    return out_ct, out_var


def cloud_ref(S, NR, RS, in_idx, in_ct, in_var, out_bidx, row_ptr, col, coef, want_output=True):
    """The reference's own cloud_compute_score -> (out_ct, out_var, fhe_wall_seconds)."""
    in_idx, in_ct = _c(in_idx, np.uint32), _c(in_ct, np.uint32).reshape(-1, CT_WORDS)
    in_var = _c(in_var, np.float64)
    out_bidx = _c(out_bidx, np.uint32)
    row_ptr, col, coef = _c(row_ptr, np.uint64), _c(col, np.uint32), _c(coef, np.int32)
    n_out = len(out_bidx)
    if want_output:
        out_ct = np.empty((n_out, CT_WORDS), np.uint32)
        out_var = np.empty(n_out, np.float64)
        p_ct, p_var = out_ct.ctypes.data, out_var.ctypes.data
    else:
        out_ct = out_var = None
        p_ct = p_var = None
    t = ref().ref_cloud_compute_score(S, NR, RS, len(in_idx), in_idx, in_ct, in_var, n_out, out_bidx, row_ptr, col,
                                      coef, p_ct, p_var)
    return out_ct, out_var, t


def phase_exact_port(key, ct, threads=None):
    key, ct = _c(key, np.int32), _c(ct, np.uint32).reshape(-1, CT_WORDS)
    ph = np.empty((len(ct), N), np.uint32)
    if len(ct):
        port().oracle_tlwe_phase_exact(key, len(ct), ct, ph, threads or host_threads())
    return ph


def decode_port(S, phase):
    phase = _c(phase, np.uint32).reshape(-1, N)
    sc = np.empty((len(phase), S), np.float32)
    if len(phase):
        port().oracle_decode_scores(S, len(phase), phase, sc)
    return sc


def phase_ref(key, ct, use_fft: bool):
    key, ct = _c(key, np.int32), _c(ct, np.uint32).reshape(-1, CT_WORDS)
    ph = np.empty((len(ct), N), np.uint32)
    ref().ref_tlwe_phase(key, len(ct), ct, 1 if use_fft else 0, ph)
    return ph


def decrypt_ref(S, key, ct):
    key, ct = _c(key, np.int32), _c(ct, np.uint32).reshape(-1, CT_WORDS)
    assert len(ct) % 3 == 0
    sc = np.empty((len(ct), S), np.float32)
    t = ref().ref_decrypt_predictions(S, key, len(ct), ct, sc)
    return sc, t


def read_model_ref(params_file, model_dir):
    """The reference's read_params + read_model -> (geometry[7], out_bidx, row_ptr, col, coef)."""
    lib = ref()
    h = lib.ref_model_load(str(params_file).encode(), str(model_dir).encode())
    try:
        rows, nnz = lib.ref_model_rows(h), lib.ref_model_nnz(h)
        g = np.zeros(7, np.uint32)
        lib.ref_model_geometry(h, g)
        out_bidx = np.zeros(rows, np.uint32)
        row_ptr = np.zeros(rows + 1, np.uint64)
        col = np.zeros(max(nnz, 1), np.uint32)
        coef = np.zeros(max(nnz, 1), np.int32)
        lib.ref_model_export(h, out_bidx, row_ptr, col, coef)
        return g, out_bidx, row_ptr, col[:nnz], coef[:nnz]
    finally:
        lib.ref_model_free(h)


def run_ref_bin(name, args, cwd, threads=None):
    """Runs one of the reference's CLI stages (keygen/encrypt/cloud/decrypt) in `cwd`."""
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(threads or host_threads())
    out = subprocess.run([str(REF_BIN / name), *map(str, args)], cwd=str(cwd), env=env, check=True,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    return out
