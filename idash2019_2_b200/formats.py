"""Readers / writers for the reference's on-disk formats touched by the hot path.

    params.bin                 write_params_ostream / read_params_istream   eval/idash.cpp:128-186, 200-261
    keys.bin                   write_key / read_key                         eval/idash.cpp:476-502
                               + TLWE key text header and body              tfhe_io.cpp:243-252, 395-416
    encrypted_data.bin         write/read_encrypted_data                    eval/idash.cpp:513-556
    encrypted_prediction.bin   write/read_encrypted_predictions             eval/idash.cpp:568-613
                               one record = u32 index, i32 84, 2048 words, f64 variance (tfhe_io.cpp:303-323)
    <pos>_<variant>.hr         parse_vw read()                              eval/parse_vw.cpp:8-30

Ciphertext files are handled as ONE byte image (np.uint8) that the C ABI consumes directly
(IDASH_B200_LAYOUT_RECORDS); nothing here allocates per-polynomial objects.
"""
from __future__ import annotations

import re
import struct
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

N = 1024
CT_WORDS = 2048
RECORD_BYTES = 8208
TLWE_SAMPLE_UID = 84
TLWE_KEY_UID = 85
CONSTANT_BIDX = 0xFFFFFFFF


@dataclass
class Params:
    """IdashParams (eval/idash.h:45-111): geometry + position -> bigIndex maps."""
    NUM_SAMPLES: int = 0
    NUM_INPUT_POSITIONS: int = 0
    NUM_OUTPUT_POSITIONS: int = 0
    NUM_INPUT_FEATURES: int = 0
    NUM_OUTPUT_FEATURES: int = 0
    NUM_REGIONS: int = 0
    REGION_SIZE: int = 0
    in_positions: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint64))      # file (hash) order
    in_bidx: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.uint32))
    out_positions: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint64))     # targets-file order
    out_names: list = field(default_factory=list)
    out_bidx: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.uint32))

    def in_index(self) -> dict:
        return {int(p): self.in_bidx[i] for i, p in enumerate(self.in_positions)}

    def out_index(self) -> dict:
        return {int(p): self.out_bidx[i] for i, p in enumerate(self.out_positions)}


def parse_params(buf: bytes, off: int = 0):
    """Returns (Params, offset after the params stream)."""
    p = Params()
    (p.NUM_SAMPLES, p.NUM_INPUT_POSITIONS, p.NUM_OUTPUT_POSITIONS, p.NUM_INPUT_FEATURES, p.NUM_OUTPUT_FEATURES,
     p.NUM_REGIONS, p.REGION_SIZE) = struct.unpack_from("<7I", buf, off)
    off += 28
    rec = np.dtype([("pos", "<u8"), ("bidx", "<u4", (3,))])
    a = np.frombuffer(buf, rec, p.NUM_INPUT_POSITIONS, off)
    off += rec.itemsize * p.NUM_INPUT_POSITIONS
    p.in_positions, p.in_bidx = a["pos"].copy(), a["bidx"].copy()
    pos, names, bidx = [], [], []
    for _ in range(p.NUM_OUTPUT_POSITIONS):
        pos.append(struct.unpack_from("<Q", buf, off)[0])
        off += 8
        end = buf.index(b"\0", off)
        names.append(buf[off:end].decode())
        off = end + 1
        bidx.append(struct.unpack_from("<3I", buf, off))
        off += 12
    p.out_positions = np.array(pos, np.uint64)
    p.out_names = names
    p.out_bidx = np.array(bidx, np.uint32).reshape(-1, 3)
    return p, off


def read_params(path) -> Params:
    return parse_params(Path(path).read_bytes())[0]


def serialize_params(p: Params) -> bytes:
    out = [struct.pack("<7I", p.NUM_SAMPLES, p.NUM_INPUT_POSITIONS, p.NUM_OUTPUT_POSITIONS, p.NUM_INPUT_FEATURES,
                       p.NUM_OUTPUT_FEATURES, p.NUM_REGIONS, p.REGION_SIZE)]
    rec = np.zeros(len(p.in_positions), np.dtype([("pos", "<u8"), ("bidx", "<u4", (3,))]))
    rec["pos"], rec["bidx"] = p.in_positions, p.in_bidx
    out.append(rec.tobytes())
    for pos, name, b in zip(p.out_positions, p.out_names, p.out_bidx):
        out.append(struct.pack("<Q", int(pos)) + name.encode() + b"\0" + struct.pack("<3I", *map(int, b)))
    return b"".join(out)


def read_key(path):
    """keys.bin -> (Params, key[1024] int32, tlwe text properties)."""
    buf = Path(path).read_bytes()
    p, off = parse_params(buf)
    m = re.compile(rb"-----BEGIN TLWEPARAMS-----\n(.*?)-----END TLWEPARAMS-----\n", re.S).match(buf, off)
    if not m:
        raise ValueError("keys.bin: TLWEPARAMS block not found")
    props = dict(line.split(b": ", 1) for line in m.group(1).splitlines() if b": " in line)
    props = {k.decode(): v.decode().strip() for k, v in props.items()}
    off = m.end()
    (uid,) = struct.unpack_from("<i", buf, off)
    if uid != TLWE_KEY_UID:
        raise ValueError(f"keys.bin: bad TLWE key type uid {uid}")
    n, k = int(props["N"]), int(props["k"])
    if n != N or k != 1:
        raise ValueError(f"keys.bin: unsupported TLWE parameters N={n} k={k}")
    key = np.frombuffer(buf, "<i4", n, off + 4).copy()
    return p, key, props


def read_ct_image(path) -> np.ndarray:
    """Whole encrypted_*.bin as a uint8 image whose record stream (image[8:]) sits at an address = 8 mod 16,
    i.e. the word arrays are 16-byte aligned -- the shape the C ABI's RECORDS layout takes."""
    size = Path(path).stat().st_size
    backing = np.empty(size + 32, np.uint8)
    shift = (-backing.ctypes.data) % 16
    img = backing[shift:shift + size]
    with open(path, "rb") as f:
        f.readinto(memoryview(img))
    check_ct_image(img)
    return img


def aligned_image(n_records: int) -> np.ndarray:
    size = 8 + n_records * RECORD_BYTES
    backing = np.empty(size + 32, np.uint8)
    shift = (-backing.ctypes.data) % 16
    img = backing[shift:shift + size]
    img[:8].view("<u8")[0] = n_records
    return img


def check_ct_image(img: np.ndarray) -> int:
    n = int(img[:8].view("<u8")[0])
    if img.size != 8 + n * RECORD_BYTES:
        raise ValueError(f"ciphertext file: size {img.size} does not match {n} records")
    if n:
        uid = np.ndarray((n,), "<i4", img, offset=12, strides=(RECORD_BYTES,))
        if not (uid == TLWE_SAMPLE_UID).all():
            raise ValueError("ciphertext file: bad TLWE sample type uid")   # tfhe_io.cpp:308 aborts
    return n


def image_views(img: np.ndarray):
    """(index [n] u32, words [n, 2048] u32, variance [n] f64) strided views into a ciphertext file image."""
    n = int(img[:8].view("<u8")[0])
    index = np.ndarray((n,), "<u4", img, offset=8, strides=(RECORD_BYTES,))
    words = np.ndarray((n, CT_WORDS), "<u4", img, offset=16, strides=(RECORD_BYTES, 4))
    var = np.ndarray((n,), "<f8", img, offset=16 + 4 * CT_WORDS, strides=(RECORD_BYTES,))
    return index, words, var


def build_ct_image(index, words, var) -> np.ndarray:
    index = np.asarray(index, np.uint32)
    img = aligned_image(len(index))
    if len(index):
        i, w, v = image_views(img)
        i[:] = index
        np.ndarray((len(index),), "<i4", img, offset=12, strides=(RECORD_BYTES,))[:] = TLWE_SAMPLE_UID
        w[:] = np.asarray(words, np.uint32).reshape(-1, CT_WORDS)
        v[:] = var
    return img


# ---- .hr model files -----------------------------------------------------------------------------
_INT_PREFIX = re.compile(r"[ \t\n\v\f\r]*([+-]?\d+)")


def read_hr(path) -> dict:
    """parse_vw read() (eval/parse_vw.cpp:8-30): per line sscanf("%s %d") -> coefs[name] = value.
    `%d` keeps the integer prefix of "-107.0"; a later duplicate name overwrites. If `%d` fails to
    convert, the reference stores an uninitialised int -- we raise instead."""
    coefs = {}
    with open(path, "r") as f:
        for line in f:
            # fgets(buf, 256): lines longer than 255 chars are split by the reference; not produced by the trainer
            toks = line.split(None, 1)
            if not toks:
                continue   # sscanf matches nothing: the reference re-stores the previous (stale) pair; harmless
            m = _INT_PREFIX.match(toks[1]) if len(toks) > 1 else None
            if not m:
                raise ValueError(f"{path}: cannot parse coefficient in line {line!r}")
            v = int(m.group(1))
            coefs[toks[0]] = ((v + 2 ** 31) % 2 ** 32) - 2 ** 31
    return coefs


def read_model(params: Params, model_dir):
    """read_model (eval/idash.cpp:66-90) -> CSR (out_bidx ascending; per row Constant first, then input
    bigIndex ascending)."""
    in_index = params.in_index()
    rows = []
    for pos, bidx3 in zip(params.out_positions, params.out_bidx):
        for snp in range(3):
            coefs = read_hr(Path(model_dir) / f"{int(pos)}_{snp}.hr")
            ent = {}
            for name, v in coefs.items():
                if name == "Constant":
                    ent[CONSTANT_BIDX] = v
                else:
                    p_s, v_s = name.split("_", 1)
                    ent[int(in_index[int(p_s)][int(v_s)])] = v      # KeyError like std::out_of_range
            rows.append((int(bidx3[snp]), ent))
    rows.sort(key=lambda r: r[0])
    out_bidx = np.array([r[0] for r in rows], np.uint32)
    row_ptr = np.zeros(len(rows) + 1, np.uint64)
    col, coef = [], []
    for i, (_, ent) in enumerate(rows):
        keys = sorted(ent, key=lambda k: (k != CONSTANT_BIDX, k))
        col.extend(keys)
        coef.extend(ent[k] for k in keys)
        row_ptr[i + 1] = len(col)
    return out_bidx, row_ptr, np.array(col, np.uint32), np.array(coef, np.int32)
