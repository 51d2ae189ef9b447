"""Synthetic inputs for the cloud-evaluation path.

Positions, genotypes and integer logistic-regression models with the statistical shape of the
iDASH data: about 5 target SNPs per tag SNP, every target predicted from its `neighbors` nearest
tag SNPs (train/to_vw.py:33-34 in the reference), three one-hot variants per SNP, one `.hr` text
file per (target, variant) (train/test_vw_hr.py:41-47).

Feature indexing follows keygen_ph1 (eval/idash.cpp:337-339, 388-390):
    output bigIndex = 3 * target_line + variant,   input bigIndex = 3 * tag_line + variant,
    "Constant" = 0xFFFFFFFF (eval/idash.h:93).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np

CONSTANT_BIDX = 0xFFFFFFFF
N = 1024
CT_WORDS = 2048


@dataclass
class Geometry:
    """NUM_SAMPLES / NUM_REGIONS / REGION_SIZE as keygen_ph1 derives them (eval/idash.cpp:409-410)."""
    S: int
    T: int
    G: int

    @property
    def NR(self) -> int:
        return N // self.S

    @property
    def RS(self) -> int:
        return N // self.NR

    @property
    def n_in_ct(self) -> int:
        """ceil(NUM_INPUT_FEATURES / NR) with NUM_INPUT_FEATURES = 3T + 1 (eval/idash.cpp:406, 621)."""
        return (3 * self.T + 1 + self.NR - 1) // self.NR

    @property
    def n_in_ct_used(self) -> int:
        """ciphertexts that actually exist: indices of the 3T real features (eval/idash.cpp:630-637)."""
        return (3 * self.T - 1) // self.NR + 1

    @property
    def n_out_ct(self) -> int:
        return 3 * self.G


def make_positions(T: int, G: int, seed: int = 1234):
    """T+G distinct sorted positions; a random subset of size T are tags, the rest targets."""
    rng = np.random.default_rng(seed)
    span = 40 * (T + G)
    pos = np.sort(rng.choice(span, size=T + G, replace=False)).astype(np.uint64) + np.uint64(16_000_000)
    is_tag = np.zeros(T + G, bool)
    is_tag[rng.choice(T + G, size=T, replace=False)] = True
    return pos[is_tag], pos[~is_tag]


def nearest_tag_windows(tag_pos: np.ndarray, target_pos: np.ndarray, n: int) -> np.ndarray:
    """For every target, the index of the first tag of its window of `n` nearest tags by |dpos|
    (stable on ties, like pandas' sort in train/to_vw.py:33-34). The n nearest elements of a sorted
    array always form a contiguous window."""
    T = len(tag_pos)
    n = min(n, T)
    tp = tag_pos.astype(np.int64)
    gp = target_pos.astype(np.int64)
    i = np.searchsorted(tp, gp)
    lo = np.clip(i - n, 0, T - n)                    # leftmost possible window start
    hi = np.clip(i, 0, T - n)                        # rightmost possible window start
    # slide right while the tag entering on the right is strictly closer than the one leaving
    start = lo.copy()
    for _ in range(n + 1):
        can = start < hi
        left = np.abs(tp[np.minimum(start, T - 1)] - gp)
        right = np.abs(tp[np.minimum(start + n, T - 1)] - gp)
        move = can & (right < left)
        if not move.any():
            break
        start = start + move
    return start.astype(np.int64)


@dataclass
class CsrModel:
    """Model as CSR over output bigIndex rows (eval/idash.h:129-134 flattened)."""
    out_bidx: np.ndarray   # [n_out] uint32, ascending
    row_ptr: np.ndarray    # [n_out+1] uint64
    col: np.ndarray        # [nnz] uint32 input bigIndex or CONSTANT_BIDX
    coef: np.ndarray       # [nnz] int32

    @property
    def n_out(self) -> int:
        return len(self.out_bidx)

    @property
    def nnz(self) -> int:
        return len(self.col)

    def rows(self, lo: int, hi: int) -> "CsrModel":
        e0, e1 = int(self.row_ptr[lo]), int(self.row_ptr[hi])
        return CsrModel(self.out_bidx[lo:hi].copy(), (self.row_ptr[lo:hi + 1] - self.row_ptr[lo]).astype(np.uint64),
                        self.col[e0:e1].copy(), self.coef[e0:e1].copy())


def make_model(tag_pos, target_pos, n: int, seed: int = 1234, coef_range: int = 200, bias_range: int = 500,
               drop_zeros: bool = True) -> CsrModel:
    """Banded synthetic model: for each (target, variant) a bias U[-bias_range, bias_range) and one
    coefficient U[-coef_range, coef_range) per (tag in window, variant); zeros omitted like
    train/test_vw_hr.py:44. Row layout: Constant first, then input bigIndex ascending."""
    rng = np.random.default_rng(seed + 1)
    T, G = len(tag_pos), len(target_pos)
    n = min(n, T)
    start = nearest_tag_windows(tag_pos, target_pos, n)            # [G]
    w = 3 * n
    n_out = 3 * G
    feat0 = np.repeat(3 * start, 3)                                 # [n_out] first input bigIndex
    cols = np.empty((n_out, w + 1), np.uint32)
    cols[:, 0] = CONSTANT_BIDX
    cols[:, 1:] = (feat0[:, None] + np.arange(w)[None, :]).astype(np.uint32)
    coefs = np.empty((n_out, w + 1), np.int32)
    coefs[:, 0] = rng.integers(-bias_range, bias_range, size=n_out)
    coefs[:, 1:] = rng.integers(-coef_range, coef_range, size=(n_out, w))
    keep = (coefs != 0) if drop_zeros else np.ones_like(coefs, bool)
    row_ptr = np.zeros(n_out + 1, np.uint64)
    row_ptr[1:] = np.cumsum(keep.sum(axis=1))
    return CsrModel(np.arange(n_out, dtype=np.uint32), row_ptr, cols[keep], coefs[keep])


def make_genotypes(T: int, S: int, seed: int = 1234, p=(0.7, 0.2, 0.1), na_frac: float = 0.0) -> np.ndarray:
    """[T, S] int8 genotypes in {0,1,2}, -1 = NA."""
    rng = np.random.default_rng(seed + 2)
    g = rng.choice(3, size=(T, S), p=p).astype(np.int8)
    if na_frac > 0:
        g[rng.random((T, S)) < na_frac] = -1
    return g


def write_tag_file(path, tag_pos, geno) -> None:
    """Tag (challenge) file as read by read_plaintext_data (eval/idash.cpp:280-312)."""
    with open(path, "w") as f:
        for p, row in zip(tag_pos, geno):
            toks = ["NA" if v < 0 else str(int(v)) for v in row]
            f.write(f"22\t{int(p)}\t{int(p) + 1}\trs{int(p)}\t" + "\t".join(toks) + "\n")


def write_target_file(path, target_pos) -> None:
    """Positions-only target file (keygen ... 1; eval/idash.cpp:327-343)."""
    with open(path, "w") as f:
        for p in target_pos:
            f.write(f"{int(p)}\n")


def write_hr_dir(path, model: CsrModel, tag_pos, target_pos) -> None:
    """One `<pos>_<variant>.hr` per output row; lines `Constant <v>` / `<tagpos>_<variant> <v>` with the
    value printed as a float like train/test_vw_hr.py:41-47 does."""
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    for r in range(model.n_out):
        ob = int(model.out_bidx[r])
        fn = path / f"{int(target_pos[ob // 3])}_{ob % 3}.hr"
        lines = []
        for e in range(int(model.row_ptr[r]), int(model.row_ptr[r + 1])):
            c, v = int(model.col[e]), int(model.coef[e])
            name = "Constant" if c == CONSTANT_BIDX else f"{int(tag_pos[c // 3])}_{c % 3}"
            lines.append(f"{name} {v:.1f}\n")
        with open(fn, "w") as f:
            f.writelines(lines)


def random_ciphertexts(n: int, seed: int = 1234) -> np.ndarray:
    """[n, 2048] uniformly random words: what TRLWE ciphertexts look like to the evaluator."""
    rng = np.random.default_rng(seed + 3)
    return rng.integers(0, 2 ** 32, size=(n, CT_WORDS), dtype=np.uint32)


def host_cpus() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
