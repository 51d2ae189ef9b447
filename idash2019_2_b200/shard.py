"""Multi-GPU partition of the cloud evaluation: contiguous target-SNP ranges, one per rank.

Every output ciphertext depends only on its own <= 3n+1 model entries (eval/idash.cpp:800-819), so the path shards
by independent units with NO collective on the data path: rank r of P evaluates the targets [G r / P, G (r+1) / P) and
needs exactly the input ciphertexts its rows touch -- a contiguous slab [ct_min, ct_max] of the tag-ciphertext array
(the model is banded in genomic order) that is derived from the model itself, not from `neighbors`. The shard's model
is re-based so that slab slot 0 is ciphertext ct_min; a feature f = ct * NR + region keeps its region.

Host logic only (numpy); the ranks run idash_b200 contexts (one process per GPU) on their shard.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .synth import CsrModel

CONSTANT_BIDX = 0xFFFFFFFF


def target_range(n_targets: int, rank: int, world: int) -> tuple[int, int]:
    """Targets [lo, hi) of `rank`: contiguous, disjoint, covering, sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return n_targets * rank // world, n_targets * (rank + 1) // world


@dataclass
class Shard:
    rank: int
    world: int
    target_lo: int          # targets [target_lo, target_hi) of the global problem
    target_hi: int
    row_lo: int             # rows [row_lo, row_hi) of the global CSR (3 consecutive rows per target)
    row_hi: int
    ct_min: int             # first input ciphertext of the slab (global index)
    n_ct: int               # slab length in ciphertexts (0 if the shard has no non-constant entry)
    model: CsrModel         # rows of this shard; col re-based to the slab, out_bidx unchanged (global)

    def slab(self, cts: np.ndarray) -> np.ndarray:
        """The rows of a global [n_ct_total, 2048] ciphertext array this shard reads."""
        return cts[self.ct_min:self.ct_min + self.n_ct]


def make_shard(model: CsrModel, num_regions: int, n_targets: int, rank: int, world: int) -> Shard:
    """`model` has 3 rows per target in target order (out_bidx = 3 * target + variant, eval/idash.cpp:337-339)."""
    if model.n_out != 3 * n_targets:
        raise ValueError("model must hold 3 rows per target SNP")
    lo, hi = target_range(n_targets, rank, world)
    sub = model.rows(3 * lo, 3 * hi)
    real = sub.col != CONSTANT_BIDX
    if real.any():
        ct = sub.col[real] // np.uint32(num_regions)
        ct_min, ct_max = int(ct.min()), int(ct.max())
        col = sub.col.copy()
        col[real] -= np.uint32(ct_min * num_regions)
        sub = CsrModel(sub.out_bidx, sub.row_ptr, col, sub.coef)
        n_ct = ct_max - ct_min + 1
    else:
        ct_min, n_ct = 0, 0
    return Shard(rank, world, lo, hi, 3 * lo, 3 * hi, ct_min, n_ct, sub)


def halo(shards: list[Shard]) -> list[int]:
    """Input ciphertexts each shard shares with its left neighbour (the overlap of consecutive slabs)."""
    out = [0]
    for a, b in zip(shards, shards[1:]):
        out.append(max(0, a.ct_min + a.n_ct - b.ct_min) if a.n_ct and b.n_ct else 0)
    return out
