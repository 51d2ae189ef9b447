#!/usr/bin/env python
"""Prints the per-tile timeline recorded by IDASH_B200_TRACE (SM cycles relative to the first event).
MMA warp: it0 loop top, te/bf/af = time lane 0 / 1 / 2 saw t_empty / b_full / first new a_full complete (-1: nothing to wait
for), waits = all waits done, issued = after the tile's MMAs and commit were issued. Epilogue warp 0: e_go = t_full seen,
e_done = tile stored. pub = publisher saw t_full. b_copy = coefficient loader issued the tile's bulk copy. mma0 / mmaN = just before
the first / after the last tcgen05.mma of the tile was issued (issued = after the commit as well)."""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1], dtype=np.uint64).astype(np.int64)
t0 = a[a > 0].min()
r = a - t0
r[a == 0] = -1
names = ["it0", "te", "bf", "af", "waits", "issued", "e_go", "mma0", "e_done", "pub", "b_copy", "mmaN"]
cols = [0, 1, 2, 3, 4, 7, 11, 5, 6, 8, 9, 10]
print("tile " + " ".join(f"{names[c]:>8}" for c in cols) + "   | mma_iter  epi_busy  epi_wait commit->e_go")
for i in range(1, len(r) - 1):
    row = r[i]
    print(f"{i:4d} " + " ".join(f"{row[c]:8d}" for c in cols) + f"   | {r[i+1][0]-row[0]:8d} {row[8]-row[6]:8d} {r[i+1][6]-row[8]:8d} {row[6]-row[5]:8d}")
d = np.diff(r[:, 0])
print("mean cycles per tile (mma loop):", d[5:-5].mean(), " epilogue busy mean:", (r[:, 8] - r[:, 6])[5:-5].mean(),
      " issue phase mean:", (r[:, 5] - r[:, 4])[5:-5].mean(), " commit->e_go mean:", (r[:, 6] - r[:, 5])[5:-5].mean())
