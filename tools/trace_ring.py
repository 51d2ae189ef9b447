#!/usr/bin/env python
"""Prints the per-tile timeline recorded by IDASH_B200_TRACE (cycles relative to the first event).
events: 0 mma iteration start, 1 t_empty seen (lane 0), 2 b_full seen (lane 1), 3 after all waits + fence,
        4 after MMA/commit issue, 5 first new a_full seen (lane 2; blank if none), 6 epilogue after t_full, 7 epilogue done"""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1], dtype=np.uint64).astype(np.int64)
t0 = a[a > 0].min()
r = a - t0
names = ["it0", "lane0ok", "synced", "fenced", "issued", "b_copy", "e_go", "e_done"]
r[a == 0] = -1
print("tile " + " ".join(f"{n:>8}" for n in names) + "   | mma_iter  epi_busy  epi_wait")
for i in range(1, len(r) - 1):
    row = r[i]
    print(f"{i:4d} " + " ".join(f"{v:8d}" for v in row) + f"   | {r[i+1][0]-row[0]:8d} {row[7]-row[6]:8d} {r[i+1][6]-row[7]:8d}")
d = np.diff(r[:, 0])
print("mean cycles per tile (mma loop):", d[5:-5].mean(), " epilogue busy mean:", (r[:, 7] - r[:, 6])[5:-5].mean(),
      " issue phase mean:", (r[:, 4] - r[:, 3])[5:-5].mean())
