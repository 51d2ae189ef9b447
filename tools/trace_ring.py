#!/usr/bin/env python
"""Prints the per-tile timeline recorded by IDASH_B200_TRACE (cycles relative to the first event).
events: 0 mma iteration start, 1 after t_empty, 2 after a_full, 3 after b_full+fence, 4 after MMA/commit issue,
        5 epilogue before t_full wait, 6 after t_full, 7 epilogue done"""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1], dtype=np.uint64).astype(np.int64)
t0 = a[a > 0].min()
r = a - t0
names = ["it0", "tEmp", "aFull", "bFull", "issued", "e_wait", "e_go", "e_done"]
print("tile " + " ".join(f"{n:>8}" for n in names) + "   | mma_iter  epi_busy  epi_wait")
for i in range(1, len(r) - 1):
    row = r[i]
    print(f"{i:4d} " + " ".join(f"{v:8d}" for v in row) + f"   | {r[i+1][0]-row[0]:8d} {row[7]-row[6]:8d} {row[6]-row[5]:8d}")
d = np.diff(r[:, 0])
print("mean cycles per tile (mma loop):", d[5:-5].mean(), " epilogue busy mean:", (r[:, 7] - r[:, 6])[5:-5].mean(),
      " issue phase mean:", (r[:, 4] - r[:, 3])[5:-5].mean())
