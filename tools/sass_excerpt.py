#!/usr/bin/env python
"""SASS evidence for the tcgen05 / TMA kernels of the built library: instruction counts per kernel + sample lines.
usage: python tools/sass_excerpt.py [lib.so] > profiles/r02_sass_tcgen05.txt"""
import re
import subprocess
import sys
from collections import Counter, defaultdict

lib = sys.argv[1] if len(sys.argv) > 1 else "idash2019_2_b200/lib/libidash_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ("UTCIMMA", "UTCBAR", "LDTM", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "SYNCS", "STG", "LDG", "STS", "LDS", "SHFL", "PRMT")
cnt, ex, total = defaultdict(Counter), defaultdict(list), Counter()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(.*?);", line)
    if not (fn and m):
        continue
    total[fn] += 1
    ins = m.group(1).strip()
    op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
    base = op.split(".")[0]
    if base in KEYS:
        cnt[fn][base] += 1
        if base in ("UTCIMMA", "UTCBAR", "LDTM", "UBLKCP", "UBLKPF") and sum(1 for e in ex[fn] if e.split()[0].startswith(base) or (e.startswith("@") and e.split()[1].startswith(base))) < 2:
            ex[fn].append(ins)
print(f"# cuobjdump -sass {lib} (sm_100a)")
print("# UTCIMMA = tcgen05.mma kind::i8 (.WS = weight-stationary), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld (.PACK16BIT = pack::16b),")
print("# UBLKCP = cp.async.bulk (TMA bulk copy; .S.G global -> shared, .G.S shared -> global: the zero fill), UBLKPF = cp.async.bulk.prefetch.L2, SYNCS = mbarrier,")
print("# UTMALDG = cp.async.bulk.tensor (tensor-map TMA: the decrypt kernel's b words). The cloud kernels' ciphertext words are staged through registers (byte-plane split), DESIGN.md 3.2")
for f in sorted(cnt):
    if not any(cnt[f][k] for k in ("UTCIMMA", "LDTM", "UBLKCP")):
        continue
    print(f"\n{f}: {total[f]} instructions")
    print("   " + "  ".join(f"{k} {cnt[f][k]}" for k in KEYS if cnt[f][k]))
    for e in ex[f]:
        print("      " + e)
