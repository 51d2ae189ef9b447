import time, ctypes as C, numpy as np, os, sys
from idash2019_2_b200 import synth, _lib as L
lib = L.lib()
tag, tgt = synth.make_positions(16184, 80882, 1234)
m = synth.make_model(tag, tgt, int(sys.argv[1]), 1234)
desc, keep = L.make_desc(1004, 1, 1024, m.out_bidx, m.row_ptr, m.col, m.coef)
for _ in range(2):
    h = C.c_void_p(); t0 = time.perf_counter()
    assert lib.idash_b200_layout_compile_ex(C.byref(desc), 0, C.byref(h)) == 0
    print("total", time.perf_counter() - t0)
    lib.idash_b200_layout_free(h)
