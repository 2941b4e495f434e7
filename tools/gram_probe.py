"""One fused / unfused Gram at bench_offline shapes, for ncu: python tools/gram_probe.py [Ns] [Nh]"""
import sys
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from rbnics_b200 import offline
from rbnics_b200.engine import SweepEngine
import ctypes as C

Ns = int(sys.argv[1]) if len(sys.argv) > 1 else 61
Nh = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
side = int(round(Nh ** 0.5))
T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(side, side))
I = sp.identity(side)
X = (sp.kron(T, I) + sp.kron(I, T) + 0.5 * sp.identity(side * side)).tocsr()
eng = SweepEngine(0)
S = np.random.default_rng(0).standard_normal((side * side, Ns))
for _ in range(3):
    offline.gram(eng, S, X)
    ms = (C.c_float * 5)()
    n = C.c_int()
    eng._lib.rb_last_timings(eng._h, ms, C.byref(n))
    print("Ns", Ns, "spmm %.3f ms  gram %.3f ms" % (ms[0], ms[1]), flush=True)
