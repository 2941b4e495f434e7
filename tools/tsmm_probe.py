import sys, ctypes as C
import numpy as np
sys.path.insert(0, "/root/repo")
from rbnics_b200 import offline
from rbnics_b200.engine import SweepEngine
eng = SweepEngine(0)
rng = np.random.default_rng(0)
Nh = 1000000
for (a, b) in [(61, 61), (20, 20), (20, 1), (64, 64)]:
    S = np.ascontiguousarray(rng.standard_normal((Nh, a))); S2 = np.ascontiguousarray(rng.standard_normal((Nh, b)))
    for _ in range(3):
        Cm = offline.gram(eng, S, None, S2)
        ms = (C.c_float * 5)(); n = C.c_int(); eng._lib.rb_last_timings(eng._h, ms, C.byref(n))
    ref = S[:20000].T @ S2[:20000]
    print(a, b, "tsmm ms %.3f" % ms[1], flush=True)
