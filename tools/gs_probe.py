"""Resident Gram-Schmidt step timing: python tools/gs_probe.py [N] [Nh]"""
import sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from rbnics_b200 import offline
from rbnics_b200.engine import SweepEngine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
Nh = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
side = int(round(Nh ** 0.5)); Nh = side * side
T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(side, side)); I = sp.identity(side)
X = (sp.kron(T, I) + sp.kron(I, T) + 0.5 * sp.identity(Nh)).tocsr()
eng = SweepEngine(0)
rng = np.random.default_rng(0)
ws = offline.OfflineWorkspace(eng)
ws.set_operator("X", X)
ws.new_block("Z", Nh, N + 4)
ws.append("Z", np.ascontiguousarray(rng.standard_normal((Nh, N)) / np.sqrt(Nh)))
z = rng.standard_normal(Nh)
for it in range(4):
    t0 = time.perf_counter()
    out, nrm = ws.gram_schmidt("Z", "X", z, append=(it == 3))
    print("N", N, "GS step wall %.2f ms  norm %.6g" % ((time.perf_counter() - t0) * 1e3, nrm), flush=True)
