"""Per-phase cycle breakdown of rb_lu_mw_kernel.  Needs the instrumented build (not part of the product library):

    cd rbnics_b200/csrc && make && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC \
        -DRB_MW_PROFILE -DRB_AB_VARIANTS -c rb_lu_mw.cu -o ../../build/csrc/rb_lu_mw_prof.o && \
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../librbnics_b200_prof.so ../../build/csrc/{rb_sweep,rb_online,rb_lu,rb_lu_inst0,rb_lu_mw_prof,rb_offline}.o -lcudart

    python tools/lu_phase_probe.py [systems per SM ...]"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rbnics_b200 import _lib as L
L.LIB_PATH = os.path.join(ROOT, "rbnics_b200", "librbnics_b200_prof.so")
from rbnics_b200.engine import SweepEngine
from rbnics_b200 import synthetic
NAMES = ["loop head / rhs", "barrier (panel top)", "wait A panel", "tiles <- staged A", "wait record", "dump",
         "barrier A", "U_qp", "barrier B", "tile updates", "transpose to rows", "8 pivot steps", "publish record",
         "proxy fence", "back-substitution", "-"]
B = 18648 * 2
N, Qa, Qf = 100, 50, 10
prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=0, riesz_rank=512)
Theta, beta, _ = synthetic.training_set(B, Qa, Qf, seed=1)
for occ in [int(x) for x in sys.argv[1:]] or [9, 1]:
    os.environ["RB_MW_OCC"] = str(occ)
    # the occupancy knob is read once per process: one subprocess per value
    if os.environ.get("_PROBE_CHILD") != "1":
        import subprocess
        env = dict(os.environ, _PROBE_CHILD="1", RB_MW_OCC=str(occ))
        subprocess.run([sys.executable, __file__, str(occ)], env=env, check=True)
        continue
    eng = SweepEngine(0)
    eng.set_system(prob["A"], prob["F"])
    eng.set_elliptic_estimator(prob["Gff"], prob["Gaf"], prob["Gaa"])
    eng.set_training_set(Theta, beta)
    lib = eng._lib
    buf = (C.c_ulonglong * 16)()
    eng.sweep_resident()
    lib.rb_mw_profile_read(buf, 1)
    eng.sweep_resident()
    lib.rb_mw_profile_read(buf, 1)
    tm = eng.last_timings()
    tot = sum(buf)
    print(f"occupancy {occ}/SM: asm+lu {tm['assemble_lu_ms']:.2f} ms for {B} systems; cycles per system (thread 0 of each CTA):")
    for n, v in zip(NAMES, buf):
        if v:
            print(f"   {n:22s} {v / B:9.0f}  {100 * v / tot:5.1f}%")
    print(f"   {'total':22s} {tot / B:9.0f}")
    eng.close()
