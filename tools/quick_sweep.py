"""Quick device-resident sweep timing at the bench shapes (N=100, Qa=50, Qf=10): python tools/quick_sweep.py [B ...]"""
import sys, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from rbnics_b200.engine import SweepEngine
from rbnics_b200 import synthetic
eng = SweepEngine(0)
N, Qa, Qf = 100, 50, 10
prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=0, riesz_rank=512)
eng.set_system(prob["A"], prob["F"]); eng.set_elliptic_estimator(prob["Gff"], prob["Gaf"], prob["Gaa"])
for B in [int(x) for x in sys.argv[1:]] or [200000]:
    Theta, beta, _ = synthetic.training_set(B, Qa, Qf, seed=1)
    eng.set_training_set(Theta, beta)
    for it in range(2):
        t0 = time.time(); r = eng.sweep_resident(); dt = time.time() - t0
    tm = eng.last_timings()
    print(f"B={B} {B/dt:.0f} params/s  asm+lu {tm['assemble_lu_ms']:.2f} ms  est {tm['estimator_ms']:.2f} ms  launches {tm['launches']}  max {r['max']:.12g} argmax {r['argmax']}", flush=True)
