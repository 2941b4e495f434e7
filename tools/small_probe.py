import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from rbnics_b200.engine import SweepEngine
from rbnics_b200 import synthetic
eng = SweepEngine(0)
Qa, Qf, B = 9, 3, 100000
Theta, beta, _ = synthetic.training_set(B, Qa, Qf, seed=1)
eng.set_training_set(Theta, beta)
for N in [1, 2, 4, 6, 6, 12, 20]:
    prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=0, riesz_rank=64)
    t0 = time.time(); eng.set_system(prob["A"], prob["F"]); t1 = time.time()
    eng.set_elliptic_estimator(prob["Gff"], prob["Gaf"], prob["Gaa"]); t2 = time.time()
    for it in range(3):
        t3 = time.time(); r = eng.sweep_resident(); dt = time.time() - t3
        tm = eng.last_timings()
        print(f"N={N} it={it} set_system {1e3*(t1-t0):.2f} ms set_est {1e3*(t2-t1):.2f} ms sweep wall {1e3*dt:.2f} ms asm+lu {tm['assemble_lu_ms']:.3f} est {tm['estimator_ms']:.3f} launches {tm['launches']}", flush=True)
