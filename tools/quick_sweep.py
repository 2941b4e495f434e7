import time, numpy as np, sys
sys.path.insert(0, "/root/repo")
from rbnics_b200.engine import SweepEngine
from rbnics_b200 import synthetic
eng = SweepEngine(0)
N, Qa, Qf = 100, 50, 10
prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=0, riesz_rank=512)
t0=time.time(); eng.set_system(prob["A"], prob["F"]); eng.set_elliptic_estimator(prob["Gff"], prob["Gaf"], prob["Gaa"]); print("set ops", time.time()-t0, flush=True)
for B in (18944, 148*128*4, 200000):
    Theta, beta, _ = synthetic.training_set(B, Qa, Qf, seed=1)
    eng.set_training_set(Theta, beta)
    for it in range(2):
        t0=time.time(); r = eng.sweep_resident(); dt=time.time()-t0
        print(B, it, dt, B/dt, r["max"], r["argmax"], eng.last_timings(), flush=True)
