#!/usr/bin/env python
"""bench_greedy.py -- an elliptic RB greedy offline stage at scale (BASELINE.json configs[1]: a coercive block problem with
the training set scaled to 10^5 parameters), through rbnics_b200.greedy.ReducedBasisGreedy with the device-resident
workspace: N_h = 10^6 dofs, Q_a = 9 blocks, Q_f = 3 loads (the shapes of tutorials/02_elastic_block), 11 parameters.

Prints one JSON line: wall seconds per stage summed over the greedy iterations.  The truth solves and the Riesz solves
are the reference's input generators (host, sparse LU); everything between them and the selected parameter -- the
Gram-Schmidt step, the new rows / columns of the reduced operators, the Riesz Gram rows, the sweep of 10^5 parameters
-- runs on the GPU on resident data."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=1000, help="grid side: N_h = side^2")
    ap.add_argument("--nblocks", type=int, default=9)
    ap.add_argument("--nloads", type=int, default=3)
    ap.add_argument("--nmax", type=int, default=6)
    ap.add_argument("--ntrain", type=int, default=100_000)
    args = ap.parse_args()
    from rbnics_b200.engine import SweepEngine
    from rbnics_b200.greedy import ReducedBasisGreedy
    from truth_problems import ThermalBlock
    t0 = time.perf_counter()
    tp = ThermalBlock(n=args.side, nblocks=args.nblocks, nloads=args.nloads,
                      mu_range=[(1.0, 100.0)] * (args.nblocks - 1) + [(-1.0, 1.0)] * args.nloads)
    setup = time.perf_counter() - t0
    train = tp.sample_training_set(args.ntrain, seed=0)
    eng = SweepEngine(0)
    # CUDA loads a kernel lazily at its first launch (tens to hundreds of ms for each new template instantiation of
    # the small-N LU): touch every reduced size of this run once on a dummy problem, so that the per-iteration sweep
    # times below are the kernels' and not the module loader's
    import numpy as np
    from rbnics_b200 import synthetic
    for n_warm in range(1, args.nmax + 2):
        prob = synthetic.elliptic_reduced_problem(n_warm, args.nblocks, args.nloads, seed=0, riesz_rank=16)
        Th, be, _ = synthetic.training_set(512, args.nblocks, args.nloads, seed=1)
        eng.set_system(prob["A"], prob["F"])
        eng.set_elliptic_estimator(prob["Gff"], prob["Gaf"], prob["Gaa"])
        eng.sweep(Th, be)
    t0 = time.perf_counter()
    run = ReducedBasisGreedy(eng, tp, train, Nmax=args.nmax, tol=0.0, resident=True)
    init = time.perf_counter() - t0
    t0 = time.perf_counter()
    run.offline()
    wall = time.perf_counter() - t0
    T = dict(run.timings)
    iters = max(len(run.greedy_selected_indices) - 1, 1)
    print(json.dumps({
        "bench": "elliptic RB greedy offline stage, shapes of BASELINE configs[1] (Q_a = 9, Q_f = 3, 11 parameters)",
        "Nh": tp.Nh, "Qa": args.nblocks, "Qf": args.nloads, "ntrain": args.ntrain, "N": run.N,
        "selected": run.greedy_selected_indices, "error_estimators": run.greedy_error_estimators,
        "wall_s_truth_problem_setup": setup, "wall_s_workspace_upload": init, "wall_s_offline": wall,
        "wall_s_by_stage_summed_over_iterations": T,
        "per_iteration_ms": {k: 1e3 * v / (iters + (k == "gpu_sweep")) for k, v in T.items()},
        "sweep_params_per_s": args.ntrain * (iters + 1) / max(T.get("gpu_sweep", 0.0), 1e-9),
        "note": "riesz_solves_and_gram = Q_a host Riesz solves (sparse LU) per iteration + their upload + the Gram rows "
                "on the GPU; host_truth_solve = one sparse LU factorisation and solve per iteration",
    }), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
