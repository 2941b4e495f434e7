#!/usr/bin/env python
"""bench_pod_greedy.py -- BASELINE.json configs[3] end to end: tutorials/06_thermal_block_unsteady, POD-Greedy, at
N_h = 10^6 dofs (K = 60 time steps: 61 trajectory snapshots per iteration, Gram S^T X S, projections Z^T A_q Z, Riesz
Gram blocks, sweep over the training set), through rbnics_b200.pod_greedy.PODGreedyReducedBasis.

Prints one JSON line: wall seconds per stage summed over the greedy iterations, and next to the GPU stages the time of
the same products done with scipy.sparse + numpy (all host cores) on the data of the last iteration.  The truth
trajectories and the Riesz solves are the reference's input generators (host, sparse LU) and are reported separately."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=1000, help="grid side: N_h = side^2")
    ap.add_argument("--iterations", type=int, default=3)
    ap.add_argument("--n1", type=int, default=4)
    ap.add_argument("--ntrain", type=int, default=100_000)
    ap.add_argument("--k", type=int, default=60)
    ap.add_argument("--host-operands", action="store_true",
                    help="upload the operands of every product on every call instead of the resident workspace")
    args = ap.parse_args()
    from oracle import rb_oracle as O
    from rbnics_b200.engine import SweepEngine
    from rbnics_b200.pod_greedy import PODGreedyReducedBasis
    from truth_problems import ThermalBlockUnsteady
    t0 = time.perf_counter()
    tp = ThermalBlockUnsteady(n=args.side, nblocks=2, nloads=1, dt=0.05, K=args.k)
    setup = time.perf_counter() - t0
    train = tp.sample_training_set(args.ntrain, seed=0)
    eng = SweepEngine(0)
    Nmax = args.iterations * args.n1
    t0 = time.perf_counter()
    run = PODGreedyReducedBasis(eng, tp, train, Nmax=Nmax, tol=0.0, N1=args.n1, resident=not args.host_operands).offline()
    wall = time.perf_counter() - t0
    T = dict(run.timings)
    # the same products on the host cores, once, on the final data (numpy / scipy.sparse, all BLAS threads)
    Z, X = run.basis_functions, tp.inner_product
    S = np.ascontiguousarray(tp.solve_over_time(run.greedy_selected_parameters[-1]).T)
    cpu = {}
    t0 = time.perf_counter()
    c = np.linalg.solve(O.gram_blas(Z, X), Z.T @ (X @ S))
    So = S - Z @ c
    cpu["orthogonalize"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.pod(So, X, args.n1, 0.0)
    cpu["trajectory_pod"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    for A in tp.operators_m + tp.operators_a:
        O.gram_blas(Z, A)
    cpu["projections"] = time.perf_counter() - t0
    R = np.concatenate(run.riesz_a + run.riesz_m + [run.riesz_f], axis=1)
    t0 = time.perf_counter()
    O.gram_blas(R, X)
    cpu["riesz_gram"] = time.perf_counter() - t0
    gpu_stages = {k: v for k, v in T.items() if k.startswith("gpu_")}
    print(json.dumps({
        "bench": "POD-Greedy offline loop, tutorials/06 shapes (BASELINE configs[3])",
        "Nh": tp.Nh, "K": args.k, "N1": args.n1, "iterations": len(run.greedy_selected_indices) - 1, "N": run.N,
        "ntrain": args.ntrain, "selected": run.greedy_selected_indices,
        "wall_s_total": wall, "wall_s_truth_problem_setup": setup,
        "wall_s_by_stage_summed_over_iterations": T,
        "wall_s_gpu_stages": sum(gpu_stages.values()),
        "wall_s_host_input_generators": T.get("host_truth_trajectory", 0.0) + T.get("host_riesz_solves", 0.0),
        "cpu_numpy_scipy_last_iteration_s": cpu,
        "resident": not args.host_operands,
        "note": "gpu_* stages are wall times summed over the iterations (divide by the iteration count for a "
                "per-iteration figure); resident: a trajectory (488 MB) crosses PCIe once per iteration, the N1 modes "
                "and the Riesz representers once; --host-operands: every product uploads its operands",
        "sweep_params_per_s_last": args.ntrain / max(T.get("gpu_sweep", 0.0) / max(len(run.greedy_selected_indices), 1), 1e-9),
    }), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
