"""Multi-GPU offline path on real GPUs: dof rows partitioned over the ranks, partial
Gram / projection blocks and the Gram-Schmidt coefficients summed over the ranks IN HBM by NCCL (rb_set_allreduce), and
a whole resident greedy run on top.  Every result is checked against the single-GPU computation of the same inputs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tools/dist_offline_check.py [--nh 250000]
Prints one JSON line per check (rank 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rbnics_b200 import offline  # noqa: E402
from rbnics_b200.engine import SweepEngine  # noqa: E402
from rbnics_b200.greedy import ReducedBasisGreedy  # noqa: E402
from rbnics_b200.sweep import DistributedMaxLoc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nh", type=int, default=250_000)
    ap.add_argument("--bench", action="store_true", help="also time the partitioned products at N_h = 10^6")
    args = ap.parse_args()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = DistributedMaxLoc(local_rank)
    eng = SweepEngine(local_rank)
    side = int(round(args.nh ** 0.5))
    Nh = side * side
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(side, side))
    I = sp.identity(side)
    X = (sp.kron(T, I) + sp.kron(I, T) + 0.5 * sp.identity(Nh)).tocsr()
    rng = np.random.default_rng(3)

    def report(name, **kw):
        if dist.rank == 0:
            print(json.dumps(dict(check=name, n_gpus=dist.size, Nh=Nh, **kw)), flush=True)

    # ---- Gram on resident blocks, partitioned, all-reduce on the device buffer
    S = np.ascontiguousarray(rng.standard_normal((Nh, 40)))
    ws_d = offline.OfflineWorkspace(eng, dist=dist)
    ws_d.set_operator("X", X)
    ws_d.new_block("S", Nh, 64)
    ws_d.append("S", S)
    Cd = ws_d.gram("S", "X")
    C1 = offline.gram(eng, S, X)                      # single-GPU product of the same inputs (hook switched off)
    err = dist.max_float(float(np.max(np.abs(Cd - C1)) / np.max(np.abs(C1))))
    report("row-partitioned resident Gram, NCCL all-reduce in HBM", Ns=40, max_rel_diff_vs_single_gpu=err)
    # several CTA tiles: symmetric tiles mirrored at warp-tile granularity per rank, THEN summed over the ranks
    S2 = np.ascontiguousarray(rng.standard_normal((Nh, 200)))
    ws_d.new_block("S200", Nh, 200)
    ws_d.append("S200", S2)
    Cd = ws_d.gram("S200", "X")
    C1 = offline.gram(eng, S2, X)
    err = dist.max_float(float(np.max(np.abs(Cd - C1)) / np.max(np.abs(C1))))
    sym = dist.max_float(float(np.max(np.abs(Cd - Cd.T)) / np.max(np.abs(Cd))))
    report("row-partitioned resident Gram, several tiles (tensor-map TMA, mirrored upper triangle)", Ns=200,
           max_rel_diff_vs_single_gpu=err, max_rel_asymmetry=sym)
    assert err <= 1e-12, err

    # ---- Gram-Schmidt: 24 vectors orthonormalised one after the other, partitioned vs single GPU
    ws_d.new_block("Z", Nh, 32)
    ws_1 = offline.OfflineWorkspace(eng, dist=None)    # slots are per engine: use different names
    ws_1._csr, ws_1._blk = dict(ws_d._csr), dict(ws_d._blk)
    ws_1.Nh = Nh
    ws_1.new_block("Z1", Nh, 32)
    worst, t_d, t_1 = 0.0, 0.0, 0.0
    for k in range(24):
        v = np.ascontiguousarray(rng.standard_normal(Nh) + (S[:, k] if k < 40 else 0.0))
        dist.barrier()
        t0 = time.perf_counter()
        zd, nd = ws_d.gram_schmidt("Z", "X", v)
        t_d += time.perf_counter() - t0
        t0 = time.perf_counter()
        z1, n1 = ws_1.gram_schmidt("Z1", "X", v)
        t_1 += time.perf_counter() - t0
        worst = max(worst, float(np.max(np.abs(zd - z1)) / np.max(np.abs(z1))), abs(nd - n1) / n1)
    worst = dist.max_float(worst)
    Zd = ws_d.read("Z")
    G = ws_d.gram("Z", "X")
    orth = dist.max_float(float(np.max(np.abs(G - np.eye(24)))))
    report("row-partitioned Gram-Schmidt (MGS order kept, one all-reduce per projection)", N=24,
           max_rel_diff_vs_single_gpu=worst, max_abs_deviation_from_orthonormal=orth,
           wall_ms_per_step_partitioned=1e3 * t_d / 24, wall_ms_per_step_single=1e3 * t_1 / 24)
    assert worst <= 1e-10 and orth <= 1e-10, (worst, orth)

    # ---- timing at N_h = 10^6 (kernel ms from CUDA events on the engine's stream, max over ranks; the all-reduce is outside
    #      the events, its wall time is reported separately)
    if args.bench:
        import ctypes as C

        def kernel_ms():
            ms = (C.c_float * 5)()
            n = C.c_int()
            eng._lib.rb_last_timings(eng._h, ms, C.byref(n))
            return float(ms[4])

        side2 = 1000
        Nb = side2 * side2
        T2 = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(side2, side2))
        I2 = sp.identity(side2)
        Xb = (sp.kron(T2, I2) + sp.kron(I2, T2) + 0.5 * sp.identity(Nb)).tocsr()
        wsb = offline.OfflineWorkspace(eng, dist=dist)
        wsb.set_operator("X", Xb)
        for Ns in (61, 400):
            Sb = np.ascontiguousarray(np.random.default_rng(Ns).standard_normal((Nb, Ns)))
            wsb.new_block("Sb", Nb, Ns)
            wsb.append("Sb", Sb)
            best_k, best_w = None, None
            for _ in range(3):
                dist.barrier()
                t0 = time.perf_counter()
                wsb.gram("Sb", "X")
                w = dist.max_float(time.perf_counter() - t0)
                k = dist.max_float(kernel_ms())
                best_k = k if best_k is None else min(best_k, k)
                best_w = w if best_w is None else min(best_w, w)
            report("bench: row-partitioned resident Gram S^T X S", Nh_bench=Nb, Ns=Ns, kernel_ms_max_over_ranks=best_k,
                   wall_ms_with_allreduce_and_d2h=1e3 * best_w)
            del Sb
        wsb.new_block("Zb", Nb, 101)
        wsb.append("Zb", np.ascontiguousarray(np.random.default_rng(1).standard_normal((Nb, 100))))
        zb = np.random.default_rng(2).standard_normal(Nb)
        wsb.gram_schmidt("Zb", "X", zb, append=False)
        best = None
        for _ in range(3):
            dist.barrier()
            t0 = time.perf_counter()
            wsb.gram_schmidt("Zb", "X", zb, append=False)
            w = dist.max_float(time.perf_counter() - t0)
            best = w if best is None else min(best, w)
        report("bench: row-partitioned Gram-Schmidt step, N = 100 (100 all-reduces of 4.7 KB)", Nh_bench=Nb,
               wall_ms=1e3 * best)

    # ---- a whole greedy run: sharded sweep + partitioned resident offline stage == single-GPU run
    from truth_problems import ThermalBlock
    tp = ThermalBlock(n=40, nblocks=3, nloads=2)
    train = tp.sample_training_set(4000, seed=7)
    run_d = ReducedBasisGreedy(eng, tp, train, Nmax=8, tol=1e-7, dist=dist, resident=True).offline()
    eng1 = SweepEngine(local_rank)
    run_1 = ReducedBasisGreedy(eng1, tp, train, Nmax=8, tol=1e-7, resident=True).offline()
    same = run_d.greedy_selected_indices == run_1.greedy_selected_indices
    dA = float(np.max(np.abs(run_d.operator_a - run_1.operator_a)) / np.max(np.abs(run_1.operator_a)))
    report("resident greedy: training set sharded + dof rows partitioned", ntrain=4000, N=int(run_d.N),
           identical_greedy_sequence=bool(same), selected=run_d.greedy_selected_indices,
           max_rel_diff_reduced_operators=dist.max_float(dA))
    assert same and dA <= 1e-9
    # ---- POD-Greedy (BASELINE configs[3]): training set of the parabolic sweep sharded over the ranks
    from rbnics_b200.pod_greedy import PODGreedyReducedBasis
    from truth_problems import ThermalBlockUnsteady
    tpu = ThermalBlockUnsteady(n=16, nblocks=2, nloads=1, dt=0.05, K=20)
    trainu = tpu.sample_training_set(2001, seed=5)
    pg_d = PODGreedyReducedBasis(eng, tpu, trainu, Nmax=6, tol=1e-8, N1=2, dist=dist).offline()
    pg_1 = PODGreedyReducedBasis(eng1, tpu, trainu, Nmax=6, tol=1e-8, N1=2).offline()
    same_pg = pg_d.greedy_selected_indices == pg_1.greedy_selected_indices
    de = float(np.max(np.abs(np.array(pg_d.greedy_error_estimators) - np.array(pg_1.greedy_error_estimators))
                      / pg_1.greedy_error_estimators[0]))
    report("POD-Greedy: parabolic sweep sharded over the ranks", ntrain=2001, N=int(pg_d.N),
           identical_greedy_sequence=bool(same_pg), selected=pg_d.greedy_selected_indices,
           max_diff_estimators_rel_to_first=dist.max_float(de))
    assert same_pg and de <= 1e-12
    eng1.close()
    eng.close()
    dist.close()


if __name__ == "__main__":
    main()
