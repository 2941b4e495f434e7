"""Dof-row partitioned Gram matrix over several GPUs (SURVEY.md 8e, row "Gram / projections"): every rank computes the
partial S_g^T (X S)_g of its PETSc-style ownership range on its own GPU, one NCCL all-reduce of the Ns x Ns block adds
them up.  Checks the result against the single-GPU product and prints one JSON line per shape (rank 0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_gram_check.py [--nh 1000000] [--ns 61 400]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rbnics_b200 import offline  # noqa: E402
from rbnics_b200.engine import SweepEngine  # noqa: E402
from rbnics_b200.sweep import DistributedMaxLoc  # noqa: E402


def kernel_ms(eng):
    ms = (C.c_float * 5)()
    n = C.c_int()
    eng._lib.rb_last_timings(eng._h, ms, C.byref(n))
    return float(ms[4])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nh", type=int, default=1_000_000)
    ap.add_argument("--ns", type=int, nargs="+", default=[61, 400])
    args = ap.parse_args()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = DistributedMaxLoc(local_rank)
    eng = SweepEngine(local_rank)
    side = int(round(args.nh ** 0.5))
    Nh = side * side
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(side, side))
    I = sp.identity(side)
    X = (sp.kron(T, I) + sp.kron(I, T) + 0.5 * sp.identity(Nh)).tocsr()
    for Ns in args.ns:
        S = np.ascontiguousarray(np.random.default_rng(Ns).standard_normal((Nh, Ns)))
        best = None
        for _ in range(3):
            dist.barrier()
            Cd = offline.gram(eng, S, X, dist=dist)
            ms = dist.max_float(kernel_ms(eng))
            best = ms if best is None else min(best, ms)
        Cf = offline.gram(eng, S, X)
        ms_full = kernel_ms(eng)
        err = float(np.max(np.abs(Cd - Cf)) / np.max(np.abs(Cf)))
        err = dist.max_float(err)
        if dist.rank == 0:
            print(json.dumps({"check": "row-partitioned Gram + all-reduce", "n_gpus": dist.size, "Nh": Nh, "Ns": Ns,
                              "kernel_ms_max_over_ranks": best, "kernel_ms_single_gpu": ms_full,
                              "max_rel_diff_vs_single_gpu": err, "rows_per_rank": offline.row_range(Nh, 0, dist.size)[1]}),
                  flush=True)
        assert err <= 1e-12, err
        del S
    eng.close()
    dist.close()


if __name__ == "__main__":
    main()
