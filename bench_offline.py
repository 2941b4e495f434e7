#!/usr/bin/env python
"""bench_offline.py -- the offline basis linear algebra at BASELINE.json configs[3] shapes (tutorial 06 scaled to
10^6 dofs): POD Gram matrix S^T X S, Galerkin projections Z^T A_q Z, mode construction S V and one Gram-Schmidt
step, each with its kernel time (CUDA events on the engine's stream, rb_last_timings) and the roofline it is bound by
(SURVEY.md 8d):  SpMM  12*nnz + 16*Nh*Ns bytes (HBM);  tall-skinny Gram 2*Nh*Ns*Ns2 flops / 8*Nh*(Ns+Ns2) bytes.

    python bench_offline.py [--nh 1000000] [--ns 61 400] [--n 20 100]

Prints one JSON line per kernel.  Not the headline metric (bench.py is); this is the measurement of rows B1-B5."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def lap2d(n):
    """7-point-like P1 stiffness + mass on an n x n grid (5 nnz per row), the synthetic X of SURVEY.md 8d config 4."""
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    I = sp.identity(n)
    return (sp.kron(T, I) + sp.kron(I, T) + 0.5 * sp.identity(n * n)).tocsr()


def peaks():
    hbm, fp64 = 6545.9, 37.15
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    try:
        fp64 = float(json.load(open(os.path.join(ROOT, "profiles", "FP64_PEAK.json")))["fp64_dmma_tflops"])
    except Exception:
        pass
    return hbm, fp64


def timings(eng):
    ms = (C.c_float * 5)()
    n = C.c_int()
    eng._lib.rb_last_timings(eng._h, ms, C.byref(n))
    return list(ms), n.value


def main_partitioned(args):
    """`python -m torch.distributed.run --nproc-per-node N bench_offline.py --gpus N`: SURVEY.md 8e rows 2 and 4 on N
    GPUs of one box.  The dof rows are partitioned (PETSc-style ownership ranges), every rank uploads only its own rows
    plus the halo its operator rows reference, the partial Gram matrices and the Gram-Schmidt coefficients are summed
    over the ranks in HBM by NCCL (rb_set_allreduce).  One JSON line per product: kernel time = max over ranks (CUDA
    events on the engine's stream), wall time = max over ranks with the all-reduce and the one D2H copy inside."""
    import ctypes as C
    from rbnics_b200 import offline
    from rbnics_b200.engine import SweepEngine
    from rbnics_b200.sweep import DistributedMaxLoc
    if int(os.environ.get("WORLD_SIZE", "1")) != args.gpus:
        raise SystemExit("bench_offline.py --gpus N must run under torch.distributed.run with N ranks")
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = DistributedMaxLoc(local_rank)
    eng = SweepEngine(local_rank)
    side = int(round(args.nh ** 0.5))
    Nh = side * side
    X = lap2d(side)

    def kernel_ms():
        ms = (C.c_float * 5)()
        n = C.c_int()
        eng._lib.rb_last_timings(eng._h, ms, C.byref(n))
        return float(ms[4])

    def emit(**kw):
        if dist.rank == 0:
            print(json.dumps(dict(n_gpus=dist.size, scaling="strong", **kw)), flush=True)

    ws = offline.OfflineWorkspace(eng, dist=dist)
    ws.set_operator("X", X)
    lo, hi = ws.upload_range()
    for Ns in args.ns:
        S = np.ascontiguousarray(np.random.default_rng(Ns).standard_normal((Nh, Ns)))
        name = f"S{Ns}"
        ws.new_block(name, Nh, Ns)
        dist.barrier()
        t0 = time.perf_counter()
        ws.append(name, S)
        up = dist.max_float(time.perf_counter() - t0)
        best_k = best_w = None
        for _ in range(args.reps + 1):
            dist.barrier()
            t0 = time.perf_counter()
            ws.gram(name, "X")
            w = dist.max_float(time.perf_counter() - t0)
            k = dist.max_float(kernel_ms())
            best_k = k if best_k is None else min(best_k, k)
            best_w = w if best_w is None else min(best_w, w)
        emit(kernel="row-partitioned resident Gram S^T X S (NCCL all-reduce of the partials in HBM)",
             shape={"Nh": Nh, "Ns": Ns, "nnz": int(X.nnz)}, kernel_ms_max_over_ranks=best_k,
             wall_ms_with_allreduce_and_d2h=1e3 * best_w, upload_ms_max_over_ranks=1e3 * up,
             h2d_bytes_per_rank=int((hi - lo) * Ns * 8), h2d_bytes_replicated=int(Nh * Ns * 8))
        del S
    for N in args.n:
        name = f"Z{N}"
        ws.new_block(name, Nh, N + 1)
        ws.append(name, np.ascontiguousarray(np.random.default_rng(N).standard_normal((Nh, N))))
        z = np.random.default_rng(2).standard_normal(Nh)
        ws.gram_schmidt(name, "X", z, append=False)
        best = None
        for _ in range(args.reps):
            dist.barrier()
            t0 = time.perf_counter()
            ws.gram_schmidt(name, "X", z, append=False)
            w = dist.max_float(time.perf_counter() - t0)
            best = w if best is None else min(best, w)
        emit(kernel="row-partitioned Gram-Schmidt step (MGS order kept: one all-reduce of 4.7 KB per projection)",
             shape={"Nh": Nh, "N": N}, wall_ms=1e3 * best)
    eng.close()
    dist.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nh", type=int, default=1_000_000)
    ap.add_argument("--ns", type=int, nargs="+", default=[61, 400])
    ap.add_argument("--n", type=int, nargs="+", default=[20, 100])
    ap.add_argument("--q", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--gpus", type=int, default=1,
                    help="N > 1 (under torchrun, one rank per GPU): the dof-row partitioned resident products instead")
    args = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 or args.gpus > 1:
        return main_partitioned(args)
    from rbnics_b200 import offline
    from rbnics_b200.engine import SweepEngine
    eng = SweepEngine(0)
    hbm, fp64 = peaks()
    side = int(round(args.nh ** 0.5))
    Nh = side * side
    X = lap2d(side)
    nnz = X.nnz
    rng = np.random.default_rng(0)

    def emit(kernel, shape, ms, bytes_alg, flops_alg, e2e_ms, note=""):
        gbs = bytes_alg / (ms * 1e-3) * 1e-9 if ms > 0 else 0.0
        tfs = flops_alg / (ms * 1e-3) * 1e-12 if ms > 0 else 0.0
        bound = "hbm" if (flops_alg / max(bytes_alg, 1)) < fp64 * 1e12 / (hbm * 1e9) else "tensor"
        frac = gbs / hbm if bound == "hbm" else tfs / fp64
        print(json.dumps({"kernel": kernel, "shape": shape, "kernel_ms": ms, "e2e_ms_with_copies": e2e_ms,
                          "algorithmic_bytes": bytes_alg, "algorithmic_flops": flops_alg, "achieved_GBps": gbs,
                          "achieved_TFLOPs": tfs, "bound": bound, "peak": hbm if bound == "hbm" else fp64,
                          "frac": frac, "note": note}), flush=True)

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    def cpu_gram(S, Xm):
        """BASELINE.md section 4 comparator: scipy.sparse SpMM + BLAS dgemm with all host threads (oracle gram_blas)."""
        from oracle import rb_oracle as O
        t0 = time.perf_counter()
        Cc = O.gram_blas(S, Xm)
        return Cc, (time.perf_counter() - t0) * 1e3

    for Ns in args.ns:
        S = np.ascontiguousarray(rng.standard_normal((Nh, Ns)))
        best = None
        for _ in range(args.reps):
            t0 = time.perf_counter()
            offline.gram(eng, S, X)
            e2e = (time.perf_counter() - t0) * 1e3
            ms, _ = timings(eng)
            if best is None or ms[4] < best[0][4]:
                best = (ms, e2e)
        ms, e2e = best
        shape = {"Nh": Nh, "Ns": Ns, "nnz": nnz}
        # parity with the measurement + the CPU comparator on the same operands
        Cg = offline.gram(eng, S, X)
        Cc, cpu_ms = cpu_gram(S, X)
        print(json.dumps({"kernel": "Gram S^T X S: parity and CPU baseline", "shape": shape,
                          "max_abs_diff_over_max_abs_C": float(np.max(np.abs(Cg - Cc)) / np.max(np.abs(Cc))),
                          "tolerance": 1e-10, "cpu_baseline": {"ms": cpu_ms, "cores": cores, "kind": "port",
                                                               "what": "scipy.sparse CSR @ dense + numpy (OpenBLAS) S^T Y, all threads"},
                          "gpu_kernel_ms": ms[4], "gpu_e2e_ms_upload_every_call": e2e}), flush=True)
        # resident POD (snapshots appended to HBM as they are produced): one Gram + host eigh + modes + diagonal norms
        ws = offline.OfflineWorkspace(eng)
        ws.set_operator("X", X)
        pod = offline.ResidentProperOrthogonalDecomposition(ws, "X", Nh, Ns)
        t0 = time.perf_counter()
        ws.append(pod.snapshots, S)
        up_ms = (time.perf_counter() - t0) * 1e3
        pod.apply(20, 0.0, read_basis=False)
        kms = {}
        t0 = time.perf_counter()
        Cr = ws.gram(pod.snapshots, "X")
        kms["gram"] = timings(eng)[0][4]
        t1 = time.perf_counter()
        import scipy.linalg
        eigs, eigv = scipy.linalg.eigh(Cr)
        t2 = time.perf_counter()
        V = np.ascontiguousarray(eigv[:, ::-1][:, :20])
        ws.new_block("pod_basis", Nh, 20)
        ws.lincomb(pod.snapshots, V, "pod_basis")
        kms["lincomb"] = timings(eng)[0][4]
        n2 = ws.column_norms2("pod_basis", "X")
        kms["column_norms"] = timings(eng)[0][4]
        ws.scale_columns("pod_basis", 1.0 / np.sqrt(n2))
        t3 = time.perf_counter()
        print(json.dumps({"kernel": "resident POD apply (Gram + host eigh + S V + diagonal norms + scaling), Nmax = 20",
                          "shape": shape, "e2e_ms": (t3 - t0) * 1e3, "host_eigh_ms": (t2 - t1) * 1e3,
                          "gpu_kernel_ms": kms, "gpu_kernel_ms_sum": sum(kms.values()),
                          "e2e_without_eigh_over_kernel_sum": ((t3 - t0) - (t2 - t1)) * 1e3 / sum(kms.values()),
                          "snapshot_upload_ms_once_per_snapshot_set": up_ms,
                          "note": "snapshots cross PCIe once, when they are produced; nothing of size Nh x Ns returns"}),
              flush=True)
        del ws, pod
        if ms[0] == 0.0:
            # ns <= 64: one kernel gathers Y = X S chunk by chunk into shared memory and multiplies by S^T
            emit("rb_gram_fused_kernel + reduce (C = S^T X S, Y never written)", shape, ms[1], 12 * nnz + 8 * Nh * Ns,
                 2 * nnz * Ns + 2 * Nh * Ns * Ns, e2e, "algorithmic bytes of SURVEY.md 8d (one pass over S)")
        else:
            emit("rb_spmm_kernel (Y = X S)", shape, ms[0], 12 * nnz + 16 * Nh * Ns, 2 * nnz * Ns, e2e)
            # X == X^T over all dof rows: only tiles on / above the diagonal are computed -> symmetric flop count
            emit("rb_tsmm2_kernel + reduce (C = S^T Y, symmetric tiles mirrored)", shape, ms[1], 16 * Nh * Ns,
                 Nh * Ns * (Ns + 1), e2e, "flops counted as Nh*Ns*(Ns+1) (SURVEY.md 8d, symmetry exploited); "
                 "Y is re-read from HBM")
        del S
    ops = [X, (X + sp.identity(Nh)).tocsr(), (2.0 * X).tocsr()][:args.q]
    for N in args.n:
        Z = np.ascontiguousarray(rng.standard_normal((Nh, N)))
        t0 = time.perf_counter()
        offline.project_matrices(eng, Z, ops)
        e2e = (time.perf_counter() - t0) * 1e3
        ms, _ = timings(eng)
        shape = {"Nh": Nh, "N": N, "Q": len(ops), "nnz": nnz}
        fused = 48 < N <= 64
        emit("rb_project (Q x fused gather+DMMA)" if fused else "rb_project (Q x (SpMM + Z^T Y))", shape, ms[4],
             len(ops) * (12 * nnz + (8 if fused else 24) * Nh * N),
             len(ops) * (2 * nnz * N + 2 * Nh * N * N), e2e,
             "Z is re-read once per operator" + ("" if fused else "; Y written and re-read"))
        V = np.ascontiguousarray(rng.standard_normal((N, min(N, 20))))
        t0 = time.perf_counter()
        offline.lincomb(eng, Z, V)
        e2e = (time.perf_counter() - t0) * 1e3
        ms, _ = timings(eng)
        emit("rb_lincomb_kernel (Z = S V)", {"Nh": Nh, "Ns": N, "n": V.shape[1]}, ms[4],
             8 * Nh * (N + V.shape[1]), 2 * Nh * N * V.shape[1], e2e)
        gs = offline.GramSchmidt(eng, X)
        z = rng.standard_normal(Nh)
        t0 = time.perf_counter()
        gs.apply(z, Z)
        e2e = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"kernel": "rb_gram_schmidt (MGS step, N dots + axpys)", "shape": {"Nh": Nh, "N": N},
                          "e2e_ms_with_copies": e2e}), flush=True)
        # the same step on a resident basis: only z (8 Nh bytes) crosses PCIe, X Z is cached per column
        ws = offline.OfflineWorkspace(eng)
        ws.set_operator("X", X)
        ws.new_block("Z", Nh, N + 1)
        ws.append("Z", Z)
        ws.gram_schmidt("Z", "X", z, append=False)      # builds the column mirrors once
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            ws.gram_schmidt("Z", "X", z, append=False)
            dt = (time.perf_counter() - t0) * 1e3
            best = dt if best is None else min(best, dt)
        print(json.dumps({"kernel": "rb_off_gram_schmidt (resident basis, MGS step)", "shape": {"Nh": Nh, "N": N},
                          "e2e_ms": best, "h2d_d2h_bytes": 16 * Nh,
                          "algorithmic_bytes": 8 * Nh * (3 * N + 4)}), flush=True)
        del ws
        del Z
    # incremental reduced-operator update of one greedy iteration on resident operands: the new
    # column Z^T A_q z and the new row z^T A_q Z of Q operators -- 2Q fused Gram launches against 2Q sparse mat-vecs
    # written side by side + ONE pass over Z
    Q, N = 10, 20
    ws = offline.OfflineWorkspace(eng)
    ws.new_block("Z", Nh, N)
    ws.append("Z", np.ascontiguousarray(rng.standard_normal((Nh, N))))
    ws.new_block("YA", Nh, 2 * Q)
    for q in range(Q):
        A = (X + 0.1 * (q + 1) * sp.identity(Nh)).tocsr()
        ws.set_operator(f"A{q}", A)
        ws.set_operator(f"A{q}T", A.T.tocsr())

    def old_way():
        return [(ws.gram("Z", f"A{q}", "Z", cols2=(N - 1, N)), ws.gram("Z", f"A{q}", "Z", cols=(N - 1, N)))
                for q in range(Q)]

    def new_way():
        for q in range(Q):
            ws.spmm(f"A{q}", "Z", (N - 1, N), "YA", q)
            ws.spmm(f"A{q}T", "Z", (N - 1, N), "YA", Q + q)
        return ws.gram("Z", None, "YA")

    res = {}
    for name, fn in (("2Q fused Gram launches", old_way), ("2Q mat-vecs + one Gram", new_way)):
        fn()
        t0 = time.perf_counter()
        for _ in range(3):
            out = fn()
        res[name] = (time.perf_counter() - t0) / 3 * 1e3
    ref = old_way()
    new = new_way()
    err = max(max(np.max(np.abs(new[:, q] - ref[q][0][:, 0])), np.max(np.abs(new[:, Q + q] - ref[q][1][0, :])))
              for q in range(Q)) / max(np.max(np.abs(r[0])) for r in ref)
    print(json.dumps({"kernel": "incremental reduced-operator update (new row + column of Q operators)",
                      "shape": {"Nh": Nh, "N": N, "Q": Q, "nnz": nnz}, "wall_ms": res, "max_rel_diff": err,
                      "algorithmic_bytes": 2 * Q * 12 * nnz + 8 * Nh * (N + 2 * Q)}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
