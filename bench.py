#!/usr/bin/env python
"""bench.py -- greedy-sweep parameters/second at N=100, Q_a=50, Q_f=10 over a 10^6-parameter training set.

    python bench.py --gpus N --steps K --warmup W            # the B200 arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm on the host cores

A "step" is one full greedy sweep (assembly + LU solve + error estimator + arg-max) over the whole training set
(BASELINE.json configs[2]); at N > 1 the SAME 10^6 parameters are sharded cyclically over the ranks (the reference's
own data-parallel strategy, rbnics/sampling/parameter_space_subset.py:56-73) => strong scaling.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_RB, Q_A, Q_F = 100, 50, 10
NTRAIN = 1_000_000
METRIC = "greedy_sweep_params_per_sec"
UNIT = "params/s"


def workload_config(world, ntrain):
    """The SAME dict in both arms (the driver compares them key by key)."""
    return {"workload": "synthetic affine elliptic reduced system N=100 Qa=50 Qf=10, training set 10^6 parameters "
                        "(BASELINE.json configs[2])",
            "N": N_RB, "Qa": Q_A, "Qf": Q_F, "ntrain": ntrain, "parallelism": f"dp{world} (training-set shards)",
            "l2": "no flush needed: per-step inputs (Theta 480 MB, assembled-matrix scratch 1.5 GB, packed G' 101 MB) "
                  "exceed the 126 MB L2",
            "timing": "device arm: perf_counter around K blocking sweeps bracketed by barrier+stream sync, max over "
                      "ranks, kernel times from CUDA events on the engine's stream; reference arm: perf_counter around "
                      "each sampled sweep, max over the worker processes"}


TRAIN_CHUNK = 100_000


def global_training_rows(n):
    """The first n <= TRAIN_CHUNK rows (Theta, beta) of the GLOBAL training set every arm works on (the b200 arm
    generates it chunk by chunk from default_rng(1); this is chunk 0)."""
    assert n <= TRAIN_CHUNK
    g = np.random.default_rng(1)
    m = min(TRAIN_CHUNK, NTRAIN)
    ta = g.uniform(0.1, 10.0, size=(m, Q_A))
    tf = g.uniform(-1.0, 1.0, size=(m, Q_F))
    Theta = np.concatenate([ta[:n], tf[:n]], axis=1)
    return Theta, Theta[:, :Q_A].min(axis=1)


def flops_per_param(N, Qa, Qf):
    """Algorithmic FP64 flops of one parameter (DESIGN.md 'Algorithmic work')."""
    asm = 2 * Qa * N * N + 2 * Qf * N
    lu = (2 * N ** 3) // 3 + 2 * N * N
    Kz = Qa * N + Qf
    est_ref = 2 * Qf * Qf + 2 * Qa * Qf * N + 2 * N + 2 * Qa * Qa * N * N + 2 * N * N + 2 * N  # reference formulation
    est_sym = Kz * (Kz + 1) + 3 * Kz  # z^T tril(G+G^T) z (symmetry of G only), z built on the fly, row dot
    # executed formulation: the Q_a blocks of z all scale the same u, so block pair (q, q') only needs the symmetric
    # part of G_qq' + G_q'q in (n, m) as well: Q_a(Q_a+1)/2 pairs x N(N+1)/2 entries (+ af and ff terms)
    macs_fold = (Qa * (Qa + 1) // 2) * (N * (N + 1) // 2) + Qa * N * Qf + Qf * (Qf + 1) // 2
    est_fold = 2 * macs_fold + 3 * Kz
    return {"assembly": asm, "lu": lu, "estimator_reference": est_ref, "estimator_sym": est_sym,
            "estimator_folded": est_fold, "F_alg": asm + lu + est_ref, "F_sym": asm + lu + est_sym,
            "F_fold": asm + lu + est_fold}


def estimator_kernel_name():
    """The estimator kernel the library runs: the pair-product GEMM unless RB_EST_KERNEL=fold (rb_est_pair.cu)."""
    return "rb_estimator_kernel" if os.environ.get("RB_EST_KERNEL") == "fold" else "rb_estimator_pair_kernel"


def estimator_traffic():
    """dram__bytes_read + dram__bytes_write of one launch of the estimator kernel at the bench configuration, from the
    newest `ncu --set full` summary of THAT kernel under profiles/ (tools/ncu_summary.py full ...); (None, reason) if
    there is none."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_estimator_full.json")),
                   key=lambda f: [int(x) if x.isdigit() else x for x in re.split(r"(\d+)", os.path.basename(f))])
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for f in reversed(files):
        try:
            cap = json.load(open(f))["captures"][0]
            if estimator_kernel_name() not in cap.get("kernel", ""):
                continue
            tot = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(cap[k]["value"]) * unit[cap[k]["unit"]]
            return tot, "profiles/" + os.path.basename(f)
        except Exception:
            continue
    return None, "no ncu --set full summary of " + estimator_kernel_name() + " under profiles/"


def fp64_peak_tflops():
    """FP64 roofline denominator: MEASURED_PEAKS.json has no FP64 entry, so the DMMA peak measured on this pool
    by tools/fp64_peak.cu (profiles/FP64_PEAK.json) is used; say which."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mp = json.load(f)
        for k in ("fp64_tflops", "fp64_dmma_tflops"):
            if k in mp:
                return float(mp[k]), "MEASURED_PEAKS.json:" + k
    except Exception:
        pass
    with open(os.path.join(ROOT, "profiles", "FP64_PEAK.json")) as f:
        p = json.load(f)
    return float(p["fp64_dmma_tflops"]), "profiles/FP64_PEAK.json (DMMA.8x8x4 peak measured with tools/fp64_peak.cu; MEASURED_PEAKS.json has no FP64 entry)"


# ----------------------------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference's algorithm; the only place bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One emulated MPI rank of the reference: cyclic shard of the sample, per-parameter loop, local arg-max."""
    rank, size, n_sample, seed_prob, seed_train = args
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import rb_oracle as O
    from rbnics_b200 import synthetic
    prob = synthetic.elliptic_reduced_problem(N_RB, Q_A, Q_F, seed=seed_prob)
    Theta, beta = global_training_rows(n_sample)
    A_q = [prob["A"][q] for q in range(Q_A)]
    f_q = [prob["F"][q] for q in range(Q_F)]
    G_ff = [[prob["Gff"][i, j] for j in range(Q_F)] for i in range(Q_F)]
    G_af = [[prob["Gaf"][i, j] for j in range(Q_F)] for i in range(Q_A)]
    G_aa = [[prob["Gaa"][i, j] for j in range(Q_A)] for i in range(Q_A)]
    idx = list(range(rank, n_sample, size))
    t0 = time.perf_counter()
    d, _, i, v = O.sweep_loop(Theta[idx, :Q_A], Theta[idx, Q_A:], beta[idx], A_q, f_q, G_ff, G_af, G_aa, N_RB)
    return time.perf_counter() - t0, len(idx), float(v), idx[i]


def cpu_reference_loop(n_per_core, cores):
    """R2 of BASELINE.md: the reference's per-parameter loop under its own cyclic MPI sharding, emulated with one
    process per core (mpi4py is not installed).  Returns params/s."""
    import multiprocessing as mp
    n_sample = n_per_core * cores
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        out = pool.map(_cpu_worker, [(r, cores, n_sample, 0, 1) for r in range(cores)])
    tmax = max(o[0] for o in out)
    return n_sample / tmax, n_sample, tmax


# ---- the REAL reference (RBniCS from baseline/_ref or /root/reference, oracle/reference_driver.py), one process per core
_REF_STATE = {}


def _ref_worker_init():
    """Runs once in every worker: import the unmodified reference (numpy online backend) and build its
    EllipticCoerciveRBReducedProblem around the bench operators."""
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    import tempfile as _tf
    os.chdir(_tf.mkdtemp(prefix="rbnics_ref_"))
    from oracle import reference_driver as R
    R.setup_reference_import("numpy")
    import rbnics  # noqa: F401
    from rbnics_b200 import synthetic
    R.disable_reduced_solution_cache()
    prob = synthetic.elliptic_reduced_problem(N_RB, Q_A, Q_F, seed=0)
    _REF_STATE["R"] = R
    _REF_STATE["rp"] = R.elliptic_reduced_problem(prob["A"], prob["F"], prob["Gff"], prob["Gaf"], prob["Gaa"])


def _ref_worker_sweep(args):
    """One rank of the reference's own data-parallel strategy: cyclic shard (parameter_space_subset.py:57), the
    closure of RBReduction._greedy per parameter, local arg-max."""
    rank, size, n_sample = args
    R, rp = _REF_STATE["R"], _REF_STATE["rp"]
    Theta, _ = global_training_rows(n_sample)
    idx = list(range(rank, n_sample, size))
    ts = R.training_set(Theta[idx])
    t0 = time.perf_counter()
    v, i = R.greedy_sweep(rp, ts)
    return time.perf_counter() - t0, len(idx), float(v), idx[int(i)]


class ReferencePool:
    """Worker processes that keep the imported reference alive across steps."""

    def __init__(self, cores):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_ref_worker_init)

    def sweep(self, n_per_core):
        n_sample = n_per_core * self.cores
        out = self.pool.map(_ref_worker_sweep, [(r, self.cores, n_sample) for r in range(self.cores)])
        tmax = max(o[0] for o in out)
        best = max(out, key=lambda o: (o[2], -o[3]))
        return n_sample / tmax, n_sample, tmax, best[2], best[3]

    def close(self):
        self.pool.close()
        self.pool.join()


def reference_available():
    from oracle import reference_driver as R
    return R.reference_root() is not None


def cpu_vectorized(n_sample, prob):
    """R3 of BASELINE.md: chunked einsum / batched numpy.linalg.solve with all BLAS threads."""
    from oracle import rb_oracle as O
    Theta, beta = global_training_rows(n_sample)
    t0 = time.perf_counter()
    O.sweep_batched(Theta[:, :Q_A], Theta[:, Q_A:], beta, prob["A"], prob["F"], prob["Gff"], prob["Gaf"], prob["Gaa"],
                    chunk=2048)
    dt = time.perf_counter() - t0
    return n_sample / dt, dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    """The reference's own CPU implementation of the path on the host cores: the REAL RBniCS classes (kind
    "reference") when the package is present, else the loop-faithful oracle port (kind "port").  Every step is a
    bounded sample of the workload; `ms_per_step` is the measured time of a sampled step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_core = max(10, int(args.cpu_params_per_core))
    real = reference_available() and not args.reference_port
    vals, step_s = [], []
    t_all = time.perf_counter()
    if real:
        pool = ReferencePool(cores)
        for _ in range(max(1, min(args.warmup, 2))):
            pool.sweep(4)  # imports + first-call dispatch caches
        for _ in range(args.steps):
            v, n_sample, tmax, vmax, imax = pool.sweep(per_core)
            vals.append(v)
            step_s.append(tmax)
        pool.close()
        kind = "reference"
        what = ("the unmodified RBniCS package (baseline/_ref): EllipticCoerciveRBReducedProblem.set_mu / solve / "
                "estimate_error on the numpy online backend, driven by ParameterSpaceSubset.max with the closure of "
                "RBReduction._greedy; third-party imports absent from the image served by oracle/shims; reduced-"
                "solution cache disabled")
    else:
        cpu_reference_loop(4, cores)
        for _ in range(args.steps):
            v, n_sample, tmax = cpu_reference_loop(per_core, cores)
            vals.append(v)
            step_s.append(tmax)
        kind = "port"
        what = "loop-faithful numpy port of the per-parameter path (oracle/rb_oracle.py sweep_loop)"
    wall = time.perf_counter() - t_all
    value = statistics.median(vals)
    from rbnics_b200 import synthetic
    prob = synthetic.elliptic_reduced_problem(N_RB, Q_A, Q_F, seed=0)
    vec, vec_dt = cpu_vectorized(4096 * 2, prob)
    fl = flops_per_param(N_RB, Q_A, Q_F)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(step_s),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, NTRAIN),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{n_sample} parameters per step ({per_core} per core; the first rows of the same "
                                   f"global training set as the b200 arm, cyclic shards, one process per core = the "
                                   f"reference's mpirun strategy); value = sampled parameters / max worker time; {what}"},
        "sampled_params_per_step": n_sample,
        "extrapolated_ms_per_full_sweep": 1e3 * NTRAIN / value,
        "cpu_vectorized": {"value": vec, "unit": UNIT, "cores": cores,
                           "sample": f"{4096 * 2} parameters, chunked einsum + batched numpy.linalg.solve"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops_fp64_F_alg": fl["F_alg"] * value * 1e-9,
        "wall_s": wall,
    }
    emit_json_line(line)


# ----------------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.tmp.read().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            load = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
            out = {"sm_mhz": statistics.median(load), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": max(power)}
        try:
            os.unlink(self.tmp.name)
        except Exception:
            pass
        return out


def run_b200_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    ntrain = args.ntrain

    # ---- CPU baseline first (rank 0, N=1 only), before any CUDA context exists in this process
    cpu_baseline = None
    cpu_vec = None
    cores = host_cores()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        per_core = max(10, int(args.cpu_params_per_core))
        if reference_available() and not args.reference_port:
            pool = ReferencePool(cores)
            pool.sweep(4)
            v, n_sample, tmax, _, _ = pool.sweep(per_core)
            pool.close()
            kind, what = "reference", ("the unmodified RBniCS package (baseline/_ref, numpy online backend): "
                                       "EllipticCoerciveRBReducedProblem.solve / estimate_error under "
                                       "ParameterSpaceSubset.max, one process per core")
        else:
            v, n_sample, tmax = cpu_reference_loop(per_core, cores)
            kind, what = "port", "loop-faithful numpy port of the per-parameter path (oracle/rb_oracle.py sweep_loop)"
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"{n_sample} parameters ({n_sample // cores} per core, the first rows of the same "
                                  f"training set, cyclic shards over {cores} processes; {tmax:.1f} s); {what}"}

    from rbnics_b200 import synthetic
    from rbnics_b200.engine import SweepEngine, pinned_empty, pinned_free
    from rbnics_b200.sweep import DistributedMaxLoc

    prob = synthetic.elliptic_reduced_problem(N_RB, Q_A, Q_F, seed=0)
    if cpu_baseline is not None:
        vec, vec_dt = cpu_vectorized(8192, prob)
        cpu_vec = {"value": vec, "unit": UNIT, "cores": cores,
                   "sample": f"8192 parameters ({vec_dt:.1f} s), chunked einsum + batched numpy.linalg.solve, all BLAS threads"}

    dist = DistributedMaxLoc(local_rank) if world > 1 else None
    eng = SweepEngine(local_rank)
    eng.set_system(prob["A"], prob["F"])
    eng.set_elliptic_estimator(prob["Gff"], prob["Gaf"], prob["Gaa"])

    # this rank's cyclic shard of the SAME global training set (identical seed on every rank)
    n_local = len(range(rank, ntrain, world))
    Theta_p, hT = pinned_empty((n_local, Q_A + Q_F))
    beta_p, hB = pinned_empty((n_local,))
    chunk = TRAIN_CHUNK
    rng_done = 0
    # generate the global set in chunks and keep rows rank::world (bitwise identical across ranks)
    g = np.random.default_rng(1)
    lo = 0
    while rng_done < ntrain:
        n = min(chunk, ntrain - rng_done)
        ta = g.uniform(0.1, 10.0, size=(n, Q_A))
        tf = g.uniform(-1.0, 1.0, size=(n, Q_F))
        first = (rank - rng_done) % world
        rows = slice(first, n, world)
        k = len(range(first, n, world))
        Theta_p[lo:lo + k, :Q_A] = ta[rows]
        Theta_p[lo:lo + k, Q_A:] = tf[rows]
        lo += k
        rng_done += n
    beta_p[:] = Theta_p[:, :Q_A].min(axis=1)

    def barrier():
        if dist is not None:
            dist.barrier()

    def step_resident():
        if dist is not None:   # the collective reads the engine's device-resident (max, argmax, #non-finite) directly
            eng.sweep_resident(raise_on_singular=False)
            v, i, bad = dist.maxloc_device(eng)
            if bad:
                raise RuntimeError("non-finite error estimators in the bench sweep")
            return v, i
        r = eng.sweep_resident()
        return r["max"], r["argmax"]

    def step_e2e():
        r = eng.sweep(Theta_p, beta_p, index_offset=rank, index_stride=world)
        if dist is not None:
            return dist.maxloc(r["max"], r["argmax"])
        return r["max"], r["argmax"]

    # ---- device-resident timing
    eng.set_training_set(Theta_p, beta_p, index_offset=rank, index_stride=world)
    for _ in range(args.warmup):
        res = step_resident()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    est_ms, tot_ms, launches = [], [], 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_resident()
        tm = eng.last_timings()
        est_ms.append(tm["estimator_ms"])
        tot_ms.append(tm["total_ms"])
        launches += tm["launches"]
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    if dist is not None:
        dt = dist.max_float(dt)
        launches = int(dist.sum_float(float(launches)))
    ms_per_step = 1e3 * dt / args.steps
    value = ntrain / (dt / args.steps)

    # ---- end-to-end timing: pinned host buffers in, (max, argmax) out, copies inside
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_e2e = step_e2e()
    barrier()
    dt2 = time.perf_counter() - t0
    if dist is not None:
        dt2 = dist.max_float(dt2)
    e2e_value = ntrain / (dt2 / args.steps)
    h2d = (n_local * (Q_A + Q_F) + n_local) * 8 * world
    d2h = 24 * world

    # ---- parity with the measurement: oracle on a seeded subsample of the SAME Theta, top-2 gap
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import rb_oracle as O
        full = eng.sweep_resident(want_delta=True)
        dl = full["delta"]
        top2 = np.sort(dl[np.argpartition(dl, -2)[-2:]])
        sub = np.sort(np.random.default_rng(12345).choice(n_local, size=min(2000, n_local), replace=False))
        sub[0] = (int(full["argmax"]) - rank) // world  # the arg-max itself is always among the checked rows
        sub = np.unique(sub)
        Ts, bs = np.ascontiguousarray(Theta_p[sub]), np.ascontiguousarray(beta_p[sub])
        gpu = eng.sweep(Ts, bs, want_delta=True, want_u=True)
        d_ref, U_ref, _, _ = O.sweep_batched(Ts[:, :Q_A], Ts[:, Q_A:], bs, prob["A"], prob["F"], prob["Gff"],
                                             prob["Gaf"], prob["Gaa"], chunk=500)
        k = int(np.where(sub == (int(full["argmax"]) - rank) // world)[0][0])
        parity = {
            "checked_params": int(sub.size),
            "u_rel_max": float(np.max(np.max(np.abs(gpu["u"] - U_ref), axis=1) / np.max(np.abs(U_ref), axis=1))),
            "delta_rel_max": float(np.max(np.abs(gpu["delta"] - d_ref) / np.abs(d_ref))),
            "delta_rel_resident_vs_subsample": float(np.max(np.abs(gpu["delta"] - dl[sub]) / np.abs(dl[sub]))),
            "argmax_delta_rel": float(abs(full["max"] - d_ref[k]) / abs(d_ref[k])),
            "top2_gap_rel": float((top2[1] - top2[0]) / top2[1]),
            "tolerance": 1e-10,
            "oracle": "oracle/rb_oracle.py sweep_batched (numpy einsum + LAPACK dgesv) on rows of the timed training set",
        }
        parity["ok"] = bool(max(parity["u_rel_max"], parity["delta_rel_max"], parity["argmax_delta_rel"]) <= 1e-10
                            and parity["top2_gap_rel"] > 1e-10)
        eng.set_training_set(Theta_p, beta_p, index_offset=rank, index_stride=world)

    if rank == 0:
        fl = flops_per_param(N_RB, Q_A, Q_F)
        peak, peak_src = fp64_peak_tflops()
        est_launch_ms = statistics.mean(est_ms)
        achieved = fl["estimator_folded"] * n_local / (est_launch_ms * 1e-3) * 1e-12
        work = eng.estimator_work()
        issued = 2 * work["issued_macs_per_param"] * n_local / (est_launch_ms * 1e-3) * 1e-12
        traffic, traffic_src = estimator_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, ntrain),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": estimator_kernel_name() + " (FP64 DMMA.8x8x4)", "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "peak_source": peak_src, "launch_ms": est_launch_ms,
                         "algorithmic_flops_per_param": fl["estimator_folded"], "params_per_launch": n_local,
                         "note": "algorithmic = minimal multiply-adds of the executed (doubly folded) quadratic form; "
                                 "the kernel issues these plus padding (pair-product GEMM: K to 8 and columns to 64 per "
                                 "class pair; folded kernel: its triangular 16-column tiles); launch_ms spans the "
                                 "estimator launch(es) of one sweep (large sweeps: whole waves + a finer-split launch "
                                 "for the last row blocks)",
                         "dmma_issued_tflops": issued, "dmma_issue_frac": issued / peak,
                         "issued_flops_per_param": 2 * work["issued_macs_per_param"]},
            "sweep": {"F_fold_flops_per_param": fl["F_fold"], "F_sym_flops_per_param": fl["F_sym"],
                      "F_alg_flops_per_param": fl["F_alg"],
                      "fp64_tflops_F_fold": fl["F_fold"] * value * 1e-12,
                      "frac_of_peak_F_fold": fl["F_fold"] * value * 1e-12 / (peak * world),
                      "equivalent_tflops_reference_formulation_F_alg": fl["F_alg"] * value * 1e-12,
                      "kernel_ms_mean": {"assemble_lu": statistics.mean(tot_ms) - est_launch_ms, "estimator": est_launch_ms},
                      "argmax": int(res[1]), "max": float(res[0]), "argmax_e2e": int(res_e2e[1])},
        }
        if parity is not None:
            line["parity"] = parity
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
            line["cpu_vectorized"] = cpu_vec
        emit_json_line(line)
    pinned_free(hT)
    pinned_free(hB)
    eng.close()
    if dist is not None:
        dist.close()


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL prints its version banner) must not add lines to stdout: route fd 1 to stderr and keep the
    real stdout for the single JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ntrain", type=int, default=NTRAIN)
    ap.add_argument("--cpu-params-per-core", type=int, default=40,
                    help="sample size per host core and step of the CPU reference (the real reference takes ~0.14 s per parameter and core at N=100, Q_a=50)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check that runs with the measurement")
    ap.add_argument("--reference-port", action="store_true",
                    help="time the oracle port instead of the real reference package even if the latter is present")
    args = ap.parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
