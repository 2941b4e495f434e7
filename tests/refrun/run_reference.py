#!/usr/bin/env python
"""Subprocess entry of the tests that execute the REAL reference package (oracle/reference_driver.py): the online
backend of RBniCS is chosen at import time, so every backend gets its own interpreter.

    python tests/refrun/run_reference.py names   <backend>           -> JSON: exported names, type modules, dispatch tables
    python tests/refrun/run_reference.py sweep   <backend> <N> <Qa> <Qf> <B> <seed> <compliant> <out.npz>
    python tests/refrun/run_reference.py stokes  <backend> <case> <out.npz>   StokesRBReducedProblem (components u, s, p:
                 component-dictionary slicing, 9 error-estimation term pairs) on the operators of reference_stokes_<case>.npz
    python tests/refrun/run_reference.py parabolic <backend> <out.npz> [plugin]   ParabolicCoerciveRBReducedProblem + the
                 POD-Greedy closure on the operators of reference_parabolic.npz
    python tests/refrun/run_reference.py storage <backend> <directory>   -> writes storages with the reference's own save()
    python tests/refrun/run_reference.py greedy  <backend> <case> <plugin 0|1> <out.npz>
                 the reference's complete offline stage, ReducedBasis(problem).offline(), on a dense FEniCS-free truth
                 problem (oracle/reference_truth.py); plugin = 1 decorates the problem with @B200Sweep(), so that the
                 reference's own greedy loop calls the batched sweep on the GPU
    python tests/refrun/run_reference.py pod_galerkin numpy <out.npz>
                 the reference's complete POD-Galerkin offline stage, PODGalerkin(problem).offline(), on the dense
                 thermal block: snapshots, POD eigenvalues / retained energy, basis, reduced operators
    python tests/refrun/run_reference.py pod_greedy <backend> <case> <out.npz> [plugin]
                 the reference's complete POD-Greedy offline stage of a parabolic problem (BASELINE configs[3]):
                 ReducedBasis(ParabolicCoerciveProblem).offline() with set_Nmax(Nmax, POD_Greedy=N1)
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_driver as R  # noqa: E402


# real-reference greedy runs: thermal block of tests/truth_problems.py (n x n grid, vertical strips)
GREEDY_CASES = {
    # tutorials/01_thermal_block as it is: Q_a = 2, Q_f = 1, ntrain = 100, Nmax = 4, tol = 1e-5, compliant estimator
    "tutorial01": dict(n=12, nblocks=2, nloads=1, compliant=True, a_range=(0.1, 10.0), ntrain=100, Nmax=4, tol=1e-5, seed=0),
    # BASELINE.json configs[0] as written: Q_a = 3, ntrain = 100, N_max = 20
    # (tol stops the loop above the round-off floor of the estimator, ~1e-8 relative, below which the reference itself
    # re-selects parameters that are already in the basis)
    "stress3": dict(n=14, nblocks=3, nloads=1, compliant=True, a_range=(0.1, 10.0), ntrain=100, Nmax=20, tol=2e-7, seed=1),
    # shapes of tutorials/02_elastic_block: Q_a = 9, Q_f = 3, 11 parameters (8 stiffness in [1, 100] + 3 loads), non-compliant
    "config2": dict(n=18, nblocks=9, nloads=3, compliant=False, a_range=(1.0, 100.0), ntrain=200, Nmax=20, tol=1e-12, seed=2),
    # the reference's OWN tutorial meshes (tests/golden/mesh_*.npz) with the P1 assembler of tests/truth_problems.py
    "tutorial01_mesh": dict(truth="ThermalBlockTutorialMesh", nblocks=2, nloads=1, compliant=True, a_range=(0.1, 10.0),
                            ntrain=100, Nmax=4, tol=1e-5, seed=0),
    "tutorial02_mesh": dict(truth="ElasticBlockTutorialMesh", nblocks=9, nloads=3, compliant=False, a_range=(1.0, 100.0),
                            ntrain=200, Nmax=20, tol=1e-12, seed=3),
}


# real-reference POD-Greedy runs: unsteady thermal block of tests/truth_problems.py (tutorial 06 terms)
POD_GREEDY_CASES = {
    "unsteady2": dict(n=10, nblocks=2, nloads=1, dt=0.05, K=20, ntrain=60, Nmax=8, N1=2, tol=1e-6, seed=10),
    "unsteady3": dict(n=12, nblocks=3, nloads=1, dt=0.05, K=30, ntrain=60, Nmax=9, N1=3, tol=1e-6, seed=12),
}


def main():
    mode, backend = sys.argv[1], sys.argv[2]
    for k in range(3, len(sys.argv)):  # output paths are given relative to the caller's directory
        if sys.argv[k].endswith(".npz") or (mode == "storage" and k == 3):
            sys.argv[k] = os.path.abspath(sys.argv[k])
    work = tempfile.mkdtemp(prefix="rbnics_ref_")
    os.chdir(work)
    R.setup_reference_import(backend, rbnicsrc_dir=work)
    import numpy as np
    import rbnics  # noqa: F401
    import rbnics.backends.online as online

    if mode == "names":
        import rbnics.backends as generic
        import rbnics.backends.online.numpy as numpy_backend
        import rbnics.backends.online.numpy.wrapping as numpy_wrapping
        wrapping = sys.modules["rbnics.backends.online.wrapping"]
        out = {
            "online_all": sorted(online.__all__),
            "numpy_all": sorted(numpy_backend.__all__),
            "wrapping_all": sorted(wrapping.__all__),
            "numpy_wrapping_all": sorted(numpy_wrapping.__all__),
            "wrapping_module": wrapping.__name__,
            "modules": {n: getattr(online, n).__module__ for n in online.__all__},
            "types": {n: getattr(online, n).Type().__module__ + "." + getattr(online, n).Type().__name__
                      for n in ("OnlineMatrix", "OnlineVector", "OnlineFunction")},
            "distinct_types": all(getattr(online, n).Type() is not getattr(numpy_backend, n[6:]).Type()
                                  for n in ("OnlineMatrix", "OnlineVector", "OnlineFunction")),
            "subclass_of_numpy_types": all(issubclass(getattr(online, n).Type(), getattr(numpy_backend, n[6:]).Type())
                                           for n in ("OnlineMatrix", "OnlineVector", "OnlineFunction")),
            # signatures registered in the generic dispatchers rbnics.backends.<name> (module names of the implementations)
            "dispatch": {n: sorted(set(f.__module__ for f in getattr(generic, n).funcs.values()))
                         for n in ("product", "sum", "LinearSolver", "transpose", "Matrix", "Vector", "Function",
                                   "AffineExpansionStorage")},
        }
        print(json.dumps(out))
        return

    if mode == "sweep":
        N, Qa, Qf, B, seed, compliant = (int(x) for x in sys.argv[3:9])
        outfile = sys.argv[9]
        from rbnics_b200 import synthetic
        R.disable_reduced_solution_cache()
        prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=seed, riesz_rank=4 * N)
        Theta, beta, _ = synthetic.training_set(B, Qa, Qf, seed=seed + 1)
        rp = R.elliptic_reduced_problem(prob["A"], prob["F"], prob["Gff"], prob["Gaf"], prob["Gaa"],
                                        compliant=bool(compliant))
        ts = R.training_set(Theta)
        col = {}
        mx, am = R.greedy_sweep(rp, ts, col)
        np.savez(outfile, N=N, Qa=Qa, Qf=Qf, B=B, seed=seed, compliant=compliant, Theta=Theta, beta=beta,
                 A=prob["A"], F=prob["F"], Gff=prob["Gff"], Gaf=prob["Gaf"], Gaa=prob["Gaa"],
                 delta=np.array(col["delta"]), u=np.array(col["u"]), max=float(mx), argmax=int(am),
                 backend=online.OnlineMatrix.__module__)
        print(json.dumps({"max": float(mx), "argmax": int(am), "backend": online.OnlineMatrix.__module__}))
        return

    if mode == "stokes":
        case, outfile = sys.argv[3], sys.argv[4]
        g = np.load(os.path.join(ROOT, "tests", "golden", "reference_stokes_%s.npz" % case))
        assert int(g["n_bc"]) == 0
        pairs = [("f", "f"), ("g", "g"), ("a", "f"), ("bt", "f"), ("b", "g"), ("a", "a"), ("a", "bt"), ("bt", "bt"), ("b", "b")]
        R.disable_reduced_solution_cache()
        rp = R.stokes_reduced_problem({t: g["op_" + t] for t in ("a", "b", "bt", "f", "g")},
                                      {(t1, t2): g["E_%s_%s" % (t1, t2)] for (t1, t2) in pairs}, int(g["n"]))
        Theta = np.concatenate([g["theta_" + t] for t in ("a", "b", "bt", "f", "g")] + [g["beta"][:, None]], axis=1)
        col = {}
        ts = R.training_set(Theta)
        mx, am = R.greedy_sweep(rp, ts, col)
        out = {"max": float(mx), "argmax": int(am), "backend": online.OnlineMatrix.__module__}
        if len(sys.argv) > 5 and sys.argv[5] == "plugin":
            # the Stokes family of the _greedy override (integration/rbnics_b200_plugin/sweep.py) on the same reduced problem:
            # ONE batched sweep instead of the per-parameter closure
            import types
            integ = os.path.join(ROOT, "integration")
            if integ not in sys.path:
                sys.path.append(integ)
            from rbnics_b200_plugin.sweep import GreedySweep
            gs = GreedySweep(types.SimpleNamespace(truth_problem=rp.truth_problem, training_set=ts))
            v, i = gs.max(rp)
            out["plugin_max"], out["plugin_argmax"] = float(v), int(i)
            gs.close()
        np.savez(outfile, u=np.array(col["u"]), delta=np.array(col["delta"]), max=float(mx), argmax=int(am),
                 backend=online.OnlineMatrix.__module__)
        print(json.dumps(out))
        return

    if mode == "parabolic":
        outfile = sys.argv[3]
        g = np.load(os.path.join(ROOT, "tests", "golden", "reference_parabolic.npz"))
        pairs = [("f", "f"), ("a", "f"), ("a", "a"), ("m", "f"), ("m", "a"), ("m", "m")]
        # (the RAM cache stays on: the reference's time-dependent monitor re-reads what it has just stored in it)
        rp = R.parabolic_reduced_problem(g["M"], g["A"], g["F"], {(a, b): g["E_%s_%s" % (a, b)] for (a, b) in pairs},
                                         float(g["dt"]), int(g["K"]))
        Theta = np.concatenate([g["Theta_m"], g["Theta_a"], g["Theta_f"], g["beta"][:, None]], axis=1)
        ts = R.training_set(Theta)
        col = {}
        mx, am = R.pod_greedy_sweep(rp, ts, col)
        out = {"max": float(mx), "argmax": int(am), "backend": online.OnlineMatrix.__module__}
        if len(sys.argv) > 4 and sys.argv[4] == "plugin":
            import types
            integ = os.path.join(ROOT, "integration")
            if integ not in sys.path:
                sys.path.append(integ)
            from rbnics_b200_plugin.sweep import GreedySweep
            gs = GreedySweep(types.SimpleNamespace(truth_problem=rp.truth_problem, training_set=ts))
            v, i = gs.max(rp)
            out["plugin_max"], out["plugin_argmax"] = float(v), int(i)
            gs.close()
        np.savez(outfile, delta=np.array(col["delta"]), max=float(mx), argmax=int(am))
        print(json.dumps(out))
        return

    if mode == "storage":
        directory = sys.argv[3]
        from rbnics.backends.online import OnlineAffineExpansionStorage, OnlineMatrix, OnlineVector
        from rbnics.utils.io import OnlineSizeDict
        rng = np.random.default_rng(5)
        # single-component operators (as written for an elliptic problem)
        st = OnlineAffineExpansionStorage(3)
        for q in range(3):
            m = OnlineMatrix({"u": 4}, {"u": 4})
            m.content[:] = rng.standard_normal((4, 4))
            st[q] = m
        st.save(directory, "operator_a")
        sv = OnlineAffineExpansionStorage(2)
        for q in range(2):
            v = OnlineVector({"u": 4})
            v.content[:] = rng.standard_normal(4)
            sv[q] = v
        sv.save(directory, "operator_f")
        ss = OnlineAffineExpansionStorage(2, 2)
        for i in range(2):
            for j in range(2):
                ss[i, j] = float(rng.standard_normal())
        ss.save(directory, "error_estimation_ff")
        # items created with integer sizes (no component dictionaries)
        si = OnlineAffineExpansionStorage(2)
        for q in range(2):
            m = OnlineMatrix(3, 3)
            m.content[:] = rng.standard_normal((3, 3))
            si[q] = m
        si.save(directory, "int_sized_matrices")
        sj = OnlineAffineExpansionStorage(2)
        for q in range(2):
            v = OnlineVector(3)
            v.content[:] = rng.standard_normal(3)
            sj[q] = v
        sj.save(directory, "int_sized_vectors")
        # multi-component (Stokes: u, s, p) operator, order 2 storage of matrices
        M = OnlineSizeDict()
        M["u"], M["s"], M["p"] = 3, 3, 2
        sm = OnlineAffineExpansionStorage(2, 1)
        for i in range(2):
            m = OnlineMatrix(M, M)
            m.content[:] = rng.standard_normal((8, 8))
            sm[i, 0] = m
        sm.save(directory, "stokes_block")
        # and the sliced view the reference takes online: [:N, :N] with a component dict
        Nd = OnlineSizeDict()
        Nd["u"], Nd["s"], Nd["p"] = 2, 1, 2
        sl = sm[:Nd, :Nd]
        np.save(os.path.join(directory, "stokes_block_sliced_0.npy"), np.asarray(sl[0, 0].content))
        np.save(os.path.join(directory, "stokes_block_sliced_1.npy"), np.asarray(sl[1, 0].content))
        print(json.dumps({"ok": True}))
        return
    if mode == "greedy":
        case, plugin, outfile = sys.argv[3], int(sys.argv[4]), sys.argv[5]
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import truth_problems
        from truth_problems import ThermalBlock
        from oracle import reference_truth as T
        cfg = GREEDY_CASES[case]
        T.install_numpy_truth_backend_helpers()
        T.ram_cache_only()
        decorator = None
        if plugin:
            integ = os.path.join(ROOT, "integration")
            if integ not in sys.path:
                sys.path.append(integ)
            from rbnics_b200_plugin.sweep import B200Sweep
            decorator = B200Sweep()
        if "truth" in cfg:
            tb = getattr(truth_problems, cfg["truth"])()
        else:
            tb = ThermalBlock(n=cfg["n"], nblocks=cfg["nblocks"], nloads=cfg["nloads"])
        problem = T.make_thermal_block_problem(tb, compliant=cfg["compliant"], decorator=decorator)
        mu_range = [cfg["a_range"]] * (cfg["nblocks"] - 1) + [(-1.0, 1.0)] * cfg["nloads"]
        import contextlib
        import io
        log = io.StringIO()
        with contextlib.redirect_stdout(log):
            rm, rp = T.run_offline(problem, mu_range, cfg["ntrain"], cfg["Nmax"], cfg["tol"], cfg["seed"])
        ts = [tuple(float(x) for x in mu) for mu in rm.training_set]
        selected = [tuple(float(x) for x in mu) for mu in rm.greedy_selected_parameters]
        indices = [ts.index(mu) for mu in selected]
        N = rp.N
        Qa, Qf = rp.Q["a"], rp.Q["f"]
        A = np.array([np.asarray(rp.operator["a"][q].content)[:N, :N] for q in range(Qa)])
        F = np.array([np.asarray(rp.operator["f"][q].content)[:N] for q in range(Qf)])
        e = rp.error_estimation_operator
        Gff = np.array([[e["f", "f"][i, j] for j in range(Qf)] for i in range(Qf)], dtype=np.float64)
        Gaf = np.array([[np.asarray(e["a", "f"][i, j].content)[:N] for j in range(Qf)] for i in range(Qa)])
        Gaa = np.array([[np.asarray(e["a", "a"][i, j].content)[:N, :N] for j in range(Qa)] for i in range(Qa)])
        Z = np.array([np.asarray(rp.basis_functions[i].vector().content) for i in range(N)]).T
        np.savez(outfile, case=case, training_set=np.array(ts), selected_indices=np.array(indices),
                 error_estimators=np.array([float(x) for x in rm.greedy_error_estimators]), N=N, A=A, F=F, Gff=Gff,
                 Gaf=Gaf, Gaa=Gaa, Z=Z, plugin=plugin, backend=online.OnlineMatrix.__module__,
                 reduction_method=type(rm).__mro__[1].__qualname__ if plugin else type(rm).__qualname__)
        print(json.dumps({"N": int(N), "selected_indices": indices,
                          "error_estimators": [float(x) for x in rm.greedy_error_estimators],
                          "uses_plugin": bool(plugin) and hasattr(rm, "_b200_greedy_sweep")}))
        return
    if mode == "pod_galerkin":
        outfile = sys.argv[3]
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from truth_problems import ThermalBlock
        from oracle import reference_truth as T
        from rbnics import PODGalerkin
        cfg = dict(n=12, nblocks=3, nloads=2, ntrain=30, Nmax=10, tol=1e-9, seed=4)
        T.install_numpy_truth_backend_helpers()
        T.ram_cache_only()
        tb = ThermalBlock(n=cfg["n"], nblocks=cfg["nblocks"], nloads=cfg["nloads"])
        problem = T.make_thermal_block_problem(tb, compliant=False)
        problem.set_mu_range(tb.mu_range)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            rm = PODGalerkin(problem)
            rm.set_Nmax(cfg["Nmax"])
            rm.set_tolerance(cfg["tol"])
            np.random.seed(cfg["seed"])
            rm.initialize_training_set(cfg["ntrain"])
            rp = rm.offline()
        ts = [tuple(float(x) for x in mu) for mu in rm.training_set]
        N = rp.N
        Qa, Qf = rp.Q["a"], rp.Q["f"]
        A = np.array([np.asarray(rp.operator["a"][q].content)[:N, :N] for q in range(Qa)])
        F = np.array([np.asarray(rp.operator["f"][q].content)[:N] for q in range(Qf)])
        Z = np.array([np.asarray(rp.basis_functions[i].vector().content) for i in range(N)]).T
        np.savez(outfile, training_set=np.array(ts), N=N, A=A, F=F, Z=Z,
                 eigenvalues=np.array([float(e) for e in rm.POD.eigenvalues]),
                 retained_energy=np.array([float(e) for e in rm.POD.retained_energy]),
                 **{k: np.asarray(v) for k, v in cfg.items()})
        print(json.dumps({"N": int(N), "eigenvalues": [float(e) for e in rm.POD.eigenvalues][:N]}))
        return
    if mode == "pod_greedy":
        case, outfile = sys.argv[3], sys.argv[4]
        plugin = len(sys.argv) > 5 and sys.argv[5] == "plugin"
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from truth_problems import ThermalBlockUnsteady
        from oracle import reference_truth as T
        cfg = POD_GREEDY_CASES[case]
        T.install_numpy_truth_backend_helpers()
        T.ram_cache_only()
        decorator = None
        if plugin:   # the reference's own POD-Greedy loop with every training_set.max(closure) on the GPU
            integ = os.path.join(ROOT, "integration")
            if integ not in sys.path:
                sys.path.append(integ)
            from rbnics_b200_plugin.sweep import B200Sweep
            decorator = B200Sweep()
        tb = ThermalBlockUnsteady(n=cfg["n"], nblocks=cfg["nblocks"], nloads=cfg["nloads"], dt=cfg["dt"], K=cfg["K"])
        problem = T.make_unsteady_thermal_block_problem(tb, decorator=decorator)
        import contextlib
        import io
        log = io.StringIO()
        with contextlib.redirect_stdout(log):
            rm, rp = T.run_pod_greedy_offline(problem, cfg["ntrain"], cfg["Nmax"], cfg["N1"], cfg["tol"], cfg["seed"])
        ts = [tuple(float(x) for x in mu) for mu in rm.training_set]
        selected = [tuple(float(x) for x in mu) for mu in rm.greedy_selected_parameters]
        indices = [ts.index(mu) for mu in selected]
        N = rp.N
        Qm, Qa, Qf = rp.Q["m"], rp.Q["a"], rp.Q["f"]
        Mr = np.array([np.asarray(rp.operator["m"][q].content)[:N, :N] for q in range(Qm)])
        A = np.array([np.asarray(rp.operator["a"][q].content)[:N, :N] for q in range(Qa)])
        F = np.array([np.asarray(rp.operator["f"][q].content)[:N] for q in range(Qf)])
        Z = np.array([np.asarray(rp.basis_functions[i].vector().content) for i in range(N)]).T
        np.savez(outfile, case=case, training_set=np.array(ts), selected_indices=np.array(indices),
                 error_estimators=np.array([float(x) for x in rm.greedy_error_estimators]), N=N, M=Mr, A=A, F=F, Z=Z,
                 **{k: np.asarray(v) for k, v in cfg.items()})
        print(json.dumps({"N": int(N), "selected_indices": indices,
                          "error_estimators": [float(x) for x in rm.greedy_error_estimators],
                          "uses_plugin": bool(plugin) and hasattr(rm, "_b200_greedy_sweep")}))
        return
    raise SystemExit("unknown mode " + mode)


if __name__ == "__main__":
    main()
