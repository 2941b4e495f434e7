#!/usr/bin/env python
"""Subprocess entry of the tests that execute the REAL reference package (oracle/reference_driver.py): the online
backend of RBniCS is chosen at import time, so every backend gets its own interpreter.

    python tests/refrun/run_reference.py names   <backend>           -> JSON: exported names, type modules, dispatch tables
    python tests/refrun/run_reference.py sweep   <backend> <N> <Qa> <Qf> <B> <seed> <compliant> <out.npz>
    python tests/refrun/run_reference.py storage <backend> <directory>   -> writes storages with the reference's own save()
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_driver as R  # noqa: E402


def main():
    mode, backend = sys.argv[1], sys.argv[2]
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
    raise SystemExit("unknown mode " + mode)


if __name__ == "__main__":
    main()
