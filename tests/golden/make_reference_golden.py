#!/usr/bin/env python
"""Golden vectors from the REAL reference package (not from exec'd fragments like make_golden.py): RBniCS is imported
from baseline/_ref or /root/reference with the third-party stand-ins of oracle/shims (oracle/reference_driver.py) and
its own classes produce

  tests/golden/reference_real_sweep_<case>.npz   EllipticCoercive[Compliant]RBReducedProblem.solve()/estimate_error()
                                                 over a ParameterSpaceSubset, closure of RBReduction._greedy, max()
  tests/golden/reference_real_pod_galerkin.npz   the COMPLETE POD-Galerkin offline stage, PODGalerkin(problem).offline():
                                                 POD eigenvalues, retained energy, basis, reduced operators
  tests/golden/reference_real_pod_greedy_<case>.npz  the COMPLETE POD-Greedy offline stage of a parabolic problem
                                                 (ReducedBasis(ParabolicCoerciveProblem).offline(), POD_Greedy = N1)
  tests/golden/reference_real_greedy_<case>.npz  the COMPLETE offline stage, ReducedBasis(problem).offline(), on a dense
                                                 FEniCS-free truth problem (oracle/reference_truth.py): training set,
                                                 greedy-selected indices, error estimators, final reduced operators,
                                                 error-estimation operators and basis
  tests/golden/storage_real/                     AffineExpansionStorage.save() of the numpy online backend
                                                 (rbnics/backends/online/basic/affine_expansion_storage.py:56-153),
                                                 single- and multi-component (u, s, p) items, plus the reference's own
                                                 component-dict slices of the latter

Run in the build container:  python tests/golden/make_reference_golden.py
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
RUN = os.path.join(os.path.dirname(HERE), "refrun", "run_reference.py")

# name: N, Qa, Qf, B, seed, compliant
CASES = {
    "n8_q3": (8, 3, 2, 40, 0, 0),
    "n20_q4": (20, 4, 3, 60, 3, 0),
    "n33_q6": (33, 6, 2, 48, 5, 0),
    "n16_q5_compliant": (16, 5, 5, 40, 7, 1),
}


def main():
    for name, case in CASES.items():
        out = os.path.join(HERE, "reference_real_sweep_%s.npz" % name)
        subprocess.run([sys.executable, RUN, "sweep", "numpy"] + [str(x) for x in case] + [out], check=True)
    for case in ("tutorial01", "stress3", "config2", "tutorial01_mesh", "tutorial02_mesh"):
        out = os.path.join(HERE, "reference_real_greedy_%s.npz" % case)
        subprocess.run([sys.executable, RUN, "greedy", "numpy", case, "0", out], check=True)
    subprocess.run([sys.executable, RUN, "pod_galerkin", "numpy", os.path.join(HERE, "reference_real_pod_galerkin.npz")],
                   check=True)
    for case in ("unsteady2", "unsteady3"):
        out = os.path.join(HERE, "reference_real_pod_greedy_%s.npz" % case)
        subprocess.run([sys.executable, RUN, "pod_greedy", "numpy", case, out], check=True)
    sdir = os.path.join(HERE, "storage_real")
    shutil.rmtree(sdir, ignore_errors=True)
    os.makedirs(sdir)
    subprocess.run([sys.executable, RUN, "storage", "numpy", sdir], check=True)


if __name__ == "__main__":
    main()
