"""The drop-in boundary, executed: `integration/rbnics/backends/online/b200` is imported by the
reference's OWN `set_online_backend` (rbnics/backends/online/__init__.py:19-53) under the reference's OWN dispatcher
(rbnics/utils/decorators/dispatch.py) -- the real RBniCS package from baseline/_ref or /root/reference with the
third-party stand-ins of oracle/shims.  Every backend needs its own interpreter (the choice is made at import), hence
the subprocesses (tests/refrun/run_reference.py).

CPU: name resolution (29 + 15 names), distinct container types, registrations in the generic dispatchers without
ambiguity.  GPU: the reference's EllipticCoercive[Compliant]RBReducedProblem.solve() / estimate_error() running on the
B200 backend reproduce the numpy backend's results (fixtures of tests/golden/make_reference_golden.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import reference_driver as R

HERE = os.path.dirname(os.path.abspath(__file__))
RUN = os.path.join(HERE, "refrun", "run_reference.py")
needs_reference = pytest.mark.skipif(R.reference_root() is None,
                                     reason="the reference is neither under baseline/_ref nor under /root/reference")


def _run(*args):
    out = subprocess.run([sys.executable, RUN] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@needs_reference
def test_set_online_backend_b200_resolves_every_name():
    d = _run("names", "b200")
    strip = lambda n: n[len("Online"):] if n.startswith("Online") else n[len("online_"):]  # noqa: E731
    assert sorted(strip(n) for n in d["online_all"]) == d["numpy_all"] and len(d["numpy_all"]) == 29
    assert d["wrapping_all"] == d["numpy_wrapping_all"] and len(d["wrapping_all"]) == 15
    assert d["wrapping_module"] == "rbnics.backends.online.b200.wrapping"
    hot = {"OnlineAffineExpansionStorage": "affine_expansion_storage", "OnlineFunction": "function",
           "OnlineLinearSolver": "linear_solver", "OnlineMatrix": "matrix", "online_product": "product",
           "online_sum": "sum", "online_transpose": "transpose", "OnlineVector": "vector"}
    for name, mod in hot.items():
        assert d["modules"][name] == "rbnics.backends.online.b200." + mod
    for name, mod in d["modules"].items():
        if name not in hot:
            assert mod.startswith("rbnics.backends.online.numpy."), (name, mod)
    assert d["distinct_types"] and d["subclass_of_numpy_types"]
    for name in ("product", "sum", "LinearSolver", "transpose"):
        mods = d["dispatch"][name]
        assert any(m.startswith("rbnics.backends.online.b200.") for m in mods)
        assert any(m.startswith("rbnics.backends.online.numpy.") for m in mods)   # both registered, no ambiguity
    for name in ("Matrix", "Vector"):
        assert d["dispatch"][name] == ["rbnics.backends.online.b200." + name.lower()]   # replaces= took effect


@needs_reference
def test_numpy_backend_is_untouched_without_the_plugin():
    d = _run("names", "numpy")
    assert all(m.startswith("rbnics.backends.online.numpy.") for m in d["modules"].values())


@needs_reference
def test_real_reference_sweep_reproduces_the_committed_fixture(tmp_path):
    """The fixtures are reproducible: the real package run again gives the committed numbers bit for bit."""
    g = np.load(os.path.join(HERE, "golden", "reference_real_sweep_n8_q3.npz"))
    out = tmp_path / "again.npz"
    _run("sweep", "numpy", 8, 3, 2, 40, 0, 0, out)
    h = np.load(out)
    for k in ("u", "delta", "Theta"):
        assert np.array_equal(g[k], h[k]), k
    assert (float(g["max"]), int(g["argmax"])) == (float(h["max"]), int(h["argmax"]))


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("case,args", [("n8_q3", (8, 3, 2, 40, 0, 0)), ("n20_q4", (20, 4, 3, 60, 3, 0)),
                                       ("n16_q5_compliant", (16, 5, 5, 40, 7, 1))])
def test_reference_reduced_problem_on_the_b200_backend(tmp_path, case, args):
    g = np.load(os.path.join(HERE, "golden", "reference_real_sweep_%s.npz" % case))
    out = tmp_path / "b200.npz"
    d = _run("sweep", "b200", *args, out)
    assert d["backend"] == "rbnics.backends.online.b200.matrix"
    h = np.load(out)
    assert int(h["argmax"]) == int(g["argmax"])
    assert np.max(np.abs(h["delta"] - g["delta"]) / np.abs(g["delta"])) <= 1e-10
    assert np.max(np.abs(h["u"] - g["u"])) <= 1e-10 * np.max(np.abs(g["u"]))


@needs_reference
def test_real_reference_pod_greedy_reproduces_the_committed_fixture(tmp_path):
    """The complete POD-Greedy offline stage of the real package (parabolic problem), run again, gives the committed
    selected indices and estimators."""
    g = np.load(os.path.join(HERE, "golden", "reference_real_pod_greedy_unsteady2.npz"))
    out = tmp_path / "again.npz"
    d = _run("pod_greedy", "numpy", "unsteady2", out)
    h = np.load(out)
    assert d["selected_indices"] == [int(i) for i in g["selected_indices"]]
    assert np.array_equal(g["error_estimators"], h["error_estimators"])
    assert np.array_equal(g["Z"], h["Z"])


@needs_reference
def test_real_reference_greedy_reproduces_the_committed_fixture(tmp_path):
    """ReducedBasis(problem).offline() of the unmodified reference, run again, gives the committed greedy sequence."""
    g = np.load(os.path.join(HERE, "golden", "reference_real_greedy_tutorial01.npz"))
    d = _run("greedy", "numpy", "tutorial01", 0, tmp_path / "again.npz")
    assert d["selected_indices"] == g["selected_indices"].tolist()
    assert d["error_estimators"] == g["error_estimators"].tolist()
    assert d["uses_plugin"] is False


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("case,backend", [("tutorial01", "numpy"), ("tutorial01", "b200"), ("stress3", "numpy"),
                                          ("config2", "numpy"), ("tutorial01_mesh", "numpy")])
def test_reference_offline_loop_with_the_b200_greedy_override(tmp_path, case, backend):
    """@B200Sweep() on the problem: RBniCS' own offline loop (truth solves, Gram-Schmidt, projections, Riesz products
    -- all the reference's code) with every `training_set.max(closure)` replaced by one batched sweep on the GPU selects
    the SAME parameters as the unmodified reference."""
    g = np.load(os.path.join(HERE, "golden", "reference_real_greedy_%s.npz" % case))
    d = _run("greedy", backend, case, 1, tmp_path / "plugin.npz")
    assert d["uses_plugin"] is True
    assert d["selected_indices"] == g["selected_indices"].tolist()
    est, ref = np.array(d["error_estimators"]), g["error_estimators"]
    # eps2 cancels down to ~1e-15 of its terms at the last iterations: absolute floor relative to the first estimator
    assert np.all(np.abs(est - ref) <= 1e-6 * ref + 1e-9 * ref[0])
    h = np.load(tmp_path / "plugin.npz")
    assert np.max(np.abs(h["A"] - g["A"])) <= 1e-10 * np.max(np.abs(g["A"]))


@needs_reference
def test_real_stokes_reduced_problem_reproduces_the_executed_source_fixture(tmp_path):
    """The reference's StokesRBReducedProblem (components u, s, p: component-dictionary slicing, block solve, 9 term pairs)
    run from the real package gives the fixture that round 1 produced by executing the reference's source: bit-identical."""
    g = np.load(os.path.join(HERE, "golden", "reference_stokes_n5.npz"))
    out = tmp_path / "stokes.npz"
    d = _run("stokes", "numpy", "n5", out)
    h = np.load(out)
    assert np.array_equal(h["u"], g["U"]) and np.array_equal(h["delta"], g["delta"])
    assert d["argmax"] == int(np.argmax(g["delta"]))


@pytest.mark.gpu
@needs_reference
def test_stokes_reduced_problem_on_the_b200_backend_and_plugin_sweep(tmp_path):
    """Same class with `online backend = b200`: the reference's own block slicing feeds rb_affine_combine / rb_solve_dense /
    rb_bilinear; and the Stokes family of the `_greedy` override returns the same (max, argmax) from ONE batched sweep."""
    g = np.load(os.path.join(HERE, "golden", "reference_stokes_n5.npz"))
    out = tmp_path / "stokes_b200.npz"
    d = _run("stokes", "b200", "n5", out, "plugin")
    assert d["backend"] == "rbnics.backends.online.b200.matrix"
    h = np.load(out)
    assert np.max(np.abs(h["u"] - g["U"])) <= 1e-10 * np.max(np.abs(g["U"]))
    assert np.max(np.abs(h["delta"] - g["delta"]) / g["delta"]) <= 1e-10
    assert d["argmax"] == int(np.argmax(g["delta"])) == d["plugin_argmax"]
    assert abs(d["plugin_max"] - float(g["delta"].max())) <= 1e-10 * float(g["delta"].max())


@needs_reference
def test_real_parabolic_reduced_problem_reproduces_the_executed_source_fixture(tmp_path):
    """ParabolicCoerciveRBReducedProblem of the real package (its numpy TimeStepping, per-step estimator with the m-terms,
    TimeQuadrature) against the round-1 fixture produced by executing the reference's source."""
    g = np.load(os.path.join(HERE, "golden", "reference_parabolic.npz"))
    out = tmp_path / "par.npz"
    d = _run("parabolic", "numpy", out)
    h = np.load(out)
    assert np.max(np.abs(h["delta"] - g["delta"]) / g["delta"]) <= 1e-14
    assert d["argmax"] == int(np.argmax(g["delta"]))


@pytest.mark.gpu
@needs_reference
def test_parabolic_family_of_the_plugin_sweep(tmp_path):
    """The POD-Greedy family of the `_greedy` override on the reference's own parabolic reduced problem: one batched sweep
    (K implicit-Euler steps, estimator per step, Simpson weights) == training_set.max of the reference's closure."""
    g = np.load(os.path.join(HERE, "golden", "reference_parabolic.npz"))
    d = _run("parabolic", "numpy", tmp_path / "par.npz", "plugin")
    assert d["plugin_argmax"] == d["argmax"] == int(np.argmax(g["delta"]))
    assert abs(d["plugin_max"] - d["max"]) <= 1e-10 * d["max"]


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("case", ["unsteady2", "unsteady3"])
def test_reference_pod_greedy_loop_with_the_b200_greedy_override(tmp_path, case):
    """BASELINE configs[3] inside the reference: RBniCS' own POD-Greedy offline loop (truth time stepping, trajectory POD,
    projections, Riesz products -- all the reference's code) with the closure of TimeDependentRBReduction._greedy replaced
    by the fused parabolic sweep on the GPU selects the SAME parameters as the unmodified reference."""
    g = np.load(os.path.join(HERE, "golden", "reference_real_pod_greedy_%s.npz" % case))
    d = _run("pod_greedy", "numpy", case, tmp_path / "plugin.npz", "plugin")
    assert d["uses_plugin"] is True
    assert d["selected_indices"] == g["selected_indices"].tolist()
    est, ref = np.array(d["error_estimators"]), g["error_estimators"]
    assert np.all(np.abs(est - ref) <= 1e-7 * ref[0])
    h = np.load(tmp_path / "plugin.npz")
    assert np.array_equal(h["Z"][:, :int(g["N1"])], g["Z"][:, :int(g["N1"])])   # same truth code, same first modes
