"""CPU tests: the C-ABI library loads and exports every declared symbol; host-side logic (packing of the estimator
quadratic form, serial max-loc rule, cyclic sharding, the world_size-2 gloo path).  No compute calls without a GPU."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_of_the_header():
    from rbnics_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "rbnics_b200.h")).read()
    declared = set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", header)) - {"rb_ctx"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name)


def test_engine_fails_loudly_without_gpu():
    """No CPU fallback: creating an engine without a B200 raises (skipped on the GPU box)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rbnics_b200._lib import B200LibraryError
    from rbnics_b200.engine import SweepEngine
    with pytest.raises(B200LibraryError):
        SweepEngine(0)


def test_product_path_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under rbnics_b200/ may import it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rbnics_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_elliptic_quadratic_form_packing():
    """z^T G z with z = [theta_a (x) u ; theta_f] equals ff + 2 af + aa of elliptic_rb_reduced_problem.py:55-64."""
    from oracle import rb_oracle as O
    from rbnics_b200 import synthetic
    from rbnics_b200.engine import elliptic_quadratic_form
    N, Qa, Qf = 7, 3, 2
    prob = synthetic.elliptic_reduced_problem(N, Qa, Qf, seed=5, riesz_rank=8)
    G, (scale, src, length, voff, vlen) = elliptic_quadratic_form(prob["Gff"], prob["Gaf"], prob["Gaa"], Qa, Qf, N)
    assert (scale, src, length, voff, vlen) == ([0, 1, 2, -1], [0, 0, 0, N], [N, N, N, Qf], Qa, Qf)
    rng = np.random.default_rng(0)
    u = rng.standard_normal(N)
    tha = rng.uniform(0.1, 10, Qa)
    thf = rng.uniform(-1, 1, Qf)
    z = np.concatenate([np.kron(tha, u), thf])
    G_ff = [[prob["Gff"][i, j] for j in range(Qf)] for i in range(Qf)]
    G_af = [[prob["Gaf"][i, j] for j in range(Qf)] for i in range(Qa)]
    G_aa = [[prob["Gaa"][i, j] for j in range(Qa)] for i in range(Qa)]
    ref = O.residual_norm_squared(u, tuple(tha), tuple(thf), G_ff, G_af, G_aa, N)
    assert abs(z @ G @ z - ref) <= 1e-12 * abs(ref)
    G0, d0 = elliptic_quadratic_form(prob["Gff"], None, None, Qa, Qf, 0)
    assert G0.shape == (Qf, Qf) and d0[:3] == ([-1], [0], [Qf])


def test_maxloc_serial_rule():
    from rbnics_b200.sweep import cyclic_shard, maxloc_serial_rule
    assert maxloc_serial_rule([1.0, 3.0, 3.0], [5, 9, 2]) == (3.0, 2)            # lowest global index on ties
    v, i = maxloc_serial_rule([1.0, float("nan"), 7.0], [0, 4, 1])
    assert np.isnan(v) and i == 4                                                  # NaN wins (numpy.argmax)
    v, i = maxloc_serial_rule([float("nan"), float("nan")], [8, 3])
    assert i == 3
    assert list(cyclic_shard(10, 1, 4)) == [1, 5, 9]                               # parameter_space_subset.py:57


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
from rbnics_b200.sweep import DistributedMaxLoc, cyclic_shard
d = DistributedMaxLoc(backend="gloo")
vals = np.array([0.3, 0.9, 0.1, 0.9, 0.5, 0.2, 0.9, 0.05])     # global estimators; maxima at 1, 3, 6
mine = list(cyclic_shard(len(vals), d.rank, d.size))
loc = max(mine, key=lambda i: (vals[i], -i))
v, i = d.maxloc(vals[loc], loc)
assert (v, i) == (0.9, 1), (v, i)                                # serial rule: lowest global index
part = np.full((3, 3), float(d.rank + 1))
d.allreduce_sum_(part)
assert np.all(part == 3.0)
assert d.max_float(d.rank) == 1.0
# row-partitioned Gram matrix and Gram-Schmidt (SURVEY.md 8e): every rank works on its PETSc-style ownership range,
# one all-reduce per Gram matrix / per projection coefficient; equal to the single-rank result
import scipy.sparse as sp
from oracle import rb_oracle as O
from rbnics_b200 import offline
rng = np.random.default_rng(0)
Nh = 301
X = (sp.diags([-1.0, 2.5, -1.0], [-1, 0, 1], shape=(Nh, Nh)) + sp.diags([0.2, 0.2], [-9, 9], shape=(Nh, Nh))).tocsr()
S = rng.standard_normal((Nh, 6))
rb, re = offline.row_range(Nh, d.rank, d.size)
assert (rb, re) == ((0, 151) if d.rank == 0 else (151, 301))
part = S[rb:re].T @ (X[rb:re] @ S)                    # the rank's partial S_g^T (X S)_g
d.allreduce_sum_(part)
assert np.max(np.abs(part - O.gram_blas(S, X))) <= 1e-12 * np.max(np.abs(part))
Z = np.zeros((Nh, 0))
for k in range(5):
    v = rng.standard_normal(Nh)
    gather = lambda piece: d.allreduce_sum_(piece.copy())
    z, nrm = O.gram_schmidt_row_partitioned(v, Z, X, rb, re, d.sum_float, gather)
    z1 = O.gram_schmidt(v, Z, X)
    assert np.max(np.abs(z - z1)) <= 1e-12 * np.max(np.abs(z1)), (k, np.max(np.abs(z - z1)))
    Z = np.concatenate([Z, z1[:, None]], axis=1)
d.barrier()
d.close()
sys.stdout.write("rank %d ok\n" % d.rank); sys.stdout.flush()
"""


def test_world_size_2_gloo_maxloc_allreduce_and_row_partitioned_offline(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    import socket
    with socket.socket() as sock:   # a free port: a fixed one collides with concurrent runs on the same host
        sock.bind(("127.0.0.1", 0))
        port = str(sock.getsockname()[1])
    env = dict(os.environ)
    env.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": port})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", port, str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout
