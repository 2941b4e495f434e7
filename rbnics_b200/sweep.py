"""Batched greedy sweep over a (sharded) training set: the B200 replacement of
``ParameterSpaceSubset.max(solve_and_estimate_error)``.

Reference call stack: rbnics/reduction_methods/base/rb_reduction.py:160-185 (_greedy) ->
rbnics/sampling/parameter_space_subset.py:52-77 (max: cyclic shard, loop, arg-max) ->
rbnics/utils/mpi/parallel_max.py:11-35 (max-loc over ranks).

Multi-GPU: one process per GPU, the training set is sharded cyclically exactly like the reference
(``range(rank, len, size)``), reduced operators are replicated, and the only collective is one all-gather of
16 bytes per rank (value, global index) -- NCCL over NVLink via ``torch.distributed`` (gloo in the CPU tests).
The contract is the SERIAL arg-max (lowest global index among equal maxima, NaN wins), not the reference's MPI
"highest rank wins" rule, so that the greedy sequence does not depend on the number of GPUs.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import _lib as L


def maxloc_serial_rule(values, indices):
    """Combine per-rank (max value, global index) pairs with numpy.argmax semantics over the whole set:
    NaN beats everything, then the larger value, then the LOWER global index."""
    best_v, best_i = None, None
    for v, i in zip(values, indices):
        v = float(v)
        i = int(i)
        if best_v is None:
            best_v, best_i = v, i
            continue
        vn, bn = math.isnan(v), math.isnan(best_v)
        if vn != bn:
            better = vn
        elif not vn and v != best_v:
            better = v > best_v
        else:
            better = i < best_i
        if better:
            best_v, best_i = v, i
    return best_v, best_i


def cyclic_shard(n, rank, size):
    """Indices owned by ``rank``: rbnics/sampling/parameter_space_subset.py:57."""
    return range(rank, n, size)


class DistributedMaxLoc:
    """torch.distributed plumbing for the one collective of the sweep (and the Gram all-reduce of the offline path).

    Reads RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the environment (torchrun)."""

    def __init__(self, local_rank=0, backend=None):
        import torch
        import torch.distributed as dist
        self.torch = torch
        self.dist = dist
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        self.backend = backend
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            self.device = torch.device("cuda", local_rank)
        else:
            self.device = torch.device("cpu")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        self._own = not dist.is_initialized()
        if self._own:
            dist.init_process_group(backend=backend)
        self.rank = dist.get_rank()
        self.size = dist.get_world_size()

    def barrier(self):
        if self.backend == "nccl":
            self.torch.cuda.synchronize()
        self.dist.barrier()
        if self.backend == "nccl":
            self.torch.cuda.synchronize()

    def maxloc(self, value, index):
        """All-gather (value, index) from every rank and apply the serial rule.  16 bytes per rank."""
        t = self.torch
        mine = t.tensor([float(value), 0.0], dtype=t.float64, device=self.device)
        # the index travels bit-exactly as an int64 viewed as float64
        mine[1:2].view(t.int64)[0] = int(index)
        out = t.empty(2 * self.size, dtype=t.float64, device=self.device)
        self.dist.all_gather_into_tensor(out, mine)
        out = out.cpu()
        vals = out[0::2].tolist()
        idxs = out[1::2].clone().view(t.int64).tolist()
        return maxloc_serial_rule(vals, idxs)

    def maxloc_device(self, engine):
        """The same collective fed straight from the engine's device-resident result {value f64, global index i64,
        number of non-finite estimators i64} (rb_result_device_ptrs): NCCL all-gathers 24 bytes per rank out of HBM, no
        host staging of the send buffer.  Returns (value, index, total non-finite count)."""
        if self.backend != "nccl":
            raise RuntimeError("maxloc_device needs the NCCL backend")
        t = self.torch
        ptr, _ = engine.result_device_ptrs()

        class _Holder:  # zero-copy view of library-owned device memory (CUDA array interface v3)
            __cuda_array_interface__ = {"shape": (3,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}

        mine = t.as_tensor(_Holder(), device=self.device)
        out = t.empty(3 * self.size, dtype=t.float64, device=self.device)
        self.dist.all_gather_into_tensor(out, mine)
        out = out.cpu()
        vals = out[0::3].tolist()
        idxs = out[1::3].clone().view(t.int64).tolist()
        bad = int(out[2::3].clone().view(t.int64).sum().item())
        v, i = maxloc_serial_rule(vals, idxs)
        return v, i, bad

    def max_float(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_float(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def allreduce_sum_(self, array):
        """In-place SUM all-reduce of a numpy float64 array (partial Gram / projection blocks)."""
        t = self.torch.from_numpy(array)
        if self.backend == "nccl":
            d = t.to(self.device)
            self.dist.all_reduce(d, op=self.dist.ReduceOp.SUM)
            t.copy_(d.cpu())
        else:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return array

    def device_allreduce_callback(self):
        """`rb_allreduce_fn` for rb_set_allreduce: in-place SUM all-reduce (NCCL) of a device buffer the library hands
        over by pointer -- zero-copy through the CUDA array interface, no host staging."""
        if self.backend != "nccl":
            raise RuntimeError("the device all-reduce callback needs the NCCL backend")
        t, dist, device = self.torch, self.dist, self.device

        def allreduce(user, ptr, count):
            try:
                class _Holder:
                    __cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                                "version": 3}
                buf = t.as_tensor(_Holder(), device=device)
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)
                t.cuda.synchronize(device)
                return 0
            except Exception:  # an exception must not unwind through the C frames
                import traceback
                traceback.print_exc()
                return 1

        return L.ALLREDUCE_FN(allreduce)

    def close(self):
        if self._own and self.dist.is_initialized():
            self.dist.destroy_process_group()


class TrainingSetSweep:
    """``training_set.max(...)`` for a whole training set at once.

    The caller evaluates theta(mu) / beta(mu) for every mu ONCE (vectorised ``compute_theta``; the thetas do not
    change during the greedy loop) and registers them; every greedy iteration then uploads the new reduced
    operators and calls :meth:`max`.
    """

    def __init__(self, engine, Theta, beta, ThetaBC=None, dist=None):
        self.engine = engine
        self.dist = dist
        rank, size = (dist.rank, dist.size) if dist is not None else (0, 1)
        self.rank, self.size = rank, size
        Theta = np.ascontiguousarray(Theta, dtype=np.float64)
        self.n_global = Theta.shape[0]
        sl = slice(rank, None, size)
        self._Theta = np.ascontiguousarray(Theta[sl])
        self._beta = np.ascontiguousarray(np.asarray(beta, dtype=np.float64)[sl])
        self._ThetaBC = None if ThetaBC is None else np.ascontiguousarray(np.asarray(ThetaBC, dtype=np.float64)[sl])
        self._resident = False

    def _ensure_resident(self):
        if not self._resident:
            self.engine.set_training_set(self._Theta, self._beta, self._ThetaBC, index_offset=self.rank,
                                         index_stride=self.size)
            self._resident = True

    def max(self, kind=L.RB_EST_SQRT_EPS2_OVER_BETA, want_delta=False):
        """(max estimator, global arg-max index[, local deltas]) -- the return value of ParameterSpaceSubset.max."""
        self._ensure_resident()
        if self.dist is None:
            r = self.engine.sweep_resident(kind=kind, want_delta=want_delta)
            v, i = r["max"], r["argmax"]
        else:
            # a rank whose shard holds a singular reduced matrix must not raise before the collective (the other ranks
            # would wait in the all-gather for ever): gather first -- NaN wins the arg-max anyway -- then raise everywhere
            r = self.engine.sweep_resident(kind=kind, want_delta=want_delta, raise_on_singular=False)
            if self.dist.backend == "nccl":
                v, i, bad = self.dist.maxloc_device(self.engine)
            else:
                v, i = self.dist.maxloc(r["max"], r["argmax"])
                bad = int(self.dist.sum_float(float(math.isnan(r["max"]) or math.isinf(r["max"]))))
            if bad:
                raise L.SingularReducedMatrixError(
                    "rb_sweep: non-finite error estimators on at least one rank (singular reduced matrix or beta == 0); "
                    "numpy.linalg.solve would raise LinAlgError")
        if want_delta:
            return v, i, r["delta"]
        return v, i


class TestingSetSweep(TrainingSetSweep):
    """Reduced solutions, error estimators and outputs for a whole testing set: the reduced-order half of
    ``error_analysis`` / ``speedup_analysis`` (rbnics/reduction_methods/base/rb_reduction.py:312-390,478-580;
    the truth solves they compare against stay with the reference)."""

    __test__ = False  # not a pytest class

    def evaluate(self, kind=L.RB_EST_SQRT_EPS2_OVER_BETA, S=None, ThetaS=None):
        """-> dict(u (local B x N), delta, eps2[, output])  for this rank's shard (rows rank::size)."""
        self._ensure_resident()
        r = self.engine.sweep_resident(kind=kind, want_delta=True, want_u=True, want_eps2=True)
        out = {"u": r["u"], "delta": r["delta"], "eps2": r["eps2"]}
        if S is not None:
            ThetaS = np.ascontiguousarray(np.asarray(ThetaS, dtype=np.float64)[self.rank::self.size])
            out["output"] = self.engine.compute_outputs(S, ThetaS)
        return out
