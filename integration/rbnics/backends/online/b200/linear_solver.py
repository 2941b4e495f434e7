"""B200 online backend: LinearSolver.  Construction, lifting rows (DirichletBC) and the monitor are the reference's
(rbnics/backends/online/basic/linear_solver.py:19-71); solve() replaces numpy.linalg.solve
(rbnics/backends/online/numpy/linear_solver.py:29-33) by rb_solve_dense (LU with partial pivoting on the device)."""
from rbnics.backends.abstract import LinearProblemWrapper
from rbnics.backends.online.numpy.linear_solver import LinearSolver as NumpyLinearSolver
from rbnics.backends.online.b200.function import Function
from rbnics.backends.online.b200.matrix import Matrix
from rbnics.backends.online.b200.vector import Vector
from rbnics.utils.decorators import BackendFor, DictOfThetaType, ThetaType


@BackendFor("b200", inputs=((Matrix.Type(), LinearProblemWrapper), Function.Type(), (Vector.Type(), None),
                            ThetaType + DictOfThetaType + (None,)))
class LinearSolver(NumpyLinearSolver):
    def set_parameters(self, parameters):
        assert len(parameters) == 0, "B200 linear solver does not accept parameters yet"

    def solve(self):
        from rbnics_b200 import online as _dev
        # the lifting rows have been applied to self.lhs / self.rhs by the constructor, as in the reference
        self.solution.vector().content[:] = _dev.solve_dense(_dev.get_engine(), self.lhs.content, self.rhs.content)
        if self.monitor is not None:
            self.monitor(self.solution)
