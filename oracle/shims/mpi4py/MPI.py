"""`mpi4py.MPI` for a single rank: the calls RBniCS makes (rbnics/utils/mpi/parallel_max.py:11-35, parallel_io.py:11-40,
rbnics/utils/io/timer.py:17-28, rbnics/sampling/parameter_space_subset.py:57) reduce to identities."""
import copy
import time


class Op(object):
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "<shim MPI.Op %s>" % self.name


MAX, MIN, SUM, PROD, MAXLOC, MINLOC, LAND, LOR = (Op(n) for n in ("MAX", "MIN", "SUM", "PROD", "MAXLOC", "MINLOC",
                                                                  "LAND", "LOR"))
ANY_SOURCE, ANY_TAG = -1, -1


class Intracomm(object):
    rank = 0
    size = 1

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def barrier(self):
        pass

    Barrier = barrier

    def bcast(self, obj, root=0):
        assert root == 0
        return obj

    def allreduce(self, sendobj, op=SUM):
        return copy.copy(sendobj)

    def reduce(self, sendobj, op=SUM, root=0):
        return copy.copy(sendobj)

    def gather(self, sendobj, root=0):
        return [sendobj]

    def allgather(self, sendobj):
        return [sendobj]

    def scatter(self, sendobj, root=0):
        return sendobj[0]

    def Dup(self):
        return self

    def Free(self):
        pass

    def py2f(self):
        return 0


Comm = Intracomm
COMM_WORLD = Intracomm()
COMM_SELF = Intracomm()


def Wtime():
    return time.time()


def Is_initialized():
    return True


def Is_finalized():
    return False
