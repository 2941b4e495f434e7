"""One-rank stand-in for `mpi4py` (see ../README.md).  Test infrastructure only."""
__version__ = "0.0+rbnics_b200_shim"
