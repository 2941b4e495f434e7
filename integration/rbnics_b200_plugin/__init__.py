"""RBniCS-side plug-in of rbnics_b200: makes `rbnics.backends.online.b200` importable and selects it.

A RBniCS maintainer would simply copy `integration/rbnics/backends/online/b200/` next to
`rbnics/backends/online/numpy/`.  Without touching the RBniCS installation the same effect is obtained with

    import rbnics_b200_plugin
    rbnics_b200_plugin.register()          # BEFORE the first `import rbnics`

which installs a meta-path finder resolving `rbnics.backends.online.b200[.*]` to the files of this repository, and
(optionally) writes the `.rbnicsrc` that selects the backend (rbnics/utils/config/config.py:66-90 reads it at import;
rbnics/backends/online/__init__.py:19-53 then imports the backend and re-exports its names as OnlineXxx / online_xxx).
"""
import importlib.abc
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
BACKEND_DIR = os.path.join(os.path.dirname(_HERE), "rbnics", "backends", "online", "b200")
_PREFIX = "rbnics.backends.online.b200"


class _B200BackendFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != _PREFIX and not fullname.startswith(_PREFIX + "."):
            return None
        rel = fullname[len(_PREFIX):].lstrip(".").split(".") if fullname != _PREFIX else []
        base = os.path.join(BACKEND_DIR, *rel)
        if os.path.isdir(base):
            return importlib.util.spec_from_file_location(fullname, os.path.join(base, "__init__.py"),
                                                          submodule_search_locations=[base])
        if os.path.isfile(base + ".py"):
            return importlib.util.spec_from_file_location(fullname, base + ".py")
        return None


def register(write_rbnicsrc_to=None):
    """Install the finder (idempotent).  With `write_rbnicsrc_to=<directory>` also write the `.rbnicsrc` selecting the
    backend there (RBniCS reads the file from the script's / current directory and its parents)."""
    if "rbnics.backends.online" in sys.modules:
        raise RuntimeError("rbnics_b200_plugin.register() must run before rbnics is imported: consumers bind the "
                           "OnlineXxx names at import (SURVEY.md 8b)")
    if not any(isinstance(f, _B200BackendFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _B200BackendFinder())
    repo_root = os.path.dirname(os.path.dirname(_HERE))
    if repo_root not in sys.path:
        sys.path.append(repo_root)  # rbnics_b200 (ctypes layer over librbnics_b200.so)
    if write_rbnicsrc_to is not None:
        with open(os.path.join(str(write_rbnicsrc_to), ".rbnicsrc"), "w") as f:
            f.write("[backends]\nonline backend = b200\n")
