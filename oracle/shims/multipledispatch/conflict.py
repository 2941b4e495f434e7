"""Signature ordering of the multipledispatch stand-in.

`supercedes` and `consistent` are deliberately looked up in this module's globals at call time: RBniCS replaces them
with its own versions (rbnics/utils/decorators/dispatch.py:754,798) and expects `ordering` / `ambiguities` to use them."""
from .utils import _toposort, groupby


class AmbiguityWarning(Warning):
    pass


def supercedes(a, b):
    """a is consistent with and strictly more specific than (or equal to) b"""
    if len(a) != len(b):
        return False
    return all(issubclass(x, y) for x, y in zip(a, b))


def consistent(a, b):
    """some argument list could satisfy both signatures"""
    if len(a) != len(b):
        return False
    return all(issubclass(x, y) or issubclass(y, x) for x, y in zip(a, b))


def ambiguous(a, b):
    return consistent(a, b) and not (supercedes(a, b) or supercedes(b, a))


def ambiguities(signatures):
    signatures = list(map(tuple, signatures))
    return set((a, b) for a in signatures for b in signatures
               if hash(a) < hash(b) and ambiguous(a, b)
               and not any(supercedes(c, a) and supercedes(c, b) for c in signatures))


def super_signature(signatures):
    n = len(signatures[0])
    assert all(len(s) == n for s in signatures)
    return [max([type.mro(sig[i]) for sig in signatures], key=len)[0] for i in range(n)]


def edge(a, b, tie_breaker=hash):
    """a should be checked before b"""
    return supercedes(a, b) and (not supercedes(b, a) or tie_breaker(a) > tie_breaker(b))


def ordering(signatures):
    """Topological order of the signatures: more specific ones first."""
    signatures = list(map(tuple, signatures))
    edges = [(a, b) for a in signatures for b in signatures if edge(a, b)]
    edges = groupby(lambda x: x[0], edges)
    for s in signatures:
        if s not in edges:
            edges[s] = []
    edges = dict((k, [b for a, b in v]) for k, v in edges.items())
    return _toposort(edges)
