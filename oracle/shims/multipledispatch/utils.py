"""Helpers of the multipledispatch stand-in: Kahn topological sort, groupby, signature printing."""
from collections import OrderedDict


def raises(err, lamda):
    try:
        lamda()
        return False
    except err:
        return True


def expand_tuples(L):
    """expand_tuples([1, (2, 3)]) -> [(1, 2), (1, 3)]"""
    if not L:
        return [()]
    rest = expand_tuples(L[1:])
    if not isinstance(L[0], tuple):
        return [(L[0],) + t for t in rest]
    return [(item,) + t for t in rest for item in L[0]]


def reverse_dict(d):
    result = OrderedDict()
    for key in d:
        for val in d[key]:
            result[val] = result.get(val, tuple()) + (key,)
    return result


def _toposort(edges):
    """Topological sort (Kahn): edges = {node: [nodes that depend on it]} -> list of nodes, dependencies first."""
    incoming = reverse_dict(edges)
    incoming = OrderedDict((k, set(v)) for k, v in incoming.items())
    S = OrderedDict.fromkeys(v for v in edges if v not in incoming)
    L = []
    while S:
        n, _ = S.popitem()
        L.append(n)
        for m in edges.get(n, ()):
            assert n in incoming[m]
            incoming[m].remove(n)
            if not incoming[m]:
                S[m] = None
    if any(incoming.get(v, None) for v in edges):
        raise ValueError("Input has cycles")
    return L


def groupby(func, seq):
    d = OrderedDict()
    for item in seq:
        d.setdefault(func(item), []).append(item)
    return d


def typename(type_):
    try:
        return type_.__name__
    except AttributeError:
        if len(type_) == 1:
            return typename(*type_)
        return "(%s)" % ", ".join(map(typename, type_))


def str_signature(sig):
    return ", ".join(typename(c) for c in sig)
