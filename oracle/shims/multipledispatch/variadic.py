"""Variadic signatures are not used by RBniCS; present so that `import multipledispatch.variadic` succeeds."""


class VariadicSignatureType(type):
    pass


def isvariadic(obj):
    return isinstance(obj, VariadicSignatureType)
