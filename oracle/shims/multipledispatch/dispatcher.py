"""Dispatcher of the multipledispatch stand-in (the subset RBniCS subclasses: dispatch.py:76-190)."""
import inspect
from warnings import warn

from . import conflict
from .utils import expand_tuples, str_signature


class MDNotImplementedError(NotImplementedError):
    """A NotImplementedError for multiple dispatch"""


def ambiguity_warn(dispatcher, ambiguities):
    warn("Ambiguities exist in dispatched function %s: %s" % (dispatcher.name, ambiguities), conflict.AmbiguityWarning)


def halt_ordering():
    """Deprecated no-op kept for API compatibility."""


def restart_ordering(on_ambiguity=ambiguity_warn):
    """Deprecated no-op kept for API compatibility."""


class Dispatcher(object):
    """Dispatch methods based on type signature."""
    __slots__ = "__name__", "name", "funcs", "_ordering", "_cache", "doc"

    def __init__(self, name, doc=None):
        self.name = self.__name__ = name
        self.funcs = {}
        self.doc = doc
        self._cache = {}

    def register(self, *types, **kwargs):
        """Register a new implementation for the given types."""
        def _df(func):
            self.add(types, func, **kwargs)
            return func
        return _df

    @classmethod
    def get_func_params(cls, func):
        return inspect.signature(func).parameters.values()

    def add(self, signature, func):
        """Add a new types/method pair to the dispatcher."""
        if any(isinstance(typ, tuple) for typ in signature):
            for typs in expand_tuples(signature):
                self.add(typs, func)
            return
        for typ in signature:
            if not isinstance(typ, type):
                raise TypeError("Tried to dispatch on non-type: %s in signature <%s> of function %s"
                                % (typ, str_signature(signature), self.name))
        self.funcs[tuple(signature)] = func
        self._cache.clear()
        try:
            del self._ordering
        except AttributeError:
            pass

    @property
    def ordering(self):
        try:
            return self._ordering
        except AttributeError:
            return self.reorder()

    def reorder(self, on_ambiguity=ambiguity_warn):
        """Recompute the checking order of the signatures; report ambiguities through on_ambiguity."""
        self._ordering = od = conflict.ordering(self.funcs)
        amb = conflict.ambiguities(self.funcs)
        if amb:
            on_ambiguity(self, amb)
        return od

    def __call__(self, *args, **kwargs):
        types = tuple([type(arg) for arg in args])
        try:
            func = self._cache[types]
        except KeyError:
            func = self.dispatch(*types)
            if not func:
                raise NotImplementedError("Could not find signature for %s: <%s>" % (self.name, str_signature(types)))
            self._cache[types] = func
        return func(*args, **kwargs)

    def __str__(self):
        return "<dispatched %s>" % self.name
    __repr__ = __str__

    def dispatch(self, *types):
        """Determine the implementation for these types (None if there is none)."""
        if types in self.funcs:
            return self.funcs[types]
        try:
            return next(self.dispatch_iter(*types))
        except StopIteration:
            return None

    def dispatch_iter(self, *types):
        n = len(types)
        for signature in self.ordering:
            if len(signature) == n and all(map(issubclass, types, signature)):
                yield self.funcs[signature]

    def resolve(self, types):
        return self.dispatch(*types)

    def __getstate__(self):
        return {"name": self.name, "funcs": self.funcs}

    def __setstate__(self, d):
        self.name = d["name"]
        self.funcs = d["funcs"]
        self._ordering = conflict.ordering(self.funcs)
        self._cache = dict()

    @property
    def __doc__(self):
        docs = ["Multiply dispatched method: %s" % self.name]
        if self.doc:
            docs.append(self.doc)
        other = []
        for sig in self.ordering[::-1]:
            func = self.funcs[sig]
            if func.__doc__:
                s = "Inputs: <%s>\n" % str_signature(sig)
                s += "-" * len(s) + "\n"
                s += func.__doc__.strip()
                docs.append(s)
            else:
                other.append(str_signature(sig))
        if other:
            docs.append("Other signatures:\n    " + "\n    ".join(other))
        return "\n\n".join(docs)

    def _help(self, *args):
        return self.dispatch(*map(type, args)).__doc__

    def help(self, *args, **kwargs):
        print(self._help(*args))


class MethodDispatcher(Dispatcher):
    """Dispatch methods based on type signature (descriptor bound to an instance)."""
    __slots__ = ("obj", "cls")

    @classmethod
    def get_func_params(cls, func):
        return list(inspect.signature(func).parameters.values())[1:]

    def __get__(self, instance, owner):
        self.obj = instance
        self.cls = owner
        return self

    def __call__(self, *args, **kwargs):
        types = tuple([type(arg) for arg in args])
        func = self.dispatch(*types)
        if not func:
            raise NotImplementedError("Could not find signature for %s: <%s>" % (self.name, str_signature(types)))
        return func(self.obj, *args, **kwargs)
