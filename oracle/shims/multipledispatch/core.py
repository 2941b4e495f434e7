"""`dispatch` decorator and `ismethod` of the multipledispatch stand-in."""
import inspect

from .dispatcher import Dispatcher, MethodDispatcher

global_namespace = dict()


def dispatch(*types, **kwargs):
    """Dispatch function on the types of the inputs.

    Collects implementations based on the function name (ignoring namespaces); raises NotImplementedError at call
    time if no implementation matches.
    """
    namespace = kwargs.get("namespace", global_namespace)
    types = tuple(types)

    def _df(func):
        name = func.__name__
        if ismethod(func):
            dispatcher = inspect.currentframe().f_back.f_locals.get(name, MethodDispatcher(name))
        else:
            if name not in namespace:
                namespace[name] = Dispatcher(name)
            dispatcher = namespace[name]
        dispatcher.add(types, func)
        return dispatcher
    return _df


def ismethod(func):
    """Is func a method?  (It has not been bound yet, so: is its first argument called `self`?)"""
    if hasattr(inspect, "signature"):
        signature = inspect.signature(func)
        return signature.parameters.get("self", None) is not None
    spec = inspect.getfullargspec(func)
    return spec and spec.args and spec.args[0] == "self"
