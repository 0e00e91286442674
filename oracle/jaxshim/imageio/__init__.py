"""Inert stub (TEST INFRASTRUCTURE ONLY): satisfies an import-time dependency of the reference package."""
import sys, types


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        def _unavailable(*a, **k):
            raise RuntimeError(f"{self.__name__}.{name} is a stub (not available in this image)")
        return _unavailable


sys.modules[__name__].__class__ = _Stub
