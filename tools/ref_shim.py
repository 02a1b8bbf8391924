"""Import the UNMODIFIED reference (NeuroDiffGym/neurodiffeq) from /root/reference in THIS container.

Test/fixture infrastructure only.  The reference needs matplotlib / seaborn / ordered_set, none of which is
installed here (no network), so stub modules are injected into ``sys.modules`` first.  Nothing from the reference
is copied into this repository: the package is imported in place, used to generate golden vectors
(``tests/golden/generate.py``) and to pin ``oracle/``.  ``/root/reference`` does not exist on the GPU box, so
nothing under ``tests -m gpu``, ``bench.py`` or ``__graft_entry__.smoke()`` may import this module.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PINNJET_REFERENCE_ROOT", "/root/reference")


class _Anything:
    """Attribute sink: any attribute access / call returns another sink (enough for module-level plotting code)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


def _stub_module(name):
    mod = types.ModuleType(name)
    mod.__path__ = []  # behave like a package so that submodule imports resolve

    def _getattr(attr):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        return _Anything()

    mod.__getattr__ = _getattr
    return mod


class _OrderedSet(dict):
    """Order-preserving de-duplicating container: the only behaviour solvers.py:182 relies on."""

    def __init__(self, iterable=()):
        super().__init__()
        for item in iterable:
            self[item] = None

    def add(self, item):
        self[item] = None


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "neurodiffeq"))


def import_reference():
    """Returns the reference ``neurodiffeq`` package, forced onto CPU/float64 (its import default)."""
    if "neurodiffeq" in sys.modules and getattr(sys.modules["neurodiffeq"], "_pinnjet_shimmed", False):
        return sys.modules["neurodiffeq"]
    if not reference_available():
        raise ImportError(f"reference not found under {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.tri", "matplotlib.cm", "matplotlib.animation",
                 "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.mplot3d", "seaborn"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _stub_module(name)
    if "ordered_set" not in sys.modules:
        try:
            __import__("ordered_set")
        except Exception:
            mod = types.ModuleType("ordered_set")
            mod.OrderedSet = _OrderedSet
            sys.modules["ordered_set"] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import neurodiffeq  # noqa: E402  (the reference)
    from neurodiffeq.utils import set_tensor_type
    set_tensor_type("cpu", 64)
    neurodiffeq._pinnjet_shimmed = True
    return neurodiffeq
