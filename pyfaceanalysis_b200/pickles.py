"""Loader for the reference's model store (Python-2 pickles of MDP / cuicuilco objects).

Replaces ``cuicuilco.object_cache.Cache.load_obj_from_cache(None, base_dir=..., base_filename=...)``
as it is called by ``load_networks_from_pipeline`` (reference ``face_analysis.py:457,485``): the file
is ``base_dir/base_filename + ".pckl"``; a network pickle may hold a tuple whose ``[0]`` is the flow
(``face_analysis.py:473-478``).

Neither ``mdp`` nor ``cuicuilco`` is importable here, so classes are materialised as
plain attribute bags that remember the pickled ``module.name``; module-level *functions* (the
``nonlinear_expansion`` functions a ``GeneralExpansionNode`` stores by name) become ``FuncRef`` objects.
The legacy top-level module aliases the reference registers before unpickling
(``FaceDetectUpdated.py:57-68``) are folded onto their cuicuilco names.
"""
from __future__ import annotations

import io
import os
import pickle

import numpy as np

# FaceDetectUpdated.py:57-68 -- old pickles name these modules without the "cuicuilco." prefix.
LEGACY_MODULE_ALIASES = {
    "more_nodes": "cuicuilco.more_nodes",
    "patch_mdp": "cuicuilco.patch_mdp",
    "sfa_libs": "cuicuilco.sfa_libs",
    "system_parameters": "cuicuilco.system_parameters",
    "network_builder": "cuicuilco.network_builder",
    "nonlinear_expansion": "cuicuilco.nonlinear_expansion",
    "gsfa_node": "cuicuilco.gsfa_node",
    "igsfa_node": "cuicuilco.igsfa_node",
    "GSFA_node": "cuicuilco.gsfa_node",
    "imageLoader": "cuicuilco.image_loader",
    "histogram_equalization": "cuicuilco.histogram_equalization",
    "inversion": "cuicuilco.inversion",
    "lattice": "cuicuilco.lattice",
}

_FUNCTION_MODULES = ("cuicuilco.nonlinear_expansion", "cuicuilco.sfa_libs")


class PickledObject(object):
    """Attribute bag standing in for an instance of an un-importable class."""

    _pickled_module = "?"
    _pickled_name = "?"

    def __init__(self, *args, **kwargs):
        self._init_args = args
        self.__dict__.update(kwargs)

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):
            # (dict_state, slots_state)
            if state[0]:
                self.__dict__.update(state[0])
            self.__dict__.update(state[1])
        elif isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state

    @classmethod
    def pickled_class(cls):
        return "%s.%s" % (cls._pickled_module, cls._pickled_name)

    def __repr__(self):
        keys = sorted(k for k in self.__dict__ if not k.startswith("__"))
        return "<%s %s>" % (self.pickled_class(), ", ".join(keys[:12]))


class FuncRef(object):
    """A module-level function pickled by reference (``GLOBAL module name``)."""

    def __init__(self, module, name):
        self.module = module
        self.name = name
        self.__name__ = name

    def __call__(self, *a, **k):
        raise RuntimeError("FuncRef %s.%s is a name only; resolve it through "
                           "pyfaceanalysis_b200.expansions" % (self.module, self.name))

    def __repr__(self):
        return "<FuncRef %s.%s>" % (self.module, self.name)

    def __eq__(self, other):
        return isinstance(other, FuncRef) and (self.module, self.name) == (other.module, other.name)

    def __hash__(self):
        return hash((self.module, self.name))


_class_cache = {}


def _stub_class(module, name):
    key = (module, name)
    cls = _class_cache.get(key)
    if cls is None:
        cls = type(str(name), (PickledObject,), {"_pickled_module": module, "_pickled_name": name})
        _class_cache[key] = cls
    return cls


def canonical_module(module):
    head = module.split(".")[0]
    if head in LEGACY_MODULE_ALIASES and not module.startswith("cuicuilco."):
        rest = module[len(head):]
        return LEGACY_MODULE_ALIASES[head] + rest
    return module


# Exact (module, name) pairs a model pickle may resolve outside mdp / cuicuilco.  Anything else -- eval, exec,
# getattr, __import__, os.system, numpy.testing helpers ... -- is refused: REDUCE on an arbitrary global is
# arbitrary code execution.
_SAFE_GLOBALS = {
    ("numpy.core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy", "float64"), ("numpy", "float32"), ("numpy", "int64"),
    ("numpy", "int32"), ("numpy", "uint8"), ("numpy", "bool_"), ("numpy", "complex128"),
    ("numpy.core.numeric", "_frombuffer"), ("numpy._core.numeric", "_frombuffer"),
    ("numpy.random", "__RandomState_ctor"), ("numpy.random.mtrand", "RandomState"),
    ("copy_reg", "_reconstructor"), ("copyreg", "_reconstructor"),
    ("__builtin__", "object"), ("__builtin__", "set"), ("__builtin__", "frozenset"), ("__builtin__", "slice"),
    ("__builtin__", "complex"), ("__builtin__", "bytearray"), ("__builtin__", "tuple"), ("__builtin__", "list"),
    ("__builtin__", "dict"), ("__builtin__", "long"), ("__builtin__", "int"), ("__builtin__", "float"),
    ("__builtin__", "bool"), ("__builtin__", "str"), ("__builtin__", "unicode"), ("__builtin__", "range"),
    ("__builtin__", "xrange"),
    ("builtins", "object"), ("builtins", "set"), ("builtins", "frozenset"), ("builtins", "slice"),
    ("builtins", "complex"), ("builtins", "bytearray"), ("builtins", "tuple"), ("builtins", "list"),
    ("builtins", "dict"), ("builtins", "int"), ("builtins", "float"), ("builtins", "bool"), ("builtins", "str"),
    ("builtins", "bytes"), ("builtins", "range"),
    ("collections", "OrderedDict"), ("collections", "defaultdict"),
    ("_codecs", "encode"),      # how Python 3 writes the bytes of a numpy array under protocol 2 (synthetic pickles)
}
_PY2_BUILTIN_NAMES = {"long": "int", "unicode": "str", "xrange": "range"}


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _SAFE_GLOBALS:
            if module.startswith("numpy.core"):
                module = module.replace("numpy.core", "numpy._core")
            elif module in ("__builtin__", "copy_reg"):
                module = {"__builtin__": "builtins", "copy_reg": "copyreg"}[module]
                name = _PY2_BUILTIN_NAMES.get(name, name)
            return super().find_class(module, name)
        module = canonical_module(module)
        top = module.split(".")[0]
        if top not in ("mdp", "cuicuilco", "bimdp"):
            raise pickle.UnpicklingError(
                "refusing to materialise %s.%s: only mdp / cuicuilco classes and an allow-list of numpy / builtin "
                "constructors are expected in a PyFaceAnalysis model pickle" % (module, name))
        if module in _FUNCTION_MODULES and name[:1].islower() or (
                module in _FUNCTION_MODULES and name in ("QT", "CT", "QE", "TE", "QN", "CN")):
            return FuncRef(module, name)
        return _stub_class(module, name)


def loads(data):
    """Unpickle ``bytes`` produced by Python 2 (protocol <= 2) or by :func:`dumps`."""
    return _StubUnpickler(io.BytesIO(data), encoding="latin1").load()


def load_obj(base_dir, base_filename, verbose=False):
    """Drop-in for ``Cache.load_obj_from_cache(None, base_dir=..., base_filename=...)``.

    ``"None0"`` (reference ``face_analysis.py:456``) yields ``None``.
    """
    if base_filename == "None0":
        return None
    path = os.path.join(base_dir, base_filename + ".pckl")
    if verbose:
        print("loading", path)
    with open(path, "rb") as f:
        return loads(f.read())


def class_path(obj):
    """``module.ClassName`` an object was pickled under (after alias folding)."""
    if isinstance(obj, PickledObject):
        return obj.pickled_class()
    return "%s.%s" % (type(obj).__module__, type(obj).__name__)


def new_object(module, name, **attrs):
    """Build an attribute bag of the given pickled class (used by the synthetic-flow factory so
    that synthetic and real flows go through one compiler)."""
    obj = _stub_class(canonical_module(module), name)()
    obj.__dict__.pop("_init_args", None)
    obj.__dict__.update(attrs)
    return obj


class _StubPickler(pickle._Pickler):
    """Protocol-2 pickler that writes attribute bags back as ``GLOBAL module name`` + ``NEWOBJ`` +
    ``BUILD(dict)`` and function references as bare ``GLOBAL``s -- the opcodes a Python-2 MDP pickle
    holds -- so a synthetic flow saved with :func:`dumps` has the shape of a real
    ``SavedNetworks/*.pckl`` and exercises the same loader path."""

    def __init__(self, f):
        super().__init__(f, protocol=2)

    def _global(self, module, name):
        self.write(pickle.GLOBAL + module.encode("ascii") + b"\n" + name.encode("ascii") + b"\n")

    def save(self, obj, save_persistent_id=True):
        if isinstance(obj, (FuncRef, PickledObject)):
            x = self.memo.get(id(obj))
            if x is not None:
                self.write(self.get(x[0]))
                return
            if isinstance(obj, FuncRef):
                self._global(obj.module, obj.name)
                self.memoize(obj)
                return
            self._global(type(obj)._pickled_module, type(obj)._pickled_name)
            self.write(pickle.EMPTY_TUPLE + pickle.NEWOBJ)
            self.memoize(obj)
            state = dict(obj.__dict__)
            state.pop("_init_args", None)
            self.save(state)
            self.write(pickle.BUILD)
            return
        super().save(obj, save_persistent_id)


def dumps(obj):
    buf = io.BytesIO()
    _StubPickler(buf).dump(obj)
    return buf.getvalue()
