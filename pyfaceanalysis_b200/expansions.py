"""Expansion-function vocabulary of cuicuilco ``GeneralExpansionNode`` lowered to a term table.

A ``GeneralExpansionNode`` stores ``funcs``: module-level functions of ``cuicuilco.nonlinear_expansion``
pickled *by name* (the reason for the alias at reference ``FaceDetectUpdated.py:62``); ``execute``
concatenates ``f_k(x)`` column-wise.  The cuicuilco source is not
available: the arithmetic of every name below is the restatement this build treats as
normative, and an unknown name raises with the name in the message.

On the GPU an expansion is not a list of functions but a flat **term table**: output column ``e`` is
``op_e(x[i_e], x[j_e]; p_e)`` with ``op`` one of the ``OP_*`` codes.  ``lower(funcs, d)`` returns that
table; the CUDA chunk builder (``csrc/flow.cu``) evaluates it row by row.
"""
from __future__ import annotations

import re

import numpy as np

# op codes shared with csrc/flow.cu (enum TermOp)
OP_ID = 0        # x_i
OP_MUL = 1       # x_i * x_j
OP_ABSPOW = 2    # |x_i| ** p
OP_SGNPOW = 3    # sign(x_i) * |x_i| ** p
OP_MUL3 = 4      # x_i * x_j * x_k   (k carried in the p slot as an integer)
OP_ABS = 5       # |x_i|
OP_CLIP = 6      # min(max(x_i, -p), p)

TERM_DTYPE = np.dtype([("op", "<i4"), ("i", "<i4"), ("j", "<i4"), ("p", "<f4")])


def _ident(d):
    return [(OP_ID, i, 0, 0.0) for i in range(d)]


def _qt(d):
    # all products x_i x_j, i <= j, i-major (cuicuilco pair_prod_ex ordering)
    return [(OP_MUL, i, j, 0.0) for i in range(d) for j in range(i, d)]


def _ct(d):
    return [(OP_MUL3, i, j, float(k)) for i in range(d) for j in range(i, d) for k in range(j, d)]


def _abspow(p):
    return lambda d: [(OP_ABSPOW, i, 0, p) for i in range(d)]


def _sgnpow(p):
    return lambda d: [(OP_SGNPOW, i, 0, p) for i in range(d)]


def _adj(max_shift):
    def f(d):
        out = []
        for s in range(1, max_shift + 1):
            out += [(OP_MUL, i, i + s, 0.0) for i in range(d - s)]
        return out
    return f


_BASE = {
    "identity": _ident,
    "I": _ident,
    "QT": _qt,
    "pair_prod_ex": _qt,
    "QE": lambda d: _ident(d) + _qt(d),
    "CT": _ct,
    "TE": lambda d: _ident(d) + _qt(d) + _ct(d),
    "unsigned_08expo": _abspow(0.8),
    "unsigned_06expo": _abspow(0.6),
    "unsigned_04expo": _abspow(0.4),
    "unsigned_sqrt": _abspow(0.5),
    "signed_08expo": _sgnpow(0.8),
    "signed_06expo": _sgnpow(0.6),
    "signed_04expo": _sgnpow(0.4),
    "signed_sqrt": _sgnpow(0.5),
    "abs": lambda d: [(OP_ABS, i, 0, 0.0) for i in range(d)],
    "pair_prod_adj1_ex": _adj(1),
    "pair_prod_adj2_ex": _adj(2),
    "pair_prod_adj3_ex": _adj(3),
}

# short suffixes usable behind a prefix selector: s15QT, s10u08ex, ...
_SUFFIX = {
    "QT": "QT", "CT": "CT", "QE": "QE", "TE": "TE", "I": "identity",
    "u08ex": "unsigned_08expo", "u06ex": "unsigned_06expo", "u04ex": "unsigned_04expo",
    "s08ex": "signed_08expo", "s06ex": "signed_06expo", "s04ex": "signed_04expo",
    "usqrt": "unsigned_sqrt", "ssqrt": "signed_sqrt",
}

_SEL_RE = re.compile(r"^s(\d+)_?([A-Za-z0-9_]+)$")
_EXPO_RE = re.compile(r"^(un)?signed_(\d)(\d)expo$")
_CLIP_RE = re.compile(r"^clip_?(\d+(?:p\d+)?)$")


def clip_limit(name):
    """L of a ``clip_<L>`` function name, else None."""
    m = _CLIP_RE.match(name)
    return float(m.group(1).replace("p", ".")) if m else None


def func_name(f):
    """Name of an expansion function however it is represented (FuncRef, python function, str)."""
    if isinstance(f, str):
        return f
    name = getattr(f, "name", None) or getattr(f, "__name__", None)
    if name is None:
        raise ValueError("expansion function without a name: %r" % (f,))
    return name


def terms_for(name, d):
    """Term list of one function applied to a ``d``-dimensional input."""
    if name in _BASE:
        return _BASE[name](d)
    m = _EXPO_RE.match(name)
    if m:
        # cuicuilco names carry the exponent as "0Y" = 0.Y (unsigned_08expo = |x|^0.8); other forms are refused
        if m.group(2) != "0":
            raise KeyError("nonlinear_expansion function %r: exponent form %s%s is not 0Y (register it in "
                           "pyfaceanalysis_b200/expansions.py)" % (name, m.group(2), m.group(3)))
        p = int(m.group(3)) / 10.0
        return (_abspow(p) if m.group(1) else _sgnpow(p))(d)
    m = _CLIP_RE.match(name)
    if m:
        p = float(m.group(1).replace("p", "."))
        return [(OP_CLIP, i, 0, p) for i in range(d)]
    m = _SEL_RE.match(name)
    if m:
        k = int(m.group(1))
        inner = _SUFFIX.get(m.group(2), m.group(2))
        if inner in _BASE or _EXPO_RE.match(inner):
            if k > d:
                raise ValueError("expansion %r selects the first %d components of a %d-dim input"
                                 % (name, k, d))
            return terms_for(inner, k)
    raise KeyError("unknown nonlinear_expansion function %r (register it in "
                   "pyfaceanalysis_b200/expansions.py)" % (name,))


def lower(funcs, d):
    """Concatenated term table (structured array, ``TERM_DTYPE``) of ``funcs`` on a ``d``-dim input."""
    terms = []
    for f in funcs:
        terms += terms_for(func_name(f), d)
    arr = np.zeros(len(terms), dtype=TERM_DTYPE)
    for e, (op, i, j, p) in enumerate(terms):
        arr[e] = (op, i, j, p)
    return arr


def expanded_dim(funcs, d):
    return sum(len(terms_for(func_name(f), d)) for f in funcs)


def term_flops(terms):
    """Algorithmic flop count of evaluating a term table once (QT one mul per term,
    pow counted as 1, identity free)."""
    ops = terms["op"]
    return int(np.sum(ops == OP_MUL) + 2 * np.sum(ops == OP_MUL3) + np.sum(ops == OP_ABSPOW)
               + np.sum(ops == OP_SGNPOW) + np.sum(ops == OP_ABS) + np.sum(ops == OP_CLIP))
