"""Oracle (test infrastructure): numpy float64 restatement of cuicuilco ``nonlinear_expansion``.

cuicuilco @9bfd24201b0e4107b9689c13b2da55e3a01cfb55 is not vendored -- PARITY UNPINNED;
each function restates the published behaviour of the name it carries.  Written as direct array
functions (the way the reference evaluates them), independently of the product's term tables.
"""
import re

import numpy as np


def identity(x):
    return x


def pair_prod_ex(x):
    """All products x_i * x_j with i <= j, i-major ("QT": quadratic terms)."""
    n, d = x.shape
    cols = []
    for i in range(d):
        cols.append(x[:, i:i + 1] * x[:, i:])
    if not cols:
        return np.zeros((n, 0))
    return np.concatenate(cols, axis=1)


QT = pair_prod_ex


def QE(x):
    return np.concatenate((x, pair_prod_ex(x)), axis=1)


def CT(x):
    """All products x_i x_j x_k with i <= j <= k."""
    n, d = x.shape
    cols = []
    for i in range(d):
        for j in range(i, d):
            cols.append(x[:, i:i + 1] * x[:, j:j + 1] * x[:, j:])
    if not cols:
        return np.zeros((n, 0))
    return np.concatenate(cols, axis=1)


def TE(x):
    return np.concatenate((x, pair_prod_ex(x), CT(x)), axis=1)


def _unsigned_expo(p):
    def f(x):
        return np.abs(x) ** p
    return f


def _signed_expo(p):
    def f(x):
        return np.sign(x) * np.abs(x) ** p
    return f


def _adj(max_shift):
    def f(x):
        cols = [x[:, :-s] * x[:, s:] for s in range(1, max_shift + 1)]
        return np.concatenate(cols, axis=1)
    return f


FUNCS = {
    "identity": identity, "I": identity,
    "QT": QT, "pair_prod_ex": pair_prod_ex, "QE": QE, "CT": CT, "TE": TE,
    "unsigned_08expo": _unsigned_expo(0.8), "unsigned_06expo": _unsigned_expo(0.6),
    "unsigned_04expo": _unsigned_expo(0.4), "unsigned_sqrt": _unsigned_expo(0.5),
    "signed_08expo": _signed_expo(0.8), "signed_06expo": _signed_expo(0.6),
    "signed_04expo": _signed_expo(0.4), "signed_sqrt": _signed_expo(0.5),
    "abs": np.abs,
    "pair_prod_adj1_ex": _adj(1), "pair_prod_adj2_ex": _adj(2), "pair_prod_adj3_ex": _adj(3),
}

_SHORT = {"u08ex": "unsigned_08expo", "u06ex": "unsigned_06expo", "u04ex": "unsigned_04expo",
          "s08ex": "signed_08expo", "s06ex": "signed_06expo", "s04ex": "signed_04expo",
          "usqrt": "unsigned_sqrt", "ssqrt": "signed_sqrt"}


def resolve(f):
    name = f if isinstance(f, str) else (getattr(f, "name", None) or f.__name__)
    if name in FUNCS:
        return FUNCS[name]
    m = re.match(r"^(un)?signed_0(\d)expo$", name)
    if m:
        p = int(m.group(2)) / 10.0
        return _unsigned_expo(p) if m.group(1) else _signed_expo(p)
    m = re.match(r"^clip_?(\d+(?:p\d+)?)$", name)
    if m:
        lim = float(m.group(1).replace("p", "."))
        return lambda x: np.clip(x, -lim, lim)
    m = re.match(r"^s(\d+)_?([A-Za-z0-9_]+)$", name)
    if m:
        k = int(m.group(1))
        inner = resolve(_SHORT.get(m.group(2), m.group(2)))
        return lambda x: inner(x[:, :k])
    raise KeyError("oracle: unknown nonlinear_expansion function %r" % (name,))
