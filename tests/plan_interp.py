"""Test helper: executes a compiled ``PlanSpec`` with numpy, mirroring ``csrc/flow.cu`` step by step.

This is a *checker of the lowering* (gather composition, mean folding, iGSFA two-pass / folded
algebra, padding, row allocation): it lets the CPU test-suite compare ``plan.compile_flow`` against
the oracle without a GPU.  It lives under ``tests/`` and is never imported by the package.
"""
import numpy as np

from pyfaceanalysis_b200 import expansions as ex
from pyfaceanalysis_b200.plan import DST_GLOBAL, DST_ROWS


def _eval_terms(terms, src):
    """src: (n, n_sources) -> (n, K)"""
    out = np.zeros((src.shape[0], len(terms)), dtype=src.dtype)
    for e in range(len(terms)):
        op, i, j, p = int(terms["op"][e]), int(terms["i"][e]), int(terms["j"][e]), terms["p"][e]
        a = src[:, i]
        if op == ex.OP_ID:
            out[:, e] = a
        elif op == ex.OP_MUL:
            out[:, e] = a * src[:, j]
        elif op == ex.OP_ABSPOW:
            out[:, e] = np.abs(a) ** src.dtype.type(p)
        elif op == ex.OP_SGNPOW:
            out[:, e] = np.sign(a) * np.abs(a) ** src.dtype.type(p)
        elif op == ex.OP_MUL3:
            out[:, e] = a * src[:, j] * src[:, int(p)]
        elif op == ex.OP_ABS:
            out[:, e] = np.abs(a)
        elif op == ex.OP_CLIP:
            out[:, e] = np.clip(a, -p, p)
        else:
            raise ValueError(op)
    return out


def run_op(op, x, dtype=np.float64):
    n = x.shape[0]
    y = np.zeros((n, op.out_dim), dtype=dtype)
    for nd in range(op.n_nodes):
        w_i = 0 if op.shared else nd
        x0 = x[:, op.gather[nd]].astype(dtype) - op.in_offset[w_i].astype(dtype)
        rows = np.zeros((n, op.n_rows), dtype=dtype)
        for ps in op.passes:
            src = np.concatenate([x0, rows], axis=1)
            A = _eval_terms(ps["terms"], src)
            Y = A @ ps["W"][w_i].astype(dtype) + ps["b"][w_i].astype(dtype)
            if ps["dst"] & DST_ROWS:
                rows[:, ps["row0"]:ps["row0"] + Y.shape[1]] = Y
            if ps["dst"] & DST_GLOBAL:
                nv = int(ps["n_valid"][nd])
                c0 = int(op.out_col[nd]) + int(ps["col_off"][nd])
                y[:, c0:c0 + nv] = np.clip(Y[:, :nv], op.clip[0], op.clip[1])
    return y


def run_plan(spec, x, dtype=np.float64):
    x = np.asarray(x)
    for op in spec.ops:
        x = run_op(op, x, dtype)
    return x[:, :spec.output_dim]
