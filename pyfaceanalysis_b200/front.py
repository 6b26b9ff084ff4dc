"""Vertical fusion of the first three layers of a thin fan-in-2 HiGSFA network ("fused front").

The reference runs a flow node by node (``mdp.Flow.execute``, ``FaceDetectUpdated.py:699``); round 1 of this build
ran one kernel per layer, every activation through HBM.  The first three layers of the ultra-thin networks
(4x4-pixel fields -> horizontal join -> vertical join) form 8x8-pixel *subtrees* that do not interact:
``front_kernel`` (``csrc/front_tc.cuh``) walks them with one thread per window, keeps every intermediate result
in tensor memory / registers and writes only the third layer's output.

This module recognises the pattern in a compiled :class:`plan.PlanSpec` and builds what the kernel streams:

* per subtree, the weight *chunks* of its 4 + 2 + 1 nodes in the order the kernel consumes them.  The contraction runs
  as a 2-piece FP16 split on ``tcgen05.mma.kind::f16`` (K = 16 per instruction, measured 2.5x cheaper per instruction
  and 5x per unit of K than ``kind::tf32``, ``profiles/tc_probe2_r02.txt``): ``W * 2^t = Whi + Wlo`` (FP16 each,
  ``t`` per level so that ``max |W| 2^t`` sits at 2^14), ``A = Ahi + Alo`` produced by the threads,
  ``D = Ahi Whi + Ahi Wlo + Alo Whi`` accumulated in FP32 -- 22 mantissa bits per operand, the 3xTF32 quality.
* each chunk = ``[head | hi image | lo image]``: the head carries what the *consumer* of an accumulator needs (bias
  of the child nodes, means of the power terms), the images are canonical K-major no-swizzle core matrices
  ``[k/8][n/8][n%8][k%8]`` of 16-bit elements.
* tables: pixel origin of every subtree pair, pixel offsets of the four first-layer nodes, output column.

Term order of a join level (two children, ``NP`` = child width padded to 8): the kernel gives each child to one warp,
whose term sequence ``s = 0 .. 2 NP - 1`` is ``s < NP``: identity of column s, else |x - mean|^p of column ``s - NP``; the
two warps fill the two 16-term halves of every 32-term chunk, so term ``s`` of child ``c`` sits in A column
``t = 32 (s // 16) + 16 c + s % 16``.
"""
from __future__ import annotations

import struct

import numpy as np

from . import expansions as ex

HEAD_BYTES = 768
CHUNK_TERMS = 32
SUPPORTED_NP = ((8, 8), (8, 16), (16, 16), (16, 24), (16, 32))     # (NP1, NP2) instantiated in csrc/front_tc.cuh
MAX_HALF = 16384.0            # |value| bound that keeps FP16 operands far from overflow
MAGIC = b"HGSFAFR1"


class NotFusable(Exception):
    pass


def _ceil(a, b):
    return -(-a // b) * b


def _level_tables(op):
    """(p, means (n_w, d), W_id (n_w, d, N), W_pow (n_w, d, N), bias_eff (n_w, N)) of a single-pass op whose term table
    is [identity over all inputs][|x|^p over all inputs]."""
    if len(op.passes) != 1 or op.n_rows != 0 or op.mode == "copy":
        raise NotFusable("not a single-pass op")
    ps = op.passes[0]
    t = ps["terms"]
    d = op.d_in
    if len(t) != 2 * d:
        raise NotFusable("term table is not [id][pow] over the inputs")
    if not (np.all(t["op"][:d] == ex.OP_ID) and np.all(t["op"][d:] == ex.OP_ABSPOW)
            and np.array_equal(t["i"][:d], np.arange(d)) and np.array_equal(t["i"][d:], np.arange(d))):
        raise NotFusable("term table is not [id][pow] over the inputs")
    p = float(t["p"][d])
    if not np.all(t["p"][d:] == t["p"][d]) or not (0.0 < p <= 1.0):
        raise NotFusable("power terms with different exponents")
    if np.any(ps["col_off"] != 0) or np.any(ps["n_valid"] != ps["n_valid"][0]):
        raise NotFusable("nodes of the layer write different numbers of columns")
    nv = int(ps["n_valid"][0])
    W = np.asarray(ps["W"], dtype=np.float64)[:, :, :nv]
    b = np.asarray(ps["b"], dtype=np.float64)[:, :nv]
    m = np.asarray(op.in_offset, dtype=np.float64)
    W_id, W_pow = W[:, :d], W[:, d:]
    bias_eff = b - np.einsum("wk,wkn->wn", m, W_id)       # identity terms read un-centred inputs
    return p, m, W_id, W_pow, bias_eff, nv


def _children(op, prev, prev_nv):
    """For every node of ``op``: the two nodes of ``prev`` whose complete outputs it gathers, in order."""
    col_owner = {}
    for nd in range(prev.n_nodes):
        c0 = int(prev.out_col[nd])
        for j in range(prev_nv):
            col_owner[c0 + j] = (nd, j)
    if op.d_in != 2 * prev_nv:
        raise NotFusable("fan-in is not two complete child nodes")
    kids = np.zeros((op.n_nodes, 2), dtype=np.int64)
    for nd in range(op.n_nodes):
        g = [col_owner.get(int(c)) for c in op.gather[nd]]
        for c in range(2):
            part = g[c * prev_nv:(c + 1) * prev_nv]
            if any(x is None for x in part) or any(x != (part[0][0], j) for j, x in enumerate(part)):
                raise NotFusable("a node does not gather two complete child nodes")
            kids[nd, c] = part[0][0]
    if len(set(kids.reshape(-1).tolist())) != 2 * op.n_nodes:
        raise NotFusable("child nodes are shared between parents")
    return kids


def _half_split(w):
    hi = w.astype(np.float16)
    lo = (w - hi.astype(np.float64)).astype(np.float16)
    return hi, lo


def _image(Wt, N):
    """Wt: (32, n_real) float64 block of scaled weights (terms x columns) -> bytes of [hi image | lo image],
    each [k/8][n/8][n%8][k%8] halves with N columns."""
    full = np.zeros((CHUNK_TERMS, N))
    full[:, :Wt.shape[1]] = Wt
    hi, lo = _half_split(full)

    def canon(a):
        return np.ascontiguousarray(a.reshape(CHUNK_TERMS // 8, 8, N // 8, 8).transpose(0, 2, 3, 1)).tobytes()
    return canon(hi) + canon(lo)


def _head(*parts):
    h = np.concatenate([np.asarray(p, dtype=np.float32).reshape(-1) for p in parts]) if parts else np.zeros(0, np.float32)
    raw = h.tobytes()
    if len(raw) > HEAD_BYTES:
        raise NotFusable("chunk head of %d bytes" % len(raw))
    return raw + b"\0" * (HEAD_BYTES - len(raw))


def _pad(v, n):
    out = np.zeros(n)
    out[:len(v)] = v
    return out


def join_column(NP, child, s):
    """A column of term ``s`` (position in the child's sequence [identity NP | power NP]) of child ``child``."""
    return 32 * (s // 16) + 16 * child + s % 16


class FrontSpec(object):
    """Everything ``front_kernel`` needs (``serialize`` writes the blob section parsed by csrc/flow.cu)."""

    n_levels = 3

    def __init__(self):
        self.n_sub = 0
        self.img_w = self.img_h = 0
        self.np1 = self.np2 = 0
        self.nn = (16, 16, 16)          # MMA N per level
        self.nv = (0, 0, 0)             # valid output columns per level
        self.nch = (1, 1, 1)            # chunks per level
        self.scale = (1.0, 1.0, 1.0)    # 2^-t per level (accumulator -> value)
        self.clip = ((0.0, 0.0),) * 3
        self.pexp = (0.8, 0.8, 0.8)
        self.out_dim = 0
        self.pair_xy = None             # (n_sub/2, 2) int32
        self.l0_off = None              # (n_sub, 4) int32: dy | dx << 8 inside the 16 x 8 pair box
        self.out_col = None             # (n_sub,) int32
        self.nodes = None               # (n_sub, 7) node ids: L0 a b c d, L1 ab cd, L2   (tests / diagnostics)
        self.sub_bytes = 0
        self.wimg = b""

    def chunk_bytes(self, level):
        return HEAD_BYTES + self.nn[level] * 128


def build(spec):
    """FrontSpec for ops 0..2 of ``spec`` or raises NotFusable with the reason."""
    if len(spec.ops) < 4:
        raise NotFusable("fewer than four layer operations")
    op0, op1, op2 = spec.ops[:3]
    if op0.d_in != 16:
        raise NotFusable("first-layer receptive fields are not 4 x 4 pixels")
    lv = [_level_tables(op) for op in (op0, op1, op2)]
    nv = tuple(l[5] for l in lv)
    np1, np2 = _ceil(nv[0], 8), _ceil(nv[1], 8)
    if (np1, np2) not in SUPPORTED_NP:
        raise NotFusable("child widths (%d, %d) -> (%d, %d) not instantiated" % (nv[0], nv[1], np1, np2))
    nn = (_ceil(nv[0], 16), _ceil(nv[1], 16), _ceil(nv[2], 16))
    if nn[0] != 16 or nn[1] > 32 or nn[2] > 32:
        raise NotFusable("layer widths %r exceed the accumulator slots" % (nv,))
    for op in (op0, op1):
        lo, hi = op.clip
        if not (np.isfinite(lo) and np.isfinite(hi) and lo <= 0.0 <= hi and max(-lo, hi) <= MAX_HALF):
            raise NotFusable("inputs of a fused level are not bounded by a clip node")
    # ---- geometry of the first layer: 4 x 4 raster blocks of a W-pixel-wide image
    g0 = op0.gather.astype(np.int64)
    W = int(g0[0, 4] - g0[0, 0])
    if W < 16 or W % 16 or spec.input_dim % W:
        raise NotFusable("first layer is not a raster of a 16-aligned image width")
    H = spec.input_dim // W
    org = g0[:, 0]
    blk = (np.arange(4)[:, None] * W + np.arange(4)[None, :]).reshape(-1)
    if not np.array_equal(g0, org[:, None] + blk[None, :]):
        raise NotFusable("first-layer fields are not 4 x 4 raster blocks")
    y0, x0 = org // W, org % W
    k1 = _children(op1, op0, nv[0])
    k2 = _children(op2, op1, nv[1])
    n_sub = op2.n_nodes
    if n_sub % 2:
        raise NotFusable("odd number of subtrees")
    subs = []
    for n2 in range(n_sub):
        u, v = k2[n2]
        a, b = k1[u]
        c, d = k1[v]
        xs, ys = x0[[a, b, c, d]], y0[[a, b, c, d]]
        bx, by = int(xs.min()), int(ys.min())
        if xs.max() - bx > 4 or ys.max() - by > 4 or bx % 8 or by % 8:
            raise NotFusable("a subtree does not cover an aligned 8 x 8 pixel block")
        subs.append(dict(n2=n2, l1=(u, v), l0=(a, b, c, d), bx=bx, by=by))
    subs.sort(key=lambda s: (s["by"], s["bx"]))
    f = FrontSpec()
    f.n_sub, f.img_w, f.img_h, f.np1, f.np2, f.nn, f.nv = n_sub, W, H, np1, np2, nn, nv
    f.pexp = tuple(l[0] for l in lv)
    f.clip = tuple((float(op.clip[0]), float(op.clip[1])) for op in (op0, op1, op2))
    f.out_dim = op2.out_dim
    f.pair_xy = np.zeros((n_sub // 2, 2), dtype=np.int32)
    f.l0_off = np.zeros((n_sub, 4), dtype=np.int32)
    f.out_col = np.zeros(n_sub, dtype=np.int32)
    f.nodes = np.zeros((n_sub, 7), dtype=np.int32)
    for p in range(n_sub // 2):
        s0, s1 = subs[2 * p], subs[2 * p + 1]
        if s0["by"] != s1["by"] or s0["bx"] % 16 or s1["bx"] != s0["bx"] + 8:
            raise NotFusable("subtrees do not pair into aligned 16 x 8 pixel boxes")
        f.pair_xy[p] = (s0["bx"], s0["by"])
        for q, s in enumerate((s0, s1)):
            k = 2 * p + q
            for i, nd in enumerate(s["l0"]):
                f.l0_off[k, i] = int(y0[nd] - s0["by"]) | (int(x0[nd] - s0["bx"]) << 8)
            f.out_col[k] = int(op2.out_col[s["n2"]])
            f.nodes[k] = list(s["l0"]) + list(s["l1"]) + [s["n2"]]
    # ---- scales: max |W| 2^t in [2^13, 2^14)
    tpow = []
    for l in lv:
        wmax = max(np.abs(l[2]).max(), np.abs(l[3]).max(), 1e-30)
        tpow.append(int(np.floor(np.log2(MAX_HALF / wmax))))
    f.scale = tuple(float(2.0 ** -t) for t in tpow)
    # FP16 range of the A operands: inputs are bounded (pixels, clipped activations), so are the terms
    bound_in = (255.0, max(-f.clip[0][0], f.clip[0][1]), max(-f.clip[1][0], f.clip[1][1]))
    for k, l in enumerate(lv):
        if bound_in[k] + np.abs(l[1]).max() > MAX_HALF:
            raise NotFusable("level %d operands exceed the FP16 range" % k)
    f.nch = (1, 4 * np1 // CHUNK_TERMS, 4 * np2 // CHUNK_TERMS)

    def wsel(op, idx):
        return 0 if op.shared else idx

    def join_rows(level, node, NP, nv_child):
        """(4 NP, nv) scaled weight rows of a join node in kernel term order, and its mean vector in that order."""
        p, m, W_id, W_pow, _, nvl = lv[level]
        w = wsel((op0, op1, op2)[level], node)
        rows = np.zeros((4 * NP, nvl))
        mean = np.zeros(2 * NP)
        for c in range(2):
            for j in range(nv_child):
                i = c * nv_child + j
                rows[join_column(NP, c, j)] = W_id[w, i]
                rows[join_column(NP, c, NP + j)] = W_pow[w, i]
                mean[c * NP + j] = m[w, i]
        return rows * 2.0 ** tpow[level], mean

    chunks = []
    for k in range(n_sub):
        a, b, c, d, u, v, n2 = [int(x) for x in f.nodes[k]]
        sub = []
        for nd in (a, b, c, d):
            w = wsel(op0, nd)
            rows = np.concatenate([lv[0][2][w], lv[0][3][w]]) * 2.0 ** tpow[0]          # 16 identity + 16 power rows
            sub.append(_head(lv[0][1][w]) + _image(rows, nn[0]))
        for nd, (ca, cb) in ((u, (a, b)), (v, (c, d))):
            rows, mean = join_rows(1, nd, np1, nv[0])
            for ch in range(f.nch[1]):
                head = _head(_pad(lv[0][4][wsel(op0, ca)], np1), _pad(lv[0][4][wsel(op0, cb)], np1), mean) if ch == 0 else _head()
                sub.append(head + _image(rows[ch * CHUNK_TERMS:(ch + 1) * CHUNK_TERMS], nn[1]))
        rows, mean = join_rows(2, n2, np2, nv[1])
        for ch in range(f.nch[2]):
            head = (_head(_pad(lv[1][4][wsel(op1, u)], np2), _pad(lv[1][4][wsel(op1, v)], np2), mean,
                          _pad(lv[2][4][wsel(op2, n2)], 32)) if ch == 0 else _head())
            sub.append(head + _image(rows[ch * CHUNK_TERMS:(ch + 1) * CHUNK_TERMS], nn[2]))
        blob = b"".join(sub)
        chunks.append(blob)
    f.sub_bytes = len(chunks[0])
    assert all(len(c) == f.sub_bytes for c in chunks)
    assert f.sub_bytes == 4 * f.chunk_bytes(0) + 2 * f.nch[1] * f.chunk_bytes(1) + f.nch[2] * f.chunk_bytes(2)
    f.wimg = b"".join(chunks)
    return f


def try_build(spec):
    """(FrontSpec or None, reason)."""
    try:
        return build(spec), ""
    except NotFusable as e:
        return None, str(e)


def serialize(f):
    """Blob section: magic, 16 int64, 12 float64, tables, weight chunks (every part padded to 16 bytes)."""
    def pad16(b):
        return b + b"\0" * ((-len(b)) % 16)
    hdr = struct.pack("<16q", f.n_sub, f.img_w, f.img_h, f.np1, f.np2, f.nn[0], f.nn[1], f.nn[2], f.nv[0], f.nv[1], f.nv[2],
                      f.nch[1], f.nch[2], f.out_dim, f.sub_bytes, HEAD_BYTES)
    flt = struct.pack("<12d", *f.scale, *[c[0] for c in f.clip], *[c[1] for c in f.clip], *f.pexp)
    return b"".join([pad16(MAGIC + hdr + flt), pad16(f.pair_xy.astype(np.int32).tobytes()),
                     pad16(f.l0_off.astype(np.int32).tobytes()), pad16(f.out_col.astype(np.int32).tobytes()),
                     pad16(f.wimg)])
