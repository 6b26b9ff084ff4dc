"""Flow compiler: unpickled MDP / cuicuilco node graph -> fused layer operations -> plan blob.

The reference executes a flow node by node (``mdp.Flow.execute``; call sites
``FaceDetectUpdated.py:699``, ``face_analysis.py:1064,1257``): every Switchboard materialises a gathered
copy of the whole activation matrix, every Layer loops over its nodes in Python, every expansion
allocates temporaries.  Here the graph is lowered once, on the host and in
float64, to a short list of *layer operations* the CUDA kernels in ``csrc/flow.cu`` execute:

    op    = (gather index per node, input offset per node, passes)
    pass  = (expansion term table, W[node], b[node], destination)

* Switchboards / HeadNodes / IdentityNodes never run: they compose into the gather index of the next op.
* PCANode / WhiteningNode / SFANode / LinearRegressionNode become one pass with an identity table
  (the mean is subtracted at the operand fetch, ``in_offset``).
* GeneralExpansionNode becomes the term table of the following projection.
* iGSFANode (IEVMLRecNode) becomes either two passes -- slow features ``S = E @ W1 + b1`` kept in
  shared memory, then ``R = [x0 | S] @ W2 + b2`` with the linear-regression reconstruction folded into
  ``W2 = [V ; -beta V]`` -- or a single folded pass ``Y = E' @ W + b``, whichever executes fewer flops.

All folding is done in float64; weights are rounded to float32 once, when the blob is written.
"""
from __future__ import annotations

import struct

import numpy as np

from . import expansions as ex

TILE = 128
MAX_PASSES = 4
WARPS = 8        # most warps per CTA of layer_kernel (csrc/layer.cuh); an op runs with op.warps = 4 or 8
DST_GLOBAL, DST_ROWS = 1, 2
SMEM_TARGET = 113 * 1024    # per-CTA shared memory that still lets two CTAs share an SM
SMEM_LIMIT = 227 * 1024

_SWITCHBOARDS = ("Switchboard", "Rectangular2dSwitchboard", "PInvSwitchboard", "DoubleRect2dSwitchboard",
                 "DoubleRhomb2dSwitchboard", "MeanInverseSwitchboard", "ChannelSwitchboard")
_LAYERS = ("Layer", "CloneLayer")
_LINEAR = ("PCANode", "WhiteningNode", "SFANode", "GSFANode", "LinearRegressionNode")
_IGSFA = ("iGSFANode", "IEVMLRecNode")


class UnsupportedFlow(NotImplementedError):
    pass


def _cls(node):
    return type(node).__name__


def _get(node, *names, default=None):
    for n in names:
        v = getattr(node, n, None)
        if v is not None:
            return v
    return default


def _f64(a):
    return np.asarray(a, dtype=np.float64)


def node_input_dim(node):
    d = _get(node, "_input_dim", "input_dim")
    if d is None:
        raise UnsupportedFlow("node %s carries no input_dim" % _cls(node))
    return int(d)


def flow_nodes(flow):
    nodes = getattr(flow, "flow", None)
    if nodes is None:
        nodes = list(flow)
    return list(nodes)


# --------------------------------------------------------------------------------------------------
# per-node lowering
# --------------------------------------------------------------------------------------------------
def _linear_params(node):
    """(avg or None, W (d x m), b (m,)) with  y = (x - avg) @ W + b  (avg None: y = x @ W + b)."""
    c = _cls(node)
    out_dim = _get(node, "output_dim", "_output_dim")
    if c in ("PCANode", "WhiteningNode"):
        v = _f64(node.v)
        if out_dim is not None:
            v = v[:, :int(out_dim)]
        return _f64(node.avg).reshape(-1), v, np.zeros(v.shape[1])
    if c in ("SFANode", "GSFANode"):
        sf = _f64(node.sf)
        if out_dim is not None:
            sf = sf[:, :int(out_dim)]
        bias = _get(node, "_bias")
        avg = _get(node, "avg")
        if bias is None:
            return _f64(avg).reshape(-1), sf, np.zeros(sf.shape[1])
        bias = _f64(bias).reshape(-1)[:sf.shape[1]]
        if avg is not None:
            avg = _f64(avg).reshape(-1)
            if np.allclose(avg @ sf, bias, rtol=1e-9, atol=1e-9 * (1.0 + np.abs(bias).max(initial=0.0))):
                return avg, sf, np.zeros(sf.shape[1])   # subtract the mean first: better conditioned in fp32
        return None, sf, -bias
    if c == "LinearRegressionNode":
        beta = _f64(node.beta)
        if _get(node, "with_bias", default=True):
            return None, beta[1:], beta[0].copy()
        return None, beta, np.zeros(beta.shape[1])
    raise UnsupportedFlow("no linear lowering for node class %r" % c)


class _Pass(object):
    """One projection of one node: term table over sources, W (K x N), b (N), destination."""

    def __init__(self, terms, W, b, to_global, col_off=0, to_rows=False):
        self.terms = terms          # TERM_DTYPE array, indices are *source* indices
        self.W = W
        self.b = b
        self.to_global = to_global
        self.to_rows = to_rows
        self.col_off = col_off
        self.row0 = -1              # assigned at layer assembly


class _NodeProgram(object):
    def __init__(self, d_in):
        self.d_in = d_in
        self.offset = np.zeros(d_in)
        self.passes = []
        self.out_dim = 0
        self.alg_flops = 0


def _identity_terms(sources):
    t = np.zeros(len(sources), dtype=ex.TERM_DTYPE)
    t["op"] = ex.OP_ID
    t["i"] = np.asarray(sources, dtype=np.int64)
    return t


def _remap_terms(terms, sources):
    """Terms written over a local vector -> terms over global source indices."""
    src = np.asarray(sources, dtype=np.int64)
    out = terms.copy()
    out["i"] = src[terms["i"]]
    two = (terms["op"] == ex.OP_MUL) | (terms["op"] == ex.OP_MUL3)
    out["j"][two] = src[terms["j"][two]]
    three = terms["op"] == ex.OP_MUL3
    if three.any():
        out["p"][three] = src[terms["p"][three].astype(np.int64)].astype(np.float32)
    return out


class _Lowerer(object):
    """Lowers the node sequence applied to ONE receptive field (one child of a Layer)."""

    def __init__(self, d_in, igsfa_mode="auto", j_pad=None):
        self.prog = _NodeProgram(d_in)
        self.cur = list(range(d_in))   # source index of every component of the current vector
        self.cur_is_input = True
        self.pending = None            # term table over the current vector (from an expansion node)
        self.n_rows = 0
        self.igsfa_mode = igsfa_mode
        self.j_pad = j_pad
        self.igsfa_J = []

    # -- helpers
    def _new_rows(self, n):
        """Shared-memory rows for an intermediate result of width n.  The producing pass stores its
        padded width (``_choose_tile``), so the allocation advances by that."""
        r0 = self.n_rows
        self.n_rows += _padded_cols(n)
        return [self.prog.d_in + r0 + k for k in range(n)], r0

    def _shift_current(self, mean):
        """current vector <- current vector - mean."""
        mean = _f64(mean).reshape(-1)
        if self.pending is not None:
            raise UnsupportedFlow("mean subtraction after an un-projected expansion")
        if self.cur_is_input:
            self.prog.offset[np.asarray(self.cur)] += mean
        else:
            last = self.prog.passes[-1]
            last.b = last.b - mean

    def _emit(self, terms, W, b, final, col_off=0, also_rows=False):
        p = _Pass(terms, W, b, to_global=final, col_off=col_off, to_rows=(not final) or also_rows)
        self.prog.passes.append(p)
        return p

    # -- node kinds
    def linear(self, node, final, fold_mean=False):
        avg, W, b = _linear_params(node)
        d = len(self.cur) if self.pending is None else len(self.pending)
        if W.shape[0] != d:
            raise ValueError("%s: x has dimension %d, should be %d" % (_cls(node), d, W.shape[0]))
        if avg is not None:
            if self.pending is None and not fold_mean:
                self._shift_current(avg)
            else:
                b = b - avg @ W    # mean of the expanded vector: fold
        terms = _identity_terms(self.cur) if self.pending is None else self.pending
        self.pending = None
        self.prog.alg_flops += 2 * W.shape[0] * W.shape[1] + (W.shape[0] if avg is not None else W.shape[1])
        self._emit(terms, W, b, final)
        if not final:
            self.cur, r0 = self._new_rows(W.shape[1])
            self.prog.passes[-1].row0 = r0
            self.cur_is_input = False
        else:
            self.prog.out_dim = W.shape[1]

    def expansion(self, node):
        if self.pending is not None:
            raise UnsupportedFlow("two consecutive expansions without a projection")
        funcs = list(node.funcs)
        local = ex.lower(funcs, len(self.cur))
        self.prog.alg_flops += ex.term_flops(local)
        self.pending = _remap_terms(local, self.cur)

    def igsfa(self, node, final):
        if not final:
            raise UnsupportedFlow("an iGSFA node must be the last node of its receptive field")
        if self.pending is not None:
            raise UnsupportedFlow("iGSFA node after an un-projected expansion")
        d = len(self.cur)
        self._shift_current(node.x_mean)
        x_src = list(self.cur)
        p_src = x_src
        pre = _get(node, "pre_expansion_node")
        self.prog.alg_flops += d
        if pre is not None:
            if _cls(pre) not in _LINEAR:
                raise UnsupportedFlow("iGSFA pre_expansion_node of class %r" % _cls(pre))
            # the pre-expansion mean goes into that projection's bias: the shared input offset must stay x_mean,
            # because the PCA / residual branch below reads the same x0 = x - x_mean rows
            self.linear(pre, final=False, fold_mean=True)
            p_src = list(self.cur)
        expn = _get(node, "exp_node")
        if expn is not None:
            local = ex.lower(list(expn.funcs), len(p_src))
            self.prog.alg_flops += ex.term_flops(local)
            T = _remap_terms(local, p_src)
        else:
            T = _identity_terms(p_src)
        D = len(T)
        J = int(node.num_sfa_features_preserved)
        avg_s, sf, b_s = _linear_params(node.sfa_node)
        if sf.shape[0] != D:
            raise ValueError("iGSFA sfa_node: expanded dimension %d, should be %d" % (D, sf.shape[0]))
        if avg_s is not None:
            b_s = b_s - avg_s @ sf
        self.prog.alg_flops += 2 * D * sf.shape[1] + sf.shape[1]
        magn = _f64(_get(node, "magn_n_sfa_x", default=1.0)).reshape(-1)
        magn = np.broadcast_to(magn, (sf.shape[1],)) if magn.size == 1 else magn[:sf.shape[1]]
        W1 = sf[:, :J] * magn[:J]
        b1 = b_s[:J] * magn[:J]
        self.prog.alg_flops += J
        pca = _get(node, "pca_node")
        if pca is None:
            self._emit(T, W1, b1, final=True)
            self.prog.out_dim = J
            self.igsfa_J.append(J)
            return
        avg_p, V, b_p = _linear_params(pca)
        if V.shape[0] != d:
            raise ValueError("iGSFA pca_node: dimension %d, should be %d" % (d, V.shape[0]))
        if avg_p is None:
            avg_p = np.zeros(d)
        P = V.shape[1]
        recon = bool(_get(node, "reconstruct_with_sfa", default=True)) and J > 0
        if recon:
            _, B, b0 = _linear_params(node.lr_node)      # x_app = s_n @ B + b0
            if B.shape != (J, d):
                raise ValueError("iGSFA lr_node: beta maps %s, expected (%d, %d)" % (B.shape, J, d))
            if IGSFA_LR_INPUT == "unscaled":
                # lr_node fed with the raw sfa_node output: x_app = (s_n / magn) @ B + b0 -- one row scaling of B
                B = B / magn[:J, None]
            elif IGSFA_LR_INPUT != "scaled":
                raise ValueError("HGSFA_IGSFA_LR_INPUT must be 'scaled' or 'unscaled', not %r" % (IGSFA_LR_INPUT,))
            self.prog.alg_flops += 2 * J * d + d
        else:
            B, b0 = np.zeros((J, d)), np.zeros(d)
        self.prog.alg_flops += 2 * d + 2 * d * P
        M2 = -(B @ V)                                    # (J x P): contribution of s_n to the residual part
        c2 = b_p - (b0 + avg_p) @ V
        self.prog.out_dim = J + P
        self.igsfa_J.append(J)

        jp = self.j_pad if self.j_pad is not None else J

        def padded(k, n):   # what the kernel executes: N rounded up to the register tiling
            return 2 * k * _padded_cols(n)

        cost_two = padded(D, jp) + padded(d + (jp if recon else 0), P)
        # identity rows the folded form can reuse (only when the expansion reads x0 directly)
        id_row = {}
        if p_src is x_src or p_src == x_src:
            for k in range(D):
                if T["op"][k] == ex.OP_ID and int(T["i"][k]) not in id_row:
                    id_row[int(T["i"][k])] = k
        missing = [s for s in x_src if s not in id_row]
        cost_fold = padded(D + len(missing), J + P) if J + P <= 256 else float("inf")
        mode = self.igsfa_mode
        if mode == "auto":
            # the single folded pass has no shared-row round trip, no second epilogue and no K-split
            # reduction; measured on B200 it wins until it executes ~1.7x the flops of the two-pass form
            mode = "fold" if cost_fold <= FOLD_BIAS * cost_two else "two_pass"
        if mode == "fold":
            Tf = np.concatenate([T, _identity_terms(missing)]) if missing else T
            Wf = np.zeros((len(Tf), J + P))
            Wf[:D, :J] = W1
            Wf[:D, J:] = W1 @ M2
            for s_i, s in enumerate(x_src):
                row = id_row[s] if s in id_row else D + missing.index(s)
                Wf[row, J:] += V[s_i]
            bf = np.concatenate([b1, b1 @ M2 + c2])
            self._emit(Tf, Wf, bf, final=True)
        else:
            pa = self._emit(T, W1, b1, final=True, col_off=0, also_rows=recon)
            pa.n_keep = J
            if recon:
                s_rows, r0 = self._new_rows(jp)
                pa.row0 = r0
                pa.pad_to = jp
                T2 = _identity_terms(x_src + s_rows)
                W2 = np.zeros((d + jp, P))
                W2[:d] = V
                W2[d:d + J] = M2
            else:
                T2 = _identity_terms(x_src)
                W2 = V
            self._emit(T2, W2, c2, final=True, col_off=J)
        self.mode_used = mode

    def run(self, nodes):
        for k, node in enumerate(nodes):
            final = k == len(nodes) - 1
            c = _cls(node)
            if c == "FlowNode":
                raise UnsupportedFlow("nested FlowNode inside a FlowNode")
            if c in _LINEAR:
                self.linear(node, final)
            elif c == "GeneralExpansionNode":
                if final:
                    raise UnsupportedFlow("an expansion must be followed by a projection inside its layer")
                self.expansion(node)
            elif c in _IGSFA:
                self.igsfa(node, final)
            elif c == "IdentityNode":
                if final:
                    raise UnsupportedFlow("trailing IdentityNode in a receptive field")
            else:
                raise UnsupportedFlow("no lowering for node class %r (pyfaceanalysis_b200/plan.py)" % c)
        return self.prog


def _child_sequence(child):
    if _cls(child) == "FlowNode":
        return flow_nodes(_get(child, "_flow", "flow"))
    return [child]


# --------------------------------------------------------------------------------------------------
# layer assembly
# --------------------------------------------------------------------------------------------------
class OpSpec(object):
    """One fused layer operation (host description; ``tests/plan_interp.py`` can execute it in numpy)."""

    def __init__(self):
        self.n_nodes = 0
        self.d_in = 0
        self.in_dim = 0
        self.out_dim = 0
        self.shared = False
        self.gather = None      # (n_nodes, d_in) int32
        self.in_offset = None   # (n_w, d_in) float64
        self.out_col = None     # (n_nodes,) int32
        self.passes = []        # list of dict(terms, W (n_w,K,Npad), b (n_w,Npad), dst, row0, cfg, n_valid, col_off, K_real, N_real)
        self.n_rows = 0
        self.twc = 1
        self.npc = 1
        self.warps = 8
        self.n_runs = 1
        self.runs = None
        self.param_floats = 0
        self.alg_flops = 0
        self.exe_flops = 0
        self.mode = ""
        self.engine = "ffma"    # "tc": tensor-core kernel (csrc/layer_tc.cuh)
        self.tc = None          # dict(nd, nstx, nw, na, Npad16, Kpad, n_chunks, tmem_cols, smem)
        self.clip = (-np.inf, np.inf)   # saturation applied when results are stored to the output buffer


NT_CHOICES = tuple(int(v) for v in __import__("os").environ.get("HGSFA_NT_CHOICES", "8,12,16,20,24,28,32").split(","))
WIDE_TILES = __import__("os").environ.get("HGSFA_WIDE_TILES", "0") != "0"   # measured: no gain on U11L_64 (profiles/README_r01.md)


def _choose_tile(n_real):
    """(NT, NTL): columns per warp register tile and number of column tiles, NT * NTL >= n_real.

    Minimises padded columns; on ties prefers fewer, wider column tiles: every column tile of a node is a
    different warp that re-evaluates the expansion (MUFU work scales with NTL), and a wider tile amortises the
    operand conversion over more FMAs."""
    best = None
    for ntl in (1, 2, 4):
        for nt in NT_CHOICES:
            if nt * ntl < n_real:
                continue
            key = (nt * ntl, ntl) if WIDE_TILES else (nt * ntl, nt > 16, -nt)
            if best is None or key < best[0]:
                best = (key, nt, ntl)
    if best is None:
        raise UnsupportedFlow("a projection with %d output columns per node (max 256)" % n_real)
    return best[1], best[2]


def _padded_cols(n_real):
    nt, ntl = _choose_tile(n_real)
    return nt * ntl


class _NotTensorCore(Exception):
    pass


def _assemble_layer(children, gather_cols, in_dim, igsfa_mode="auto"):
    """Tensor-core engine first (one folded contraction per node: tcgen05 does not care about the extra
    columns) when the op qualifies, else the FFMA engine with its own fold / two-pass choice."""
    if ENGINE in ("tc", "auto") and igsfa_mode in ("auto", "fold"):
        try:
            return _assemble_layer_impl(children, gather_cols, in_dim, "fold", "tc")
        except (_NotTensorCore, UnsupportedFlow):
            pass
    return _assemble_layer_impl(children, gather_cols, in_dim, igsfa_mode, "ffma")


def _assemble_layer_impl(children, gather_cols, in_dim, igsfa_mode, engine):
    """children: list of node sequences (one per receptive field); gather_cols: list of index arrays."""
    n_nodes = len(children)
    d_in = len(gather_cols[0])
    if any(len(g) != d_in for g in gather_cols):
        raise UnsupportedFlow("a Layer whose nodes have different input dimensions")
    shared = all(all(a is b for a, b in zip(children[0], ch)) and len(ch) == len(children[0]) for ch in children)
    uniq = [children[0]] if shared else children

    # iGSFA nodes of one layer keep different numbers of slow features: pad to the layer maximum
    j_pad = None
    js = [int(n.num_sfa_features_preserved) for ch in uniq for n in ch if _cls(n) in _IGSFA]
    if js:
        j_pad = max(js)

    def lower_all(mode):
        progs = []
        for ch in uniq:
            lw = _Lowerer(d_in, igsfa_mode=mode, j_pad=j_pad)
            progs.append((lw.run(ch), lw))
        return progs

    progs = lower_all(igsfa_mode)
    if igsfa_mode == "auto":
        modes = set(getattr(lw, "mode_used", None) for _, lw in progs) - {None}
        if len(modes) > 1:   # nodes disagree: take the majority for the whole layer
            votes = [getattr(lw, "mode_used", None) for _, lw in progs]
            pick = max(modes, key=votes.count)
            progs = lower_all(pick)
    n_pass = len(progs[0][0].passes)
    if any(len(p.passes) != n_pass for p, _ in progs):
        raise UnsupportedFlow("nodes of one Layer lower to different pass counts")
    if n_pass > MAX_PASSES:
        raise UnsupportedFlow("a receptive field needs %d passes (max %d)" % (n_pass, MAX_PASSES))

    op = OpSpec()
    op.n_nodes, op.d_in, op.in_dim, op.shared = n_nodes, d_in, in_dim, shared
    op.gather = np.asarray(gather_cols, dtype=np.int32).reshape(n_nodes, d_in)
    op.in_offset = np.stack([p.offset for p, _ in progs])
    out_dims = [p.out_dim for p, _ in progs]
    if shared:
        out_dims = out_dims * n_nodes
    op.out_col = np.concatenate([[0], np.cumsum(out_dims)[:-1]]).astype(np.int32)
    op.out_dim = int(np.sum(out_dims))
    op.mode = getattr(progs[0][1], "mode_used", "")
    op.alg_flops = int(sum(p.alg_flops for p, _ in progs) * (n_nodes if shared else 1))

    row_cursor = 0
    for k in range(n_pass):
        plist = [p.passes[k] for p, _ in progs]
        t0 = plist[0].terms
        for q in plist[1:]:
            if len(q.terms) != len(t0) or not np.array_equal(q.terms, t0):
                raise UnsupportedFlow("nodes of one Layer use different expansion tables")
        K = len(t0)
        n_real = max(q.W.shape[1] for q in plist)
        to_rows = plist[0].to_rows
        pad_to = max([getattr(q, "pad_to", 0) for q in plist] + [n_real]) if to_rows else n_real
        nt, ntl = _choose_tile(max(n_real, pad_to))
        npad = nt * ntl
        n_w = len(plist)
        W = np.zeros((n_w, K, npad))
        b = np.zeros((n_w, npad))
        n_valid = np.zeros(n_nodes, dtype=np.int32)
        col_off = np.zeros(n_nodes, dtype=np.int32)
        for w_i, q in enumerate(plist):
            W[w_i, :, :q.W.shape[1]] = q.W
            b[w_i, :q.W.shape[1]] = q.b
        for nd in range(n_nodes):
            q = plist[0 if shared else nd]
            n_valid[nd] = getattr(q, "n_keep", q.W.shape[1]) if q.to_global else 0
            col_off[nd] = q.col_off
        dst = (DST_GLOBAL if plist[0].to_global else 0) | (DST_ROWS if to_rows else 0)
        row0 = 0
        if to_rows:
            # rows are allocated in emission order by the lowerer; both sides must agree
            row0 = row_cursor
            if any(q.row0 != row0 for q in plist):
                raise UnsupportedFlow("nodes of one Layer produce intermediate results of different widths")
            row_cursor += npad
        op.passes.append(dict(terms=t0.copy(), W=W, b=b, dst=dst, row0=row0, NT=nt, NTL=ntl, n_valid=n_valid,
                              col_off=col_off, K_real=K, N_real=n_real, K=K, Npad=npad))
        op.exe_flops += n_nodes * 2 * K * npad
    op.n_rows = row_cursor
    if engine == "tc":
        if not _tc_eligible(op) or (ENGINE == "auto" and op.passes[0]["K"] < TC_MIN_K):
            raise _NotTensorCore()
        _decompose_tc(op)
    else:
        _decompose(op)
    return op


def _gather_runs(gather_row):
    """Contiguous feature runs of one receptive field: list of (i0, f0, len)."""
    g = np.asarray(gather_row, dtype=np.int64)
    runs = []
    i0 = 0
    for i in range(1, len(g) + 1):
        if i == len(g) or g[i] != g[i - 1] + 1:
            runs.append((i0, int(g[i0]), i - i0))
            i0 = i
    return runs


def _op_smem(op, twc, el=4, stages=1, warps=None):
    """Shared memory of layer_kernel for this op (mirrors layout_op in csrc/flow.cu)."""
    warps = warps or op.warps

    def up(x):
        return (x + 127) // 128 * 128
    n_terms = sum(ps["K"] for ps in op.passes)
    scratch = 0
    for ps in op.passes:
        sw, tw, ks = _pass_split(ps, twc, warps)
        if ks > 1:
            scratch = max(scratch, (warps // 2) * sw * ps["NT"] * TILE * 4)
    raw = op.d_in * TILE * el
    stage = up(twc * raw + op.param_floats * 4)
    return 128 + up(n_terms * 8) + up(twc * op.n_rows * TILE * 4) + up(scratch) + stages * stage


def _pass_split(ps, twc, warps):
    """(SW, TW, KS): tile slots per warp, slot groups processed concurrently, K-split, with
    TW * NTL * KS == warps.  Two slots per warp (a 8 windows x NT register tile) halve the weight reads per
    FMA; they are used whenever the accumulators still fit (NT <= 16)."""
    sw = 2 if (ps["NT"] <= 16 and twc >= 2 and TWO_SLOTS) else 1
    tw = max(1, min(twc // sw, warps // ps["NTL"]))
    return sw, tw, warps // (ps["NTL"] * tw)


import os as _os

# tuning switches (environment overrides are for experiments; defaults are what the measurements favour)
TWO_SLOTS = _os.environ.get("HGSFA_TWO_SLOTS", "1") != "0"
FOLD_MEANS = _os.environ.get("HGSFA_FOLD_MEANS", "1") != "0"
SMALL_CTAS = _os.environ.get("HGSFA_SMALL_CTAS", "1") != "0"


def _decompose(op):
    """Choose warps per CTA, tile slots per CTA and the per-pass warp decomposition.

    Measured on B200 (profiles/README_r01.md): two 4-warp CTAs per SM beat one 8-warp CTA -- the per-node
    barrier, the mbarrier wait and the epilogue of one CTA overlap the FMA loops of the other -- so an op
    uses 4-warp CTAs whenever two of them fit in shared memory, and an 8-warp CTA otherwise."""
    d_pad = -(-op.d_in // 4) * 4
    off = d_pad
    for ps in op.passes:
        ps["b_off"] = off
        off += ps["Npad"]
        ps["w_off"] = off
        off += ps["K"] * ps["Npad"]
    op.param_floats = -(-off // 4) * 4
    wide = any(ps["NT"] > 16 or (ps["NT"] > 8 and TWO_SLOTS) for ps in op.passes)

    def best_twc(warps, budget):
        # enough slots that no pass needs a K-split, then as many as shared memory allows
        want = 1
        for ps in op.passes:
            sw = 2 if (ps["NT"] <= 16 and TWO_SLOTS) else 1
            want = max(want, min(16, sw * max(1, warps // ps["NTL"])))
        cands = [t for t in (16, 8, 4, 2, 1) if t <= want]
        fits = [t for t in cands if _op_smem(op, t, warps=warps) <= budget]
        return fits[0] if fits else None

    choice = None
    if SMALL_CTAS and all(ps["NTL"] <= 4 for ps in op.passes):
        # registers: a wide kernel (255 regs) allows 2 x 128 threads per SM, a narrow one 4 x 128
        per_sm = 2 if wide else 4
        t = best_twc(4, SMEM_LIMIT // per_sm - 1024)
        if t is None and not wide:
            t = best_twc(4, SMEM_LIMIT // 2 - 1024)
        if t is not None:
            choice = (4, t)
    if choice is None:
        t = best_twc(8, SMEM_LIMIT if wide else SMEM_TARGET) or best_twc(8, SMEM_LIMIT)
        if t is not None:
            choice = (8, t)
        elif all(ps["NTL"] <= 4 for ps in op.passes) and best_twc(4, SMEM_LIMIT) is not None:
            choice = (4, best_twc(4, SMEM_LIMIT))       # one 4-warp CTA per SM: smaller K-split scratch
        else:
            raise UnsupportedFlow("a receptive field of %d inputs does not fit in shared memory" % op.d_in)
    op.warps, op.twc = choice
    for ps in op.passes:
        ps["SW"], ps["TW"], ps["KS"] = _pass_split(ps, op.twc, op.warps)
    op.npc = max(1, min(8, op.n_nodes // 32))
    runs = [_gather_runs(g) for g in op.gather]
    op.n_runs = max(len(r) for r in runs)
    op.runs = np.zeros((op.n_nodes, op.n_runs, 4), dtype=np.int32)
    for nd, rl in enumerate(runs):
        for r, (i0, f0, ln) in enumerate(rl):
            op.runs[nd, r] = (i0, f0, ln, 0)


# "auto": single-pass ops with at least TC_MIN_K expansion terms per node run on the tensor cores (tcgen05, 3xTF32),
# the rest (gather-only ops, anything the tensor-core kernel does not take) on the FFMA kernel; "ffma" forces the
# FP32 kernel everywhere (7.5e-6 x std instead of 1.7e-4 x std against the float64 oracle, 1.5x slower).
# Measured on U11L_64 (profiles/README_r01.md): all ops on tcgen05 39.4 ms, first two layers on FFMA 41.1 ms.
ENGINE = _os.environ.get("HGSFA_ENGINE", "auto")
TC_MIN_K = int(_os.environ.get("HGSFA_TC_MIN_K", "0"))
TC_CK = int(_os.environ.get("HGSFA_TC_CK", "32"))          # terms per chunk (csrc/layer_tc.cuh, compile-time constant there)
# two CTAs per SM (256 tensor-memory columns, 113 KB each): measured 56.8 vs 87.8 ms per 1M windows against one big CTA
TC_L2_BPNS = float(_os.environ.get("HGSFA_TC_L2_BPNS", "30"))   # weight-chunk streaming rate assumed by the cost model
TC_NA_CHOICES = tuple(int(v) for v in _os.environ.get("HGSFA_TC_NA_CHOICES", "2,4").split(","))
TC_TIERS = [(int(a), int(b) * 1024) for a, b in
            (t.split(":") for t in _os.environ.get("HGSFA_TC_TIERS", "256:113,512:227").split(","))]   # (columns, KB) per CTA


def _tc_eligible(op):
    if op.mode == "copy" or len(op.passes) != 1 or op.n_rows != 0:
        return False
    ps = op.passes[0]
    if ps["dst"] != DST_GLOBAL or int(ps["n_valid"].max()) > 128:
        return False
    return all(sg[4] == 0 for sg in _segments(ps["terms"], op.d_in))


def _tc_segments(segs, K):
    """Segments as the tensor-core kernel walks them (mirrors hgsfa_plan_create): identity/power fusions
    undone, every segment padded to a multiple of 8 A columns, split at chunk boundaries.
    Returns (pieces, Kpad)."""
    out = []
    cursor = 0

    def push(o, n):
        nonlocal cursor
        pos, end = cursor, cursor + -(-n // 8) * 8
        while pos < end:
            nxt = min(end, (pos // TC_CK + 1) * TC_CK)
            out.append((o, pos, nxt))
            pos = nxt
        cursor = end
    for sg in segs:
        o, k0, k1 = sg[0], sg[1], sg[2]
        if o == OP_ID_POW:
            half = (k1 - k0) // 2
            push(ex.OP_ID, half)
            push(ex.OP_ABSPOW, half)
        else:
            push(o, k1 - k0)
    return out, cursor


def _decompose_tc(op):
    """Tensor-core configuration of a single-pass op: tiles per CTA, accumulator sets, receptive-field stages,
    weight-ring stages and A stages under the 512-column tensor-memory and 227 KB shared-memory budgets.
    The score is a rough time model (MMA issue rate measured by tools/tc_probe.cu, L2 traffic of the
    streamed weight chunks, exposed loads when a resource is single-buffered)."""
    ps = op.passes[0]
    K = ps["K"]
    n_max = max(1, int(ps["n_valid"].max()))
    npad16 = -(-n_max // 16) * 16
    segs, kpad = _tc_segments(_fuse_id_pow(_segments(ps["terms"], op.d_in)), K)
    n_chunks = -(-kpad // TC_CK)
    d_pad = -(-op.d_in // 4) * 4
    head = -(-(d_pad + npad16 + 2 * K) // 4) * 4
    wstage = 2 * TC_CK * npad16 * 4

    def up(x):
        return (x + 127) // 128 * 128
    fixed = 512 + 2 * up(K * 8) + up(len(segs) * 32) + up((n_chunks + 1) * 4) + up(8 * npad16 * 4)
    t_mma = 3.0 * (kpad / 8.0) * (17.0 + 0.2 * npad16)          # ns per (node, tile)
    best = None
    forced = {k: int(_os.environ[e]) for k, e in (("twc", "HGSFA_TC_TWC"), ("nd", "HGSFA_TC_ND"), ("nstx", "HGSFA_TC_NSTX"),
                                                  ("na", "HGSFA_TC_NA"), ("nw", "HGSFA_TC_NW")) if e in _os.environ}
    # (a 128-column / one-A-stage / three-CTA tier for single-chunk ops was worth 9.1 -> 8.3 ms on layer 0 before the
    # epilogue moved to its own warps; with them two CTAs of 320 threads do better: 7.8 ms)
    tiers = [(c, m, TC_NA_CHOICES) for c, m in TC_TIERS]
    for max_cols, max_smem, na_choices in tiers:
      if best is not None:
        break
      for twc in range(1, 9):
        for nd in (1, 2):
            for na in na_choices:
                cols = nd * twc * npad16 + na * 2 * TC_CK
                if cols > 512:
                    continue
                for nstx in (1, 2):
                    for nw in (2, 3, 4):
                        cfg = dict(twc=twc, nd=nd, nstx=nstx, na=na, nw=nw)
                        if any(cfg[k] != v for k, v in forced.items()):
                            continue
                        smem = fixed + nstx * up(twc * op.d_in * TILE * 4 + head * 4) + nw * up(wstage)
                        if smem > SMEM_LIMIT or cols > max_cols or smem > max_smem:
                            continue
                        t_w = n_chunks * wstage / twc / TC_L2_BPNS                    # L2 -> SM bytes per ns and SM
                        t = max(t_mma, t_w)
                        if nstx == 1:
                            t += (1500.0 + op.d_in * TILE * 4 * twc / 100.0) / twc    # exposed receptive-field load per node
                        if nd == 1:
                            t += 500.0 / twc                                          # MMA pipe drained before the epilogue
                        t += 300.0 / twc                                              # per-node hand-overs
                        t += (0.0 if na == 4 else (0.05 if na == 2 else 0.5) * t_mma) + (0.0 if nw >= 3 else 0.03 * t_mma)
                        key = (t, smem)
                        if best is None or key < best[0]:
                            best = (key, cfg, cols, smem)
    if best is None:
        raise UnsupportedFlow("a receptive field of %d inputs does not fit in shared memory" % op.d_in)
    _, cfg, cols, smem = best
    tmem = 32
    while tmem < cols:
        tmem *= 2
    op.engine = "tc"
    op.warps, op.twc = 4, cfg["twc"]
    op.tc = dict(cfg, Npad16=npad16, Kpad=kpad, n_chunks=n_chunks, tmem_cols=tmem, smem=smem, n_segs=len(segs))
    ps["SW"], ps["TW"], ps["KS"] = 1, 1, 1
    off = d_pad
    ps["b_off"] = off
    off += ps["Npad"]
    ps["w_off"] = off
    off += ps["K"] * ps["Npad"]
    op.param_floats = -(-off // 4) * 4
    op.exe_flops = op.n_nodes * 3 * 2 * kpad * npad16          # tensor-pipe flops of the 3xTF32 split
    op.npc = max(1, min(8, op.n_nodes // 32))
    runs = [_gather_runs(g) for g in op.gather]
    op.n_runs = max(len(r) for r in runs)
    op.runs = np.zeros((op.n_nodes, op.n_runs, 4), dtype=np.int32)
    for nd_i, rl in enumerate(runs):
        for r, (i0, f0, ln) in enumerate(rl):
            op.runs[nd_i, r] = (i0, f0, ln, 0)


OP_ID_POW = 7   # segment-only op code (csrc/layer.cuh): identity rows followed by |x|^p rows of the same inputs


def _fuse_id_pow(segs):
    """[identity over rows R][|x|^p over the same rows R] -> one fused segment: the kernel reads and centres
    every input once and feeds it to both weight rows."""
    out = []
    k = 0
    while k < len(segs):
        s0 = segs[k]
        if (FUSE_ID_POW and k + 1 < len(segs) and s0[0] == ex.OP_ID and s0[4] == 0 and s0[5] >= 0):
            s1 = segs[k + 1]
            if (s1[0] == ex.OP_ABSPOW and s1[4] == 0 and s1[5] == s0[5] and s1[1] == s0[2]
                    and s1[2] - s1[1] == s0[2] - s0[1]):
                out.append((OP_ID_POW, s0[1], s1[2], s1[3], 0, s0[5]))
                k += 2
                continue
        out.append(s0)
        k += 1
    return out


FOLD_BIAS = float(_os.environ.get("HGSFA_FOLD_BIAS", "1.5"))
# which slow features feed an iGSFA node's linear reconstruction (cuicuilco source unavailable):
# "scaled" = sfa_x[:, :J] * magn_n_sfa_x (default: lr_node is trained on the rescaled features), "unscaled" = sfa_x[:, :J].
# oracle/nodes.py has the same switch; a real SavedNetworks pickle run against recorded reference outputs settles it.
IGSFA_LR_INPUT = _os.environ.get("HGSFA_IGSFA_LR_INPUT", "scaled")
FUSE_ID_POW = _os.environ.get("HGSFA_FUSE_ID_POW", "1") != "0"


def _segments(terms, d_in):
    """Runs of consecutive terms sharing (op, exponent, operand kind): list of
    (op, k0, k1, p, kind, ibase).  kind 0: operands are receptive-field rows, 1: shared rows of an
    earlier pass, 2: mixed.  ibase >= 0 when operand i advances by one row per term (no table lookup)."""
    n = len(terms)
    ops = terms["op"]
    no_p = (ex.OP_ID, ex.OP_MUL, ex.OP_MUL3, ex.OP_ABS)

    def kind_of(k):
        rows = [int(terms["i"][k])]
        if ops[k] in (ex.OP_MUL, ex.OP_MUL3):
            rows.append(int(terms["j"][k]))
        if ops[k] == ex.OP_MUL3:
            rows.append(int(terms["p"][k]))
        lo = all(r < d_in for r in rows)
        hi = all(r >= d_in for r in rows)
        return 0 if lo else (1 if hi else 2)

    kinds = [kind_of(k) for k in range(n)]
    segs = []
    k0 = 0
    for k in range(1, n + 1):
        same = (k < n and ops[k] == ops[k0] and kinds[k] == kinds[k0]
                and (ops[k] in no_p or terms["p"][k] == terms["p"][k0]))
        if not same:
            op_code = int(ops[k0])
            p = 0.0 if op_code in no_p else float(terms["p"][k0])
            i = terms["i"][k0:k].astype(np.int64)
            affine = bool(np.array_equal(i, i[0] + np.arange(k - k0)))
            segs.append((op_code, k0, k, p, kinds[k0], int(i[0]) if affine else -1))
            k0 = k
    return segs


# --------------------------------------------------------------------------------------------------
# flow-level compiler
# --------------------------------------------------------------------------------------------------
class PlanSpec(object):
    def __init__(self, input_dim):
        self.input_dim = int(input_dim)
        self.output_dim = int(input_dim)
        self.ops = []
        self.out_cols = None   # columns of the last op's buffer that form the flow output (None = all)

    @property
    def alg_flops(self):
        return sum(op.alg_flops for op in self.ops)

    @property
    def exe_flops(self):
        return sum(op.exe_flops for op in self.ops)


def compile_flow(flow, input_dim=None, igsfa_mode="auto"):
    """Lower a flow (object with ``.flow`` or a sequence of nodes) to a :class:`PlanSpec`."""
    if igsfa_mode == "auto":
        igsfa_mode = _os.environ.get("HGSFA_IGSFA_MODE", "auto")
    nodes = flow_nodes(flow)
    if not nodes:
        raise UnsupportedFlow("empty flow")
    if input_dim is None:
        input_dim = node_input_dim(nodes[0])
    spec = PlanSpec(input_dim)
    cols = np.arange(input_dim, dtype=np.int64)   # logical column -> column of the last materialised buffer
    buf_dim = input_dim
    pending_layers = []   # per-field node sequences collected from consecutive aligned Layers

    def flush_pending():
        nonlocal cols, buf_dim, pending_layers
        if not pending_layers:
            return
        fields, gathers = pending_layers
        try:
            op = _assemble_layer(fields, gathers, buf_dim, igsfa_mode)
        except UnsupportedFlow as first:
            # a folded iGSFA pass carries a D x out weight block; when that (plus the receptive field) exceeds
            # shared memory the two-pass form (D x J and (d+J) x P blocks) may still fit
            if igsfa_mode != "auto":
                raise
            try:
                op = _assemble_layer(fields, gathers, buf_dim, "two_pass")
            except UnsupportedFlow:
                raise first
        spec.ops.append(op)
        buf_dim = op.out_dim
        cols = np.arange(buf_dim, dtype=np.int64)
        pending_layers = []

    def add_layer(children):
        """children: list of child nodes, consuming consecutive slices of the current logical vector."""
        nonlocal pending_layers
        dims = [node_input_dim(ch) for ch in children]
        seqs = [_child_sequence(ch) for ch in children]
        if not pending_layers and sum(dims) != len(cols):
            raise ValueError("%s: x has dimension %d, should be %d" % ("Layer", len(cols), sum(dims)))
        if pending_layers:
            # a Layer directly after an expansion-only Layer with the same partition: fuse per field
            fields, gathers = pending_layers
            if len(fields) != len(seqs):
                raise UnsupportedFlow("expansion Layer followed by a Layer with a different partition")
            fields = [f + s for f, s in zip(fields, seqs)]
            pending_layers = [fields, gathers]
        else:
            starts = np.concatenate([[0], np.cumsum(dims)[:-1]])
            gathers = [cols[s:s + d] for s, d in zip(starts, dims)]
            pending_layers = [seqs, gathers]
        # keep collecting only while every field still ends in an un-projected expansion
        if not all(_cls(f[-1]) == "GeneralExpansionNode" for f in pending_layers[0]):
            flush_pending()

    for node in nodes:
        c = _cls(node)
        if c in _SWITCHBOARDS or (hasattr(node, "connections") and not hasattr(node, "nodes")):
            if pending_layers:
                raise UnsupportedFlow("Switchboard directly after an expansion Layer")
            conn = np.asarray(node.connections, dtype=np.int64)
            if conn.size and (conn.min() < 0 or conn.max() >= len(cols)):
                raise ValueError("Switchboard: connection index outside the %d input columns" % len(cols))
            cols = cols[conn]
        elif c == "HeadNode":
            if pending_layers:
                raise UnsupportedFlow("HeadNode directly after an expansion Layer")
            cols = cols[:int(_get(node, "output_dim", "_output_dim"))]
        elif c == "IdentityNode":
            continue
        elif c == "PointwiseFunctionNode":
            # element-wise saturation: fused into the epilogue of the producing op
            name = ex.func_name(node.func)
            lim = ex.clip_limit(name)
            if lim is None:
                raise UnsupportedFlow("PointwiseFunctionNode with function %r (only clip_<L>)" % name)
            if pending_layers:
                raise UnsupportedFlow("a clipping node directly after an expansion Layer")
            if not spec.ops:
                # nothing materialised yet (a clip node executed on its own): saturate in a copy op
                spec.ops.append(_copy_op(cols, buf_dim))
                buf_dim = spec.ops[-1].out_dim
                cols = np.arange(buf_dim, dtype=np.int64)
            lo, hi = spec.ops[-1].clip
            spec.ops[-1].clip = (max(lo, -lim), min(hi, lim))
        elif c in _LAYERS:
            add_layer(list(node.nodes))
        elif c == "SameInputLayer":
            raise UnsupportedFlow("SameInputLayer")
        else:
            add_layer([node])
    if pending_layers:
        raise UnsupportedFlow("flow ends with an expansion that is never projected")
    spec.output_dim = len(cols)
    if not spec.ops or not np.array_equal(cols, np.arange(len(cols))):
        # the flow ends in (or consists of) a gather: materialise it with identity projections over
        # 16-column receptive fields (a Switchboard executed on its own, e.g. through a node facade)
        spec.ops.append(_copy_op(cols, buf_dim))
    return spec


def _copy_op(cols, in_dim):
    n = len(cols)
    n_nodes = -(-n // 16)
    gather = np.zeros((n_nodes, 16), dtype=np.int32)
    flat = np.concatenate([cols, np.full(n_nodes * 16 - n, cols[-1] if n else 0)])
    gather[:] = flat.reshape(n_nodes, 16)
    op = OpSpec()
    op.n_nodes, op.d_in, op.in_dim, op.out_dim, op.shared = n_nodes, 16, int(in_dim), n, True
    op.gather = gather
    op.in_offset = np.zeros((1, 16))
    op.out_col = (np.arange(n_nodes) * 16).astype(np.int32)
    n_valid = np.full(n_nodes, 16, dtype=np.int32)
    n_valid[-1] = n - 16 * (n_nodes - 1)
    op.passes = [dict(terms=_identity_terms(range(16)), W=np.eye(16)[None], b=np.zeros((1, 16)), dst=DST_GLOBAL,
                      row0=0, NT=16, NTL=1, n_valid=n_valid, col_off=np.zeros(n_nodes, dtype=np.int32), K_real=16,
                      N_real=16, K=16, Npad=16)]
    op.n_rows, op.mode = 0, "copy"
    op.exe_flops = n_nodes * 2 * 16 * 16
    _decompose(op)
    return op


# --------------------------------------------------------------------------------------------------
# blob writer (format parsed by hgsfa_plan_create, csrc/flow.cu)
# --------------------------------------------------------------------------------------------------
def _pad16(b):
    return b + b"\0" * ((-len(b)) % 16)


def _arr(a, dtype):
    return _pad16(np.ascontiguousarray(a, dtype=dtype).tobytes())


FRONT = _os.environ.get("HGSFA_FRONT", "1") != "0"


def plan_front(spec):
    """Fused-front description of the first three ops (``front.FrontSpec``) or None; the reason is kept in
    ``spec.front_reason``.  Only the tensor-core engine fuses (``HGSFA_ENGINE=ffma`` stays the exact-FP32 path)."""
    from . import front as _front
    spec.front, spec.front_reason = None, "disabled"
    if FRONT and ENGINE in ("auto", "tc") and all(op.engine == "tc" for op in spec.ops[:3]):
        spec.front, spec.front_reason = _front.try_build(spec)
    return spec.front


def serialize(spec):
    """Plan blob, version 2 (parsed by hgsfa_plan_create in csrc/flow.cu; layout in DESIGN.md section 4).
    Header: magic, input_dim, last op's out_dim, n_ops, byte offset and size of the fused-front section (0 = none)."""
    from . import front as _front
    last_dim = spec.ops[-1].out_dim
    body = _serialize_ops(spec)
    fr = getattr(spec, "front", None)
    if fr is None and not hasattr(spec, "front_reason"):
        fr = plan_front(spec)
    fsec = _front.serialize(fr) if fr is not None else b""
    hdr = b"HGSFAPL2" + struct.pack("<7q", spec.input_dim, last_dim, len(spec.ops), (64 + len(body)) if fsec else 0, len(fsec), 0, 0)
    assert len(hdr) == 64 and len(body) % 16 == 0
    return hdr + body + fsec


def _serialize_ops(spec):
    out = []
    for op in spec.ops:
        n_w = 1 if op.shared else op.n_nodes
        n_terms = sum(ps["K"] for ps in op.passes)
        out.append(struct.pack("<12q", op.n_nodes, op.d_in, op.in_dim, op.out_dim, len(op.passes),
                               int(op.shared) | (op.warps << 8) | ((1 if op.engine == "tc" else 0) << 16),
                               op.n_rows, op.twc, op.alg_flops, op.exe_flops, op.npc, op.n_runs)
                   + struct.pack("<2d", float(op.clip[0]), float(op.clip[1]))
                   + struct.pack("<2q", op.param_floats, n_terms))
        out.append(_arr(op.runs, np.int32))
        out.append(_arr(op.out_col, np.int32))
        params = np.zeros((n_w, op.param_floats), dtype=np.float32)
        params[:, :op.d_in] = op.in_offset
        t16 = np.zeros((n_terms, 4), dtype=np.int16)
        t_off = 0
        seg_lists = []
        for ps in op.passes:
            segs = _fuse_id_pow(_segments(ps["terms"], op.d_in))
            bias = ps["b"].astype(np.float64).copy()
            flagged = []
            for (o, k0, k1, p, kind, ibase) in segs:
                nomean = 0
                if o == ex.OP_ID and kind == 0 and ibase >= 0 and FOLD_MEANS:
                    # identity terms over the receptive field: x_mean goes into the bias (float64 here),
                    # the kernel then feeds the raw rows to the FMAs
                    m = op.in_offset[:, ibase:ibase + (k1 - k0)]                      # (n_w, len)
                    bias -= np.einsum("wk,wkn->wn", m, ps["W"][:, k0:k1, :])
                    nomean = 1
                flagged.append((o, k0, k1, p, kind, ibase, nomean))
            seg_lists.append(flagged)
            params[:, ps["b_off"]:ps["b_off"] + ps["Npad"]] = bias
            params[:, ps["w_off"]:ps["w_off"] + ps["K"] * ps["Npad"]] = ps["W"].reshape(n_w, -1)
            t = ps["terms"]
            t16[t_off:t_off + ps["K"], 0] = t["i"]
            two = (t["op"] == ex.OP_MUL) | (t["op"] == ex.OP_MUL3)
            t16[t_off:t_off + ps["K"], 1] = np.where(two, t["j"], 0)
            t16[t_off:t_off + ps["K"], 2] = np.where(t["op"] == ex.OP_MUL3, t["p"].astype(np.int64), 0)
            ps["term_off"] = t_off
            t_off += ps["K"]
        if not np.isfinite(params).all():
            raise ValueError("non-finite flow parameters")
        out.append(_arr(params, np.float32))
        out.append(_arr(t16, np.int16))
        for ps, segs in zip(op.passes, seg_lists):
            if op.engine == "tc":
                tile_fields = (op.tc["nd"], op.tc["nstx"], op.tc["nw"], op.tc["na"])
            else:
                tile_fields = (ps["NT"], ps["NTL"], ps["KS"], ps["TW"])
            out.append(struct.pack("<16q", ps["K"], ps["Npad"], *tile_fields, ps["dst"],
                                   ps["row0"], ps["w_off"], ps["b_off"], ps["term_off"], len(segs), ps["K_real"],
                                   ps["N_real"], ps["SW"], 0))
            sb = b"".join(struct.pack("<3if4i", o, k0, k1, p, kind, ibase, nomean, 0)
                          for (o, k0, k1, p, kind, ibase, nomean) in segs)
            out.append(_pad16(sb))
            out.append(_arr(ps["n_valid"], np.int32))
            out.append(_arr(ps["col_off"], np.int32))
    return b"".join(out)


def describe(spec):
    lines = ["plan: %d -> %d, %d ops, %.3f MFLOP/window algorithmic, %.3f executed"
             % (spec.input_dim, spec.output_dim, len(spec.ops), spec.alg_flops / 1e6, spec.exe_flops / 1e6)]
    for k, op in enumerate(spec.ops):
        if op.engine == "tc":
            p, t = op.passes[0], op.tc
            lines.append("  op%-2d nodes=%-4d d_in=%-4d out=%-5d %s %s tc [K%d->N%d pad %dx%d, %d chunks] twc=%d nd=%d nstx=%d nw=%d na=%d "
                         "tmem=%d runs=%d smem=%dK"
                         % (k, op.n_nodes, op.d_in, op.out_dim, "clone" if op.shared else "layer", op.mode, p["K_real"],
                            p["N_real"], t["Kpad"], t["Npad16"], t["n_chunks"], op.twc, t["nd"], t["nstx"], t["nw"], t["na"],
                            t["tmem_cols"], op.n_runs, t["smem"] // 1024))
            continue
        ps = ", ".join("K%d->N%d(%dx%d sw%d ks%d tw%d)" % (p["K_real"], p["N_real"], p["NT"], p["NTL"], p["SW"], p["KS"], p["TW"])
                       for p in op.passes)
        lines.append("  op%-2d nodes=%-4d d_in=%-4d out=%-5d %s %s [%s] warps=%d twc=%d runs=%d rows=%d smem=%dK"
                     % (k, op.n_nodes, op.d_in, op.out_dim, "clone" if op.shared else "layer", op.mode, ps,
                        op.warps, op.twc, op.n_runs, op.n_rows, _op_smem(op, op.twc) // 1024))
    return "\n".join(lines)
