"""Synthetic HiGSFA flows with the pickle shape of the reference's ``SavedNetworks/*.pckl``.

The eight flow pickles the reference ships were stripped from the tree, so every flow
this build can execute today is a stand-in.  The factory emits the *same object graph a real pickle
unpickles to* -- ``mdp.Flow`` with ``flow = [PInvSwitchboard, Layer([iGSFANode, ...]), ...]``, attribute
names as in SURVEY.md Appendix A.2 -- so that synthetic and (future) real flows go through one loader
and one compiler (``plan.compile_flow``).

Parameters are not random matrices: each node is fitted, layer by layer in float64, on synthetic image
patches (PCA-whitened "slow" directions, least-squares reconstruction, PCA of the residual), which
gives the conditioning a trained network has -- whitening amplifies low-variance directions, and that
is what decides whether an FP32 kernel stays inside the 1e-3 x std tolerance.

Network specs (documented so the algorithmic flops can be recomputed by hand):

``U11L_64``  -- 64x64 input, "ultra thin" fan-in-2 hierarchy, 11 layers:
    L0  4x4 pixel fields, stride 4        -> 16x16 nodes
    L1..L8 alternately join 2 nodes horizontally / vertically -> 8x16, 8x8, 4x8, 4x4, 2x4, 2x2, 1x2, 1x1
    L9, L10 single full nodes on the 1x1 map
``U11L_96``  -- 96x96 input (age-like): L0 6x6 fields stride 3 -> 31x31; then 3-wide stride-2 joins
    31->15->7->3->1 horizontally / vertically (8 layers), + 2 top nodes.
"""
from __future__ import annotations

import hashlib
import os
import pickle

import numpy as np

from . import expansions as ex
from .pickles import FuncRef, new_object

NLE = "cuicuilco.nonlinear_expansion"


# ----------------------------------------------------------------------------------------------------
# specs
# ----------------------------------------------------------------------------------------------------
def _layer(grid_in, field, stride, out_dim, J, funcs, clone=False):
    return dict(grid_in=grid_in, field=field, stride=stride, out_dim=out_dim, J=J, funcs=funcs, clone=clone)


def spec_u11l_64():
    f_low = ["identity", "unsigned_08expo"]
    f_high = ["identity", "unsigned_08expo", "s10QT"]
    L = []
    L.append(_layer((64, 64), (4, 4), (4, 4), 13, 4, f_low))          # -> 16x16
    L.append(_layer((16, 16), (2, 1), (2, 1), 20, 6, f_low))          # H -> 8x16
    L.append(_layer((8, 16), (1, 2), (1, 2), 27, 8, f_low))           # V -> 8x8
    L.append(_layer((8, 8), (2, 1), (2, 1), 35, 9, f_high))           # H -> 4x8
    L.append(_layer((4, 8), (1, 2), (1, 2), 45, 10, f_high))          # V -> 4x4
    L.append(_layer((4, 4), (2, 1), (2, 1), 55, 11, f_high))          # H -> 2x4
    L.append(_layer((2, 4), (1, 2), (1, 2), 60, 12, f_high))          # V -> 2x2
    L.append(_layer((2, 2), (2, 1), (2, 1), 60, 12, f_high))          # H -> 1x2
    L.append(_layer((1, 2), (1, 2), (1, 2), 60, 12, f_high))          # V -> 1x1
    L.append(_layer((1, 1), (1, 1), (1, 1), 60, 12, f_high))
    L.append(_layer((1, 1), (1, 1), (1, 1), 60, 12, f_high))
    return dict(name="U11L_64", input_xy=(64, 64), layers=L)


def spec_u11l_96():
    f_low = ["identity", "unsigned_08expo"]
    f_high = ["identity", "unsigned_08expo", "s10QT"]
    L = []
    L.append(_layer((96, 96), (6, 6), (3, 3), 16, 5, f_low))          # -> 31x31
    L.append(_layer((31, 31), (3, 1), (2, 1), 24, 7, f_low))          # H -> 15x31
    L.append(_layer((15, 31), (1, 3), (1, 2), 32, 8, f_low))          # V -> 15x15
    L.append(_layer((15, 15), (3, 1), (2, 1), 40, 9, f_high))         # H -> 7x15
    L.append(_layer((7, 15), (1, 3), (1, 2), 50, 10, f_high))         # V -> 7x7
    L.append(_layer((7, 7), (3, 1), (2, 1), 60, 11, f_high))          # H -> 3x7
    L.append(_layer((3, 7), (1, 3), (1, 2), 60, 12, f_high))          # V -> 3x3
    L.append(_layer((3, 3), (3, 1), (1, 1), 60, 12, f_high))          # H -> 1x3
    L.append(_layer((1, 3), (1, 3), (1, 1), 60, 12, f_high))          # V -> 1x1
    L.append(_layer((1, 1), (1, 1), (1, 1), 60, 12, f_high))
    L.append(_layer((1, 1), (1, 1), (1, 1), 60, 12, f_high))
    return dict(name="U11L_96", input_xy=(96, 96), layers=L)


def spec_tiny(input_xy=(16, 16), clone_first=True):
    """Small 4-layer network for fast CPU-side tests (same node vocabulary)."""
    f_low = ["identity", "unsigned_08expo"]
    f_high = ["identity", "signed_08expo", "s6QT"]
    L = []
    L.append(_layer(input_xy, (4, 4), (4, 4), 9, 3, f_low, clone=clone_first))
    gx, gy = input_xy[0] // 4, input_xy[1] // 4
    L.append(_layer((gx, gy), (2, 1), (2, 1), 12, 4, f_high))
    L.append(_layer((gx // 2, gy), (1, 2), (1, 2), 14, 5, f_high))
    L.append(_layer((gx // 2, gy // 2), (gx // 2, gy // 2), (1, 1), 16, 6, f_high))
    return dict(name="tiny_%dx%d" % input_xy, input_xy=input_xy, layers=L)


def spec_s5l_64():
    """Small 5-layer 64x64 network (same vocabulary as U11L_64) for cascade tests on the CPU oracle."""
    f_low = ["identity", "unsigned_08expo"]
    f_high = ["identity", "unsigned_08expo", "s8QT"]
    L = []
    L.append(_layer((64, 64), (8, 8), (8, 8), 12, 4, f_low))        # -> 8x8
    L.append(_layer((8, 8), (2, 2), (2, 2), 16, 5, f_low))          # -> 4x4
    L.append(_layer((4, 4), (2, 2), (2, 2), 20, 6, f_high))         # -> 2x2
    L.append(_layer((2, 2), (2, 2), (2, 2), 24, 8, f_high))         # -> 1x1
    L.append(_layer((1, 1), (1, 1), (1, 1), 24, 8, f_high))
    return dict(name="S5L_64", input_xy=(64, 64), layers=L)


def spec_f4l_32(widths=(7, 12, 14)):
    """Small 32x32 network with the fan-in-2 front of the ultra-thin networks (4x4 pixel fields, horizontal join,
    vertical join, [identity, |x|^0.8] expansions) and two top layers: exercises the fused front kernel with other
    node widths than U11L_64's (child widths padded to 8 / 16 / 32, one to four chunks per join)."""
    f_low = ["identity", "unsigned_08expo"]
    f_high = ["identity", "unsigned_08expo", "s10QT"]
    w0, w1, w2 = widths
    L = []
    L.append(_layer((32, 32), (4, 4), (4, 4), w0, max(2, w0 // 3), f_low))        # -> 8x8
    L.append(_layer((8, 8), (2, 1), (2, 1), w1, max(2, w1 // 3), f_low))          # H -> 4x8
    L.append(_layer((4, 8), (1, 2), (1, 2), w2, max(2, w2 // 3), f_low))          # V -> 4x4
    L.append(_layer((4, 4), (2, 2), (2, 2), 16, 6, f_high))                       # -> 2x2
    L.append(_layer((2, 2), (2, 2), (1, 1), 16, 6, f_high))                       # -> 1x1
    return dict(name="F4L_32_%d_%d_%d" % widths, input_xy=(32, 32), layers=L)


SPECS = {"U11L_64": spec_u11l_64, "U11L_96": spec_u11l_96, "tiny": spec_tiny, "S5L_64": spec_s5l_64,
         "F4L_32": spec_f4l_32, "F4L_32_wide": lambda: spec_f4l_32((13, 26, 27)), "F4L_32_mid": lambda: spec_f4l_32((12, 16, 30))}


# ----------------------------------------------------------------------------------------------------
# receptive-field connections (what mdp.hinet.Rectangular2dSwitchboard / cuicuilco lattice produce)
# ----------------------------------------------------------------------------------------------------
def rect_connections(grid_xy, field_xy, stride_xy, channel_dim):
    """Row-major (y outer, x inner) channel map; one block of ``field_x*field_y*channel_dim`` inputs per
    output node, nodes ordered row-major as well."""
    gx, gy = grid_xy
    fx, fy = field_xy
    sx, sy = stride_xy
    ox = (gx - fx) // sx + 1
    oy = (gy - fy) // sy + 1
    conn = []
    for yo in range(oy):
        for xo in range(ox):
            for yi in range(yo * sy, yo * sy + fy):
                for xi in range(xo * sx, xo * sx + fx):
                    first = (yi * gx + xi) * channel_dim
                    conn.extend(range(first, first + channel_dim))
    return np.asarray(conn, dtype=np.int64), (ox, oy)


# ----------------------------------------------------------------------------------------------------
# synthetic training patches
# ----------------------------------------------------------------------------------------------------
def synthetic_patches(n, size_xy, seed):
    """uint8 (n, h*w): band-limited noise with a random face-like blob, values 0..255."""
    rng = np.random.default_rng(seed)
    w, h = size_xy
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.fftfreq(w)[None, :]
    rad = np.sqrt(fx * fx + fy * fy)
    filt = 1.0 / (1.0 + (rad / 0.06) ** 2)
    out = np.empty((n, h * w), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    for k in range(n):
        z = np.fft.ifft2(np.fft.fft2(rng.standard_normal((h, w))) * filt).real
        z = z / (z.std() + 1e-9)
        cx, cy = rng.uniform(0.3 * w, 0.7 * w), rng.uniform(0.3 * h, 0.7 * h)
        s = rng.uniform(0.15, 0.35) * w
        blob = np.exp(-(((xx - cx) / s) ** 2 + ((yy - cy) / (1.3 * s)) ** 2))
        img = 128 + 45 * z + rng.uniform(-60, 60) * blob + 3.0 * rng.standard_normal((h, w))
        out[k] = np.clip(np.rint(img), 0, 255).astype(np.uint8).reshape(-1)
    return out


def _expand(x, funcs):
    """float64 evaluation of an expansion through the product's own term tables (used for fitting only)."""
    t = ex.lower(funcs, x.shape[1])
    cols = np.empty((x.shape[0], len(t)))
    for e in range(len(t)):
        op, i, j, p = int(t["op"][e]), int(t["i"][e]), int(t["j"][e]), float(t["p"][e])
        a = x[:, i]
        if op == ex.OP_ID:
            cols[:, e] = a
        elif op == ex.OP_MUL:
            cols[:, e] = a * x[:, j]
        elif op == ex.OP_ABSPOW:
            cols[:, e] = np.abs(a) ** p
        elif op == ex.OP_SGNPOW:
            cols[:, e] = np.sign(a) * np.abs(a) ** p
        elif op == ex.OP_MUL3:
            cols[:, e] = a * x[:, j] * x[:, int(p)]
        elif op == ex.OP_ABS:
            cols[:, e] = np.abs(a)
        elif op == ex.OP_CLIP:
            cols[:, e] = np.clip(a, -p, p)
    return cols


# ----------------------------------------------------------------------------------------------------
# node fitting
# ----------------------------------------------------------------------------------------------------
def _fit_igsfa(x, funcs, out_dim, J, rng):
    """Fit one iGSFA-shaped node on data ``x`` (n, d); returns (node object, its output on x)."""
    n, d = x.shape
    x_mean = x.mean(axis=0)
    x0 = x - x_mean
    E = _expand(x0, funcs)
    D = E.shape[1]
    avg = E.mean(axis=0)
    Ec = E - avg
    # PCA-whitening of the *standardised* expanded data: quadratic terms are orders of magnitude larger
    # than linear ones, and whitening raw covariances would make every feature a difference of huge
    # numbers (ill-conditioned in any precision, and unlike a trained network whose SFA step sees
    # variance-normalised data)
    scale = Ec.std(axis=0) + 1e-12
    Ez = Ec / scale
    lam, U = np.linalg.eigh(Ez.T @ Ez / (n - 1))
    lam, U = lam[::-1], U[:, ::-1]
    keep = int(np.sum(lam > lam[0] * 1e-3))
    K = max(J, min(keep, 3 * J))
    K = min(K, keep)
    J = min(J, K)
    Q, _ = np.linalg.qr(rng.standard_normal((K, K)))
    sf = ((U[:, :K] / np.sqrt(lam[:K])) @ Q[:, :J]) / scale[:, None]   # unit-variance, decorrelated features
    bias = avg @ sf
    s = E @ sf - bias
    # cuicuilco rescales the slow part so that it is commensurate with the PCA part
    beta_raw, *_ = np.linalg.lstsq(np.column_stack([np.ones(n), s]), x0, rcond=None)
    magn = np.maximum(np.linalg.norm(beta_raw[1:], axis=1), 1e-3 * np.linalg.norm(beta_raw[1:]) + 1e-12)
    s_n = s * magn
    beta, *_ = np.linalg.lstsq(np.column_stack([np.ones(n), s_n]), x0, rcond=None)
    x_app = np.column_stack([np.ones(n), s_n]) @ beta
    res = x0 - x_app
    P = out_dim - J
    r_avg = res.mean(axis=0)
    rc = res - r_avg
    lam_r, V = np.linalg.eigh(rc.T @ rc / (n - 1))
    lam_r, V = lam_r[::-1], V[:, ::-1]
    if P > d:
        raise ValueError("node output %d needs %d residual components from a %d-dim input" % (out_dim, P, d))
    V = V[:, :P]
    r = rc @ V

    sfa_node = new_object("mdp.nodes", "SFANode", sf=sf, avg=avg.reshape(1, -1), _bias=bias.reshape(1, -1),
                          _input_dim=D, _output_dim=J, d=np.ones(J), _dtype=np.dtype("float64"))
    exp_node = new_object("cuicuilco.more_nodes", "GeneralExpansionNode",
                          funcs=[FuncRef(NLE, f) for f in funcs], _input_dim=d, _output_dim=D,
                          exp_output_dim=D, use_pseudoinverse=True, use_hint=False)
    lr_node = new_object("mdp.nodes", "LinearRegressionNode", beta=beta, with_bias=True, use_pinv=False,
                         _input_dim=J, _output_dim=d)
    pca_node = new_object("mdp.nodes", "PCANode", v=V, avg=r_avg.reshape(1, -1), d=lam_r[:P],
                          _input_dim=d, _output_dim=P, output_dim=P)
    node = new_object("cuicuilco.igsfa_node", "iGSFANode", x_mean=x_mean.reshape(1, -1),
                      pre_expansion_node=None, exp_node=exp_node, sfa_node=sfa_node, lr_node=lr_node,
                      pca_node=pca_node, magn_n_sfa_x=magn.reshape(1, -1), num_sfa_features_preserved=J,
                      reconstruct_with_sfa=True, slow_feature_scaling_method="QR_decomposition",
                      delta_threshold=1.99, _input_dim=d, _output_dim=out_dim, _dtype=np.dtype("float64"))
    return node, np.concatenate([s_n, r], axis=1)


def make_flow(spec, seed=0, n_train=None, vary_J=True, clip_sigmas=4.0, verbose=False, train_patches=None):
    """Build and fit a synthetic flow.  Returns an ``mdp.Flow``-shaped object (``.flow`` node list).
    ``train_patches`` (n, w*h) replaces the default band-limited-noise training set."""
    if isinstance(spec, str):
        spec = SPECS[spec]()
    rng = np.random.default_rng(seed)
    w, h = spec["input_xy"]
    dmax = 0
    ch = 1
    for L in spec["layers"]:
        d = L["field"][0] * L["field"][1] * ch
        dmax = max(dmax, ex.expanded_dim(L["funcs"], d))
        ch = L["out_dim"]
    if n_train is None:
        n_train = max(1000, 12 * dmax)
    if train_patches is not None:
        X = np.asarray(train_patches, dtype=np.float64)
        n_train = X.shape[0]
    else:
        X = synthetic_patches(n_train, (w, h), seed + 1).astype(np.float64)
    nodes = []
    ch = 1
    for li, L in enumerate(spec["layers"]):
        conn, (ox, oy) = rect_connections(L["grid_in"], L["field"], L["stride"], ch)
        d = L["field"][0] * L["field"][1] * ch
        n_nodes = ox * oy
        if conn.size != n_nodes * d:
            raise ValueError("layer %d: inconsistent receptive fields" % li)
        sb = new_object("cuicuilco.more_nodes", "PInvSwitchboard", connections=conn,
                        _input_dim=X.shape[1], _output_dim=conn.size, output_dim=conn.size)
        Xg = X[:, conn]
        outs = []
        children = []
        if L["clone"]:
            # one node fitted on the pooled data of all receptive fields, applied everywhere
            pooled = Xg.reshape(-1, d)
            if pooled.shape[0] > 4 * n_train:
                pooled = pooled[rng.choice(pooled.shape[0], 4 * n_train, replace=False)]
            node, _ = _fit_igsfa(pooled, L["funcs"], L["out_dim"], L["J"], rng)
            children = [node] * n_nodes
        for k in range(n_nodes):
            xk = Xg[:, k * d:(k + 1) * d]
            if L["clone"]:
                outs.append(apply_igsfa(children[0], xk))
                continue
            J = L["J"]
            if vary_J and n_nodes > 1:
                J = max(1, J + ((k * 7 + li) % 3) - 1)
            node, yk = _fit_igsfa(xk, L["funcs"], L["out_dim"], J, rng)
            children.append(node)
            outs.append(yk)
        cls = "CloneLayer" if L["clone"] else "Layer"
        layer = new_object("mdp.hinet", cls, nodes=children, _input_dim=conn.size,
                           _output_dim=n_nodes * L["out_dim"])
        if L["clone"]:
            layer.node = children[0]
        nodes += [sb, layer]
        X = np.concatenate(outs, axis=1)
        if clip_sigmas:
            # cuicuilco-style saturation between layers: without it products of outliers grow
            # polynomially from layer to layer on inputs unlike the training set
            lim = float(np.ceil(clip_sigmas * X.std(axis=0).max()))
            nodes.append(new_object("cuicuilco.more_nodes", "PointwiseFunctionNode",
                                    func=FuncRef(NLE, "clip_%d" % int(lim)), _input_dim=X.shape[1],
                                    _output_dim=X.shape[1]))
            X = np.clip(X, -lim, lim)
        ch = L["out_dim"]
        if verbose:
            print("layer %d: %dx%d nodes, d=%d -> %d, feature std %.3g..%.3g"
                  % (li, ox, oy, d, L["out_dim"], X.std(axis=0).min(), X.std(axis=0).max()))
    flow = new_object("mdp.linear_flows", "Flow", flow=nodes, verbose=False)
    flow._synthetic_spec = spec["name"]
    flow._train_output_std = X.std(axis=0)
    return flow


def apply_igsfa(node, x):
    """float64 forward of a fitted node (fitting helper for CloneLayers; mirrors ``_fit_igsfa``)."""
    x0 = x - node.x_mean
    E = _expand(x0, [f.name for f in node.exp_node.funcs])
    s = E @ node.sfa_node.sf - node.sfa_node._bias
    s_n = s * node.magn_n_sfa_x
    x_app = np.column_stack([np.ones(len(x)), s_n]) @ node.lr_node.beta
    r = (x0 - x_app - node.pca_node.avg) @ node.pca_node.v
    return np.concatenate([s_n, r], axis=1)


# ----------------------------------------------------------------------------------------------------
# cache (fitting U11L_64 takes a few seconds; the cache directory travels with the repo snapshot)
# ----------------------------------------------------------------------------------------------------
def cached_flow(spec_name, seed=0, cache_dir=None, **kw):
    from .pickles import dumps, loads
    if cache_dir is None:
        cache_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "flows")
    key = hashlib.sha1(repr((spec_name, seed, sorted(kw.items()), 6)).encode()).hexdigest()[:12]
    path = os.path.join(cache_dir, "%s_%s.pckl" % (spec_name, key))
    if os.path.exists(path):
        with open(path, "rb") as f:
            return loads(f.read())
    flow = make_flow(spec_name, seed=seed, **kw)
    try:
        os.makedirs(cache_dir, exist_ok=True)
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(dumps(flow))
        os.replace(tmp, path)
    except OSError:
        pass
    return flow
