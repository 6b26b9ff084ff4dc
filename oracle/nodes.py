"""Oracle (test infrastructure): float64 ``execute`` of the MDP / cuicuilco node vocabulary.

Follows the call sites ``FaceDetectUpdated.py:699`` and ``face_analysis.py:1064,1257``
(``flow.execute(x, benchmark=...)``).  The node classes live in un-vendored packages
(mdp master of 2018-03, cuicuilco @9bfd242) -- **PARITY UNPINNED**: each function restates
the published algorithm of the class it is named after, operating on the attribute names those
classes pickle.  Objects are duck-typed on ``type(node).__name__``.
"""
import numpy as np

from . import expansions as _exp


def _name(node):
    return type(node).__name__


def _get(node, *names, default=None):
    for n in names:
        if hasattr(node, n):
            return getattr(node, n)
    return default


# ------------------------------------------------------------------ leaf nodes
def pca_execute(node, x):
    """mdp.nodes.PCANode / WhiteningNode: ``mult(x - avg, v)`` (whitening: v pre-scaled)."""
    v = np.asarray(node.v, dtype=np.float64)
    out_dim = _get(node, "output_dim", "_output_dim")
    if out_dim is not None:
        v = v[:, :out_dim]
    return (x - np.asarray(node.avg, dtype=np.float64).reshape(1, -1)) @ v


def sfa_execute(node, x):
    """mdp.nodes.SFANode (and cuicuilco GSFANode at execute time): ``mult(x, sf) - _bias``."""
    sf = np.asarray(node.sf, dtype=np.float64)
    bias = _get(node, "_bias")
    if bias is None:
        bias = np.asarray(node.avg, dtype=np.float64).reshape(1, -1) @ sf
    bias = np.asarray(bias, dtype=np.float64).reshape(1, -1)
    out_dim = _get(node, "output_dim", "_output_dim")
    if out_dim is not None:
        sf, bias = sf[:, :out_dim], bias[:, :out_dim]
    return x @ sf - bias


def linreg_execute(node, x):
    """mdp.nodes.LinearRegressionNode: optional constant column of ones first, then ``mult(x, beta)``."""
    beta = np.asarray(node.beta, dtype=np.float64)
    if _get(node, "with_bias", default=True):
        x = np.concatenate((np.ones((x.shape[0], 1)), x), axis=1)
    return x @ beta


def expansion_execute(node, x):
    """cuicuilco GeneralExpansionNode: concatenate ``f(x)`` over the stored function list."""
    return np.concatenate([_exp.resolve(f)(x) for f in node.funcs], axis=1)


def switchboard_execute(node, x):
    """mdp.hinet.Switchboard (+ Rectangular2dSwitchboard, cuicuilco PInvSwitchboard): ``x[:, connections]``."""
    return x[:, np.asarray(node.connections, dtype=np.int64)]


def layer_execute(node, x):
    """mdp.hinet.Layer / CloneLayer: consecutive input slices, one per node, outputs concatenated."""
    nodes = list(node.nodes)
    outs = []
    start = 0
    for sub in nodes:
        d = input_dim(sub)
        outs.append(execute(sub, x[:, start:start + d]))
        start += d
    if start != x.shape[1]:
        raise ValueError("Layer: input has %d columns, nodes consume %d" % (x.shape[1], start))
    return np.concatenate(outs, axis=1)


def same_input_layer_execute(node, x):
    return np.concatenate([execute(sub, x) for sub in node.nodes], axis=1)


def flownode_execute(node, x):
    flow = _get(node, "_flow", "flow")
    return flow_execute(flow, x)


# "scaled": x_app = lr_node(sfa_x[:, :J] * magn_n_sfa_x);  "unscaled": x_app = lr_node(sfa_x[:, :J]).
# Tests flip this together with pyfaceanalysis_b200.plan.IGSFA_LR_INPUT; real pickles decide which one is right.
IGSFA_LR_INPUT = __import__("os").environ.get("HGSFA_IGSFA_LR_INPUT", "scaled")


def igsfa_execute(node, x):
    """cuicuilco iGSFANode (older pickles: IEVMLRecNode), SURVEY.md row a-11:

    ``x0 = x - x_mean``; ``e = exp_node(pre_expansion_node(x0))``; ``s = sfa_node(e)``;
    ``s_n = s[:, :J] * magn_n_sfa_x``; ``x_app = lr_node(s_n)`` (if ``reconstruct_with_sfa``);
    ``r = pca_node(x0 - x_app)``; output ``[s_n, r]``.
    """
    x0 = x - np.asarray(node.x_mean, dtype=np.float64).reshape(1, -1)
    pre = _get(node, "pre_expansion_node")
    xp = execute(pre, x0) if pre is not None else x0
    expn = _get(node, "exp_node")
    e = execute(expn, xp) if expn is not None else xp
    s = execute(node.sfa_node, e)
    J = int(node.num_sfa_features_preserved)
    magn = np.asarray(_get(node, "magn_n_sfa_x", default=1.0), dtype=np.float64).reshape(-1)
    s_n = s[:, :J] * (magn if magn.size == 1 else magn[:J])          # n_sfa_x = sfa_x * magn_n_sfa_x, first J kept
    if _get(node, "reconstruct_with_sfa", default=True) and J > 0:
        # UNPINNED detail (cuicuilco source unavailable): is the linear reconstruction fed with the rescaled
        # slow features n_sfa_x (our reading: lr_node is trained on them) or with the raw sfa_node output?
        # IGSFA_LR_INPUT selects; the product's lowering has the same switch (plan.IGSFA_LR_INPUT).
        x_app = execute(node.lr_node, s_n if IGSFA_LR_INPUT == "scaled" else s[:, :J])
    else:
        x_app = 0.0
    pca = _get(node, "pca_node")
    if pca is None:
        return s_n
    r = execute(pca, x0 - x_app)
    return np.concatenate((s_n, r), axis=1)


def identity_execute(node, x):
    return x


def head_execute(node, x):
    """cuicuilco HeadNode: keep the first ``output_dim`` components."""
    return x[:, :_get(node, "output_dim", "_output_dim")]


def pointwise_execute(node, x):
    """cuicuilco PointwiseFunctionNode: apply the stored function element-wise."""
    return _exp.resolve(node.func)(x)


_DISPATCH = {
    "PCANode": pca_execute, "WhiteningNode": pca_execute,
    "SFANode": sfa_execute, "GSFANode": sfa_execute, "SFAPCANode": sfa_execute,
    "LinearRegressionNode": linreg_execute,
    "GeneralExpansionNode": expansion_execute,
    "Switchboard": switchboard_execute, "Rectangular2dSwitchboard": switchboard_execute,
    "PInvSwitchboard": switchboard_execute, "DoubleRect2dSwitchboard": switchboard_execute,
    "Layer": layer_execute, "CloneLayer": layer_execute,
    "SameInputLayer": same_input_layer_execute,
    "FlowNode": flownode_execute,
    "iGSFANode": igsfa_execute, "IEVMLRecNode": igsfa_execute,
    "IdentityNode": identity_execute,
    "HeadNode": head_execute,
    "PointwiseFunctionNode": pointwise_execute,
}


def input_dim(node):
    d = _get(node, "_input_dim", "input_dim")
    if d is None:
        raise ValueError("node %s has no input_dim" % _name(node))
    return int(d)


def execute(node, x):
    fn = _DISPATCH.get(_name(node))
    if fn is None:
        raise KeyError("oracle: no execute() restatement for node class %r" % _name(node))
    x = np.asarray(x, dtype=np.float64)
    d = _get(node, "_input_dim", "input_dim")
    if d is not None and x.shape[1] != int(d):
        # MDP's _pre_execution_checks raises on a dimension mismatch
        raise ValueError("%s: x has dimension %d, should be %d" % (_name(node), x.shape[1], int(d)))
    return fn(node, x)


def flow_nodes(flow):
    nodes = _get(flow, "flow")
    if nodes is None:
        nodes = list(flow)
    return list(nodes)


def flow_execute(flow, x, nodenr=None, benchmark=None):
    """mdp.Flow.execute: ``x = node.execute(x)`` over the node list (cuicuilco's patch adds the ignored
    ``benchmark=`` keyword used at ``FaceDetectUpdated.py:699``)."""
    x = np.asarray(x, dtype=np.float64)
    nodes = flow_nodes(flow)
    if nodenr is None:
        nodenr = len(nodes) - 1
    for node in nodes[:nodenr + 1]:
        x = execute(node, x)
    return x
