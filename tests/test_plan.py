"""Host logic: lowering of MDP / cuicuilco node graphs to fused layer operations.

The numpy plan interpreter (tests/plan_interp.py) executes what the compiler emitted; the oracle executes
the node graph the way the reference does.  Agreement in float64 proves the algebra of the lowering
(gather composition, mean handling, iGSFA two-pass / folded forms, padding, row allocation)."""
import struct

import numpy as np
import pytest

import plan_interp
from oracle import nodes as onodes
from pyfaceanalysis_b200 import expansions as ex
from pyfaceanalysis_b200 import plan, synthetic
from pyfaceanalysis_b200.pickles import FuncRef, new_object

NLE = "cuicuilco.nonlinear_expansion"


def _agree(flow, x, mode="auto", tol=1e-4):
    spec = plan.compile_flow(flow, igsfa_mode=mode)
    y = plan_interp.run_plan(spec, x)
    ref = onodes.flow_execute(flow, x)
    assert y.shape == ref.shape
    scale = max(1.0, np.abs(ref).max())
    # the only float64 difference: exponents are stored as float32 in the term table (0.8f vs 0.8)
    assert np.abs(y - ref).max() <= tol * scale
    return spec


@pytest.mark.parametrize("mode", ["auto", "fold", "two_pass"])
def test_tiny_igsfa_network(tiny_flow, mode):
    x = synthetic.synthetic_patches(64, (16, 16), 4).astype(np.float64)
    spec = _agree(tiny_flow, x, mode)
    assert len(spec.ops) == 4 and spec.output_dim == 16
    assert spec.ops[0].shared and not spec.ops[1].shared        # CloneLayer vs Layer
    blob = plan.serialize(spec)
    assert blob[:8] == b"HGSFAPL2" and struct.unpack("<3q", blob[8:32]) == (256, 16, 4)


def test_u11l_structure(u11l_flow):
    spec = plan.compile_flow(u11l_flow)
    assert [op.n_nodes for op in spec.ops] == [256, 128, 64, 32, 16, 8, 4, 2, 1, 1, 1]
    assert spec.input_dim == 4096 and spec.output_dim == 60
    # switchboards never materialise: 11 Layers -> 11 ops; clip nodes are fused into the epilogues
    assert all(np.isfinite(op.clip).all() for op in spec.ops)
    # per-node J differs inside a layer (iGSFA keeps a data-dependent number of slow features)
    js = {int(n.num_sfa_features_preserved) for n in u11l_flow.flow[4].nodes}
    assert len(js) > 1
    x = synthetic.synthetic_patches(24, (64, 64), 5).astype(np.float64)
    _agree(u11l_flow, x)
    assert spec.alg_flops > 1.2e6 and spec.exe_flops >= spec.alg_flops * 0.9


def _pca(d, m, rng):
    return new_object("mdp.nodes", "PCANode", v=rng.standard_normal((d, m)), avg=rng.standard_normal((1, d)) * 3,
                      d=np.ones(m), _input_dim=d, _output_dim=m, output_dim=m)


def _sfa(d, m, rng):
    sf = rng.standard_normal((d, m)) * 0.1
    avg = rng.standard_normal((1, d))
    return new_object("mdp.nodes", "SFANode", sf=sf, avg=avg, _bias=avg @ sf, _input_dim=d, _output_dim=m)


def _exp(d, funcs):
    return new_object("cuicuilco.more_nodes", "GeneralExpansionNode", funcs=[FuncRef(NLE, f) for f in funcs],
                      _input_dim=d, _output_dim=ex.expanded_dim(funcs, d))


def _layer(nodes):
    return new_object("mdp.hinet", "Layer", nodes=nodes, _input_dim=sum(n._input_dim for n in nodes),
                      _output_dim=sum(n._output_dim for n in nodes))


def _sb(conn, in_dim):
    return new_object("mdp.hinet", "Switchboard", connections=np.asarray(conn), _input_dim=in_dim,
                      _output_dim=len(conn))


def test_cuicuilco_style_separate_layers():
    """switchboard -> Layer(PCA) -> Layer(expansion) -> Layer(SFA): the expansion fuses into the SFA op."""
    rng = np.random.default_rng(0)
    d, n = 12, 4
    conn = rng.permutation(n * d)
    funcs = ["identity", "signed_08expo", "s5QT", "pair_prod_adj2_ex"]
    D = ex.expanded_dim(funcs, 7)
    flow = [_sb(conn, n * d), _layer([_pca(d, 7, rng) for _ in range(n)]), _layer([_exp(7, funcs) for _ in range(n)]),
            _layer([_sfa(D, 5, rng) for _ in range(n)])]
    x = rng.standard_normal((40, n * d)) * 5
    spec = _agree(flow, x)
    assert len(spec.ops) == 2                       # PCA layer, then expansion+SFA fused
    assert spec.ops[1].passes[0]["K_real"] == D


def test_flownode_children_and_head_node():
    rng = np.random.default_rng(1)
    d, n = 10, 3
    funcs = ["identity", "unsigned_08expo", "QT"]
    D = ex.expanded_dim(funcs, 6)
    kids = []
    for _ in range(n):
        inner = new_object("mdp.linear_flows", "Flow", flow=[_pca(d, 6, rng), _exp(6, funcs), _sfa(D, 4, rng)])
        kids.append(new_object("mdp.hinet", "FlowNode", _flow=inner, _input_dim=d, _output_dim=4))
    head = new_object("cuicuilco.more_nodes", "HeadNode", _input_dim=12, _output_dim=7, output_dim=7)
    top = _pca(7, 3, rng)
    flow = [_layer(kids), head, top]
    x = rng.standard_normal((30, n * d)) * 2
    spec = _agree(flow, x)
    assert len(spec.ops) == 2 and len(spec.ops[0].passes) == 2 and spec.ops[0].n_rows >= 6
    assert spec.ops[1].d_in == 7                     # HeadNode became a gather of the first 7 columns


def test_trailing_switchboard_and_standalone_nodes():
    rng = np.random.default_rng(2)
    flow = [_layer([_pca(8, 5, rng), _pca(8, 5, rng)]), _sb([9, 0, 3, 3, 7], 10)]
    x = rng.standard_normal((10, 16))
    spec = _agree(flow, x)
    assert spec.ops[-1].mode == "copy" and spec.output_dim == 5
    # a Switchboard executed on its own (node facade) is a pure copy op
    _agree([_sb(rng.permutation(40)[:33], 40)], rng.standard_normal((6, 40)))


def test_heterogeneous_output_dims_in_one_layer():
    rng = np.random.default_rng(3)
    flow = [_layer([_pca(6, 2, rng), _pca(6, 5, rng), _pca(6, 3, rng)])]
    spec = _agree(flow, rng.standard_normal((12, 18)))
    assert list(spec.ops[0].out_col) == [0, 2, 7] and spec.output_dim == 10


def test_errors_are_loud():
    rng = np.random.default_rng(4)
    bad = _exp(5, ["identity"])
    bad.funcs.append(FuncRef(NLE, "no_such_expansion"))
    with pytest.raises(KeyError, match="no_such_expansion"):
        plan.compile_flow([_layer([bad]), _layer([_sfa(6, 2, rng)])])
    with pytest.raises(plan.UnsupportedFlow):
        plan.compile_flow([new_object("mdp.nodes", "FANode", _input_dim=4, _output_dim=2)])
    with pytest.raises(ValueError, match="should be"):
        plan.compile_flow([_layer([_pca(6, 2, rng)]), _layer([_pca(3, 2, rng)])])     # dimension mismatch
    with pytest.raises(plan.UnsupportedFlow):
        plan.compile_flow([_layer([_exp(5, ["identity"])])])                            # expansion never projected


def test_expansion_tables_match_oracle_functions():
    from oracle import expansions as oexp
    rng = np.random.default_rng(5)
    x = rng.standard_normal((20, 9)) * 3
    for name in ["identity", "QT", "CT", "QE", "TE", "unsigned_08expo", "signed_08expo", "unsigned_06expo",
                 "signed_04expo", "unsigned_sqrt", "signed_sqrt", "abs", "pair_prod_adj1_ex", "pair_prod_adj3_ex",
                 "s4QT", "s6u08ex", "s3CT", "clip_2",
                 # every 0Y exponent the generic rule accepts, not only the hard-coded 08 / 06 / 04
                 "unsigned_02expo", "signed_09expo", "unsigned_01expo", "signed_03expo", "unsigned_05expo",
                 "signed_07expo", "s5_unsigned_09expo", "s5signed_02expo"]:
        t = ex.lower([name], 9)
        got = plan_interp._eval_terms(t, x)
        ref = oexp.resolve(name)(x)
        assert got.shape == ref.shape, name
        assert np.allclose(got, ref, rtol=1e-6, atol=1e-6), name
    assert ex.term_flops(ex.lower(["QT"], 5)) == 15 and ex.term_flops(ex.lower(["identity"], 5)) == 0
    assert ex.terms_for("unsigned_02expo", 2)[0][3] == pytest.approx(0.2) and ex.terms_for("signed_09expo", 2)[0][3] == pytest.approx(0.9)
    for bad in ("unsigned_12expo", "signed_80expo"):      # not the cuicuilco 0Y form: refuse instead of guessing
        with pytest.raises(KeyError):
            ex.terms_for(bad, 3)


def test_tile_choice_and_segments():
    assert plan._choose_tile(13) == (16, 1) and plan._choose_tile(20) == (20, 1)
    assert plan._choose_tile(12) == (12, 1) and plan._choose_tile(27) == (28, 1)
    for n in range(1, 129):         # column tiles come in multiples of 4: never more than 3 padded columns up to 32
        nt, ntl = plan._choose_tile(n)
        assert nt % 4 == 0 and 8 <= nt <= 32 and ntl in (1, 2, 4) and nt * ntl >= n
        if 8 <= n <= 32:
            assert nt * ntl - n < 4
    assert plan._choose_tile(60)[0] * plan._choose_tile(60)[1] == 64
    with pytest.raises(plan.UnsupportedFlow):
        plan._choose_tile(300)
    t = ex.lower(["identity", "unsigned_08expo", "s3QT"], 6)
    segs = plan._segments(t, 6)
    assert [(s[0], s[1], s[2], s[5]) for s in segs] == [(ex.OP_ID, 0, 6, 0), (ex.OP_ABSPOW, 6, 12, 0), (ex.OP_MUL, 12, 18, -1)]
    fused = plan._fuse_id_pow(segs)
    assert fused[0][0] == plan.OP_ID_POW and fused[0][1:3] == (0, 12) and len(fused) == 2
    assert plan._gather_runs([4, 5, 6, 10, 11, 3]) == [(0, 4, 3), (3, 10, 2), (5, 3, 1)]


def test_tensor_core_configuration(u11l_flow, monkeypatch):
    """Tensor-core ops: 8-aligned segments, tensor-memory columns and shared memory inside the hardware budgets."""
    monkeypatch.setattr(plan, "ENGINE", "auto")
    spec = plan.compile_flow(u11l_flow)
    assert all(op.engine == "tc" for op in spec.ops)
    for op in spec.ops:
        t, ps = op.tc, op.passes[0]
        assert t["Kpad"] % 8 == 0 and t["Kpad"] >= ps["K"] and t["Npad16"] % 16 == 0 and t["Npad16"] >= ps["N_real"]
        cols = t["nd"] * op.twc * t["Npad16"] + t["na"] * 2 * plan.TC_CK
        assert cols <= t["tmem_cols"] <= 512 and t["smem"] <= plan.SMEM_LIMIT and 1 <= op.twc <= 8
        assert t["n_chunks"] == -(-t["Kpad"] // plan.TC_CK)
    pieces, kpad = plan._tc_segments([(ex.OP_ID, 0, 26, 0.0, 0, 0), (ex.OP_ABSPOW, 26, 52, 0.8, 0, 0)], 52)
    assert kpad == 64 and [(a, b) for _, a, b in pieces] == [(0, 32), (32, 64)]
    pieces, kpad = plan._tc_segments([(plan.OP_ID_POW, 0, 108, 0.8, 0, 0), (ex.OP_MUL, 108, 163, 0.0, 0, -1)], 163)
    assert kpad == 168 and all(a % 8 == 0 and b % 8 == 0 and a // plan.TC_CK == (b - 1) // plan.TC_CK for _, a, b in pieces)
    # the numpy plan interpreter does not care about the engine: same algebra
    x = synthetic.synthetic_patches(8, (64, 64), 5).astype(np.float64)
    monkeypatch.setattr(plan, "ENGINE", "ffma")
    spec_f = plan.compile_flow(u11l_flow)
    assert all(op.engine == "ffma" for op in spec_f.ops)
    assert np.allclose(plan_interp.run_plan(spec, x), plan_interp.run_plan(spec_f, x), rtol=1e-9, atol=1e-9)
    blob = plan.serialize(spec)
    assert (struct.unpack_from("<q", blob, 64 + 5 * 8)[0] >> 16) & 0xff == 1        # engine flag of op 0


def _igsfa_node(rng, d=9, k_pre=None, J=3, P=4, funcs=("identity", "unsigned_08expo")):
    """A hand-made iGSFA node (random parameters) with every optional part present."""
    pre = None
    d_exp_in = d
    if k_pre:
        pre = _pca(d, k_pre, rng)              # pre_expansion_node with a NON-ZERO mean (plan.py:235)
        d_exp_in = k_pre
    D = ex.expanded_dim(list(funcs), d_exp_in)
    sfa = _sfa(D, J + 2, rng)
    lr = new_object("mdp.nodes", "LinearRegressionNode", beta=rng.standard_normal((J + 1, d)), with_bias=True,
                    _input_dim=J, _output_dim=d)
    pca = _pca(d, P, rng)
    return new_object("cuicuilco.igsfa_node", "iGSFANode", x_mean=rng.standard_normal((1, d)) * 2,
                      pre_expansion_node=pre, exp_node=_exp(d_exp_in, list(funcs)), sfa_node=sfa, lr_node=lr, pca_node=pca,
                      magn_n_sfa_x=(0.5 + rng.random((1, J + 2))) * 3, num_sfa_features_preserved=J,
                      reconstruct_with_sfa=True, _input_dim=d, _output_dim=J + P)


@pytest.mark.parametrize("mode", ["auto", "fold", "two_pass"])
def test_igsfa_pre_expansion_node_keeps_x_mean(mode):
    """The pre-expansion PCA's mean must not leak into the shared input offset: the PCA / residual branch reads
    x0 = x - x_mean, not x - x_mean - avg_pre."""
    rng = np.random.default_rng(11)
    nodes = [_igsfa_node(rng, d=9, k_pre=5) for _ in range(3)]
    x = rng.standard_normal((50, 27)) * 3
    spec = _agree([_layer(nodes)], x, mode)
    assert np.allclose(spec.ops[0].in_offset, np.concatenate([n.x_mean for n in nodes]))


@pytest.mark.parametrize("lr_input", ["scaled", "unscaled"])
@pytest.mark.parametrize("mode", ["fold", "two_pass"])
def test_igsfa_lr_input_switch(monkeypatch, mode, lr_input):
    """Both readings of cuicuilco's iGSFA reconstruction (lr_node on the rescaled or on the raw slow features) are
    lowered; oracle and compiler flip together, and the two readings really differ."""
    rng = np.random.default_rng(12)
    nodes = [_igsfa_node(rng, d=8) for _ in range(2)]
    x = rng.standard_normal((40, 16)) * 3
    ref_scaled = onodes.flow_execute([_layer(nodes)], x)
    monkeypatch.setattr(onodes, "IGSFA_LR_INPUT", lr_input)
    monkeypatch.setattr(plan, "IGSFA_LR_INPUT", lr_input)
    _agree([_layer(nodes)], x, mode)
    ref = onodes.flow_execute([_layer(nodes)], x)
    assert (lr_input == "scaled") == bool(np.allclose(ref, ref_scaled))
    assert np.allclose(ref[:, :3], ref_scaled[:, :3])       # the slow part is the same under both readings


def test_fused_front_tables_match_unfused_ops(u11l_flow):
    """The fused front (layers 0-2 in one kernel) streams its own weight chunks and tables; executed in numpy the way
    the kernel walks them they reproduce the three unfused ops (FP16 hi + lo weights: 22 bits)."""
    import front_interp
    from pyfaceanalysis_b200 import front
    spec = plan.compile_flow(u11l_flow)
    f = plan.plan_front(spec)
    assert f is not None, spec.front_reason
    assert (f.n_sub, f.np1, f.np2, f.nn, f.nch) == (64, 16, 24, (16, 32, 32), (1, 2, 3))
    x = synthetic.synthetic_patches(6, (64, 64), 7).astype(np.float64)
    ref = x
    for op in spec.ops[:3]:
        ref = plan_interp.run_op(op, ref)
    got = front_interp.run_front(f, x)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    blob = plan.serialize(spec)
    off, size = struct.unpack("<2q", blob[32:48])
    assert off % 16 == 0 and off + size == len(blob) and blob[off:off + 8] == front.MAGIC
    # flows that do not match the pattern keep the per-layer path and say why
    tiny = synthetic.cached_flow("tiny", seed=0)
    ts = plan.compile_flow(tiny)
    assert plan.plan_front(ts) is None and ts.front_reason
    assert struct.unpack("<2q", plan.serialize(ts)[32:48]) == (0, 0)
