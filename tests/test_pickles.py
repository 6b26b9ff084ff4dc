"""Model-store loader: stub unpickling of Python-2 MDP / cuicuilco pickles."""
import os
import pickle

import numpy as np
import pytest

from pyfaceanalysis_b200 import pickles, synthetic

REF = "/root/reference/SavedClassifiers"


def test_synthetic_flow_round_trip_has_real_pickle_shape(tiny_flow):
    data = pickles.dumps(tiny_flow)
    # GLOBAL opcodes name the upstream classes, exactly like a SavedNetworks/*.pckl
    for token in (b"cmdp.linear_flows\nFlow\n", b"cmdp.hinet\nCloneLayer\n", b"ccuicuilco.igsfa_node\niGSFANode\n",
                  b"ccuicuilco.nonlinear_expansion\nunsigned_08expo\n", b"cmdp.nodes\nSFANode\n"):
        assert token in data
    back = pickles.loads(data)
    assert pickles.class_path(back) == "mdp.linear_flows.Flow"
    n0, n1 = tiny_flow.flow[1].nodes[0], back.flow[1].nodes[0]
    assert back.flow[1].nodes[0] is back.flow[1].nodes[1]            # CloneLayer sharing survives (memo)
    assert np.array_equal(n0.sfa_node.sf, n1.sfa_node.sf)
    f = n1.exp_node.funcs[1]
    assert isinstance(f, pickles.FuncRef) and f.name == "unsigned_08expo"
    with pytest.raises(RuntimeError):
        f(np.zeros((1, 1)))


def test_legacy_module_aliases():
    # FaceDetectUpdated.py:57-68: old pickles say "more_nodes", "nonlinear_expansion", "GSFA_node", ...
    blob = b"\x80\x02cmore_nodes\nGeneralExpansionNode\nq\x00)\x81q\x01}q\x02U\x05funcsq\x03]q\x04cnonlinear_expansion\nQT\nq\x05asb."
    obj = pickles.loads(blob)
    assert pickles.class_path(obj) == "cuicuilco.more_nodes.GeneralExpansionNode"
    assert obj.funcs[0] == pickles.FuncRef("cuicuilco.nonlinear_expansion", "QT")
    assert pickles.canonical_module("GSFA_node") == "cuicuilco.gsfa_node"
    assert pickles.canonical_module("imageLoader") == "cuicuilco.image_loader"


def test_foreign_globals_are_refused():
    with pytest.raises(pickle.UnpicklingError, match="refusing"):
        pickles.loads(pickle.dumps(os.getcwd, protocol=2))
    # a REDUCE on any of these is arbitrary code execution: builtins and numpy are allow-listed by exact name only
    for mod, name in (("__builtin__", "eval"), ("builtins", "exec"), ("__builtin__", "getattr"), ("builtins", "__import__"),
                      ("os", "system"), ("posix", "system"), ("numpy.testing._private.utils", "runstring"),
                      ("numpy", "load"), ("subprocess", "Popen")):
        blob = b"\x80\x02c" + mod.encode() + b"\n" + name.encode() + b"\n."
        with pytest.raises(pickle.UnpicklingError, match="refusing"):
            pickles.loads(blob)
    # what a real pickle does need still resolves (Python-2 names included)
    blob = b"\x80\x02c__builtin__\nset\n]q\x00(K\x01K\x02e\x85Rq\x01."
    assert pickles.loads(blob) == {1, 2}
    assert pickles.loads(b"\x80\x02c__builtin__\nlong\n.") is int


def test_load_obj_none_sentinel(tmp_path):
    assert pickles.load_obj(str(tmp_path), "None0") is None            # face_analysis.py:456
    p = tmp_path / "x.pckl"
    p.write_bytes(pickles.dumps(pickles.new_object("mdp.nodes", "PCANode", v=np.eye(2))))
    assert np.array_equal(pickles.load_obj(str(tmp_path), "x").v, np.eye(2))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_real_classifier_pickles_match_golden(classifiers):
    names = sorted(f for f in os.listdir(REF) if f.endswith(".pckl"))
    assert len(names) == 19
    for fn, g in zip(names, classifiers):
        clf = pickles.load_obj(REF, fn[:-5])
        assert pickles.class_path(clf) == "mdp.nodes.GaussianClassifier" and g.name == fn
        assert clf._input_dim == g.input_dim and str(clf._dtype) == "float64"
        assert np.array_equal(np.asarray(clf.means), np.asarray(g.means))
        assert np.array_equal(np.asarray(clf.inv_covs), np.asarray(g.inv_covs))
        assert np.array_equal(np.asarray(clf.avg_labels), g.avg_labels)
