import json
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_classifiers():
    """The 19 shipped GaussianClassifier parameter sets as attribute bags (tests/golden/classifiers.npz)."""
    z = np.load(os.path.join(GOLDEN, "classifiers.npz"))
    out = []
    for k, name in enumerate(z["names"]):
        key = "c%02d" % k
        clf = types.SimpleNamespace(
            name=str(name),
            means=list(z[key + "_means"]),
            inv_covs=list(z[key + "_inv_covs"]),
            _sqrt_def_covs=list(z[key + "_sqrt_def_covs"]),
            p=list(z[key + "_p"]),
            labels=list(z[key + "_labels"]),
            avg_labels=z[key + "_avg_labels"],
        )
        clf._input_dim = clf.means[0].shape[0]
        clf.input_dim = clf._input_dim
        out.append(clf)
    return out


@pytest.fixture(scope="session")
def classifiers():
    return load_classifiers()


@pytest.fixture(scope="session")
def pipeline():
    with open(os.path.join(GOLDEN, "pipeline.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def crop_golden():
    return np.load(os.path.join(GOLDEN, "crop_golden.npz"))


@pytest.fixture(scope="session")
def grid_golden():
    with open(os.path.join(GOLDEN, "grid_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def tiny_flow():
    from pyfaceanalysis_b200 import synthetic
    return synthetic.make_flow("tiny", seed=3)


@pytest.fixture(scope="session")
def u11l_flow():
    from pyfaceanalysis_b200 import synthetic
    return synthetic.cached_flow("U11L_64", seed=0)


def class_samples(clf, n_per_class, rng, spread=1.0):
    """Points drawn around the class means with the class covariances (in-distribution head inputs)."""
    xs = []
    for mu, ic in zip(clf.means, clf.inv_covs):
        cov = np.linalg.inv(ic)
        cov = (cov + cov.T) / 2
        xs.append(rng.multivariate_normal(mu, cov * spread, size=n_per_class, method="eigh"))
    return np.concatenate(xs, axis=0)


@pytest.fixture(scope="session")
def u11l96_flow():
    from pyfaceanalysis_b200 import synthetic
    return synthetic.cached_flow("U11L_96", seed=1, n_train=1500)
