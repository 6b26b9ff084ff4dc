"""GPU parity of the Gaussian-classifier head (csrc/gauss.cu) against the float64 oracle on the REAL
shipped classifier parameters (tests/golden/classifiers.npz)."""
import numpy as np
import pytest

from conftest import class_samples
from oracle import gauss as ogauss

pytestmark = pytest.mark.gpu


def test_regression_all_shipped_classifiers(classifiers):
    from pyfaceanalysis_b200 import GpuGaussianClassifier
    rng = np.random.default_rng(0)
    for clf in classifiers:
        g = GpuGaussianClassifier(clf)
        assert g.input_dim == clf.input_dim
        x = class_samples(clf, 8, rng, spread=1.5)
        ref = ogauss.regression(clf, x, clf.avg_labels)
        got = g.regression(x, clf.avg_labels)
        scale = np.abs(clf.avg_labels).max()
        assert np.abs(got - ref).max() <= 1e-9 * scale, clf.name
        ref_v, ref_s = ogauss.regression(clf, x, clf.avg_labels, estimate_std=True)
        got_v, got_s = g.regression(x, clf.avg_labels, estimate_std=True)
        assert np.abs(got_v - ref_v).max() <= 1e-9 * scale
        assert np.abs(got_s - ref_s).max() <= 1e-7 * scale
        P = g.class_probabilities(x)
        assert np.abs(P - ogauss.class_probabilities(clf, x)).max() <= 1e-9
        assert g.label(x) == ogauss.label(clf, x)
        # float32 features (what the flow kernels hand over)
        got32 = g.regression(x.astype(np.float32), clf.avg_labels)
        ref32 = ogauss.regression(clf, x.astype(np.float32).astype(np.float64), clf.avg_labels)
        assert np.abs(got32 - ref32).max() <= 1e-9 * scale
        g.close()


def test_underflow_gives_nan_like_reference(classifiers):
    """All class likelihoods underflow -> 0/0 -> NaN in the reference; the kernel must agree row by row."""
    from pyfaceanalysis_b200 import GpuGaussianClassifier
    rng = np.random.default_rng(1)
    clf = classifiers[1]   # a Disc head (9 x 10)
    g = GpuGaussianClassifier(clf)
    x = class_samples(clf, 20, rng)
    far = x + rng.standard_normal(x.shape) * 3000.0
    mix = np.concatenate([x, far])
    ref = ogauss.regression(clf, mix, clf.avg_labels)
    got = g.regression(mix, clf.avg_labels)
    assert np.isnan(ref).any() and not np.isnan(ref).all()
    assert np.array_equal(np.isnan(ref), np.isnan(got))
    ok = ~np.isnan(ref)
    assert np.abs(got[ok] - ref[ok]).max() <= 1e-9
    g.close()


def test_shapes_and_errors(classifiers):
    from pyfaceanalysis_b200 import GpuGaussianClassifier
    clf = classifiers[0]
    g = GpuGaussianClassifier(clf)
    assert g.regression(np.zeros((0, g.input_dim)), clf.avg_labels).shape == (0,)
    with pytest.raises(ValueError):
        g.regression(np.zeros((4, g.input_dim + 1)), clf.avg_labels)
    # strided view, like sl[:, 0:D] of a wider feature matrix
    wide = np.random.default_rng(2).standard_normal((33, g.input_dim + 7)) * 100
    view = wide[:, :g.input_dim]
    assert np.allclose(g.regression(view, clf.avg_labels), ogauss.regression(clf, view, clf.avg_labels),
                       rtol=1e-9, atol=1e-9, equal_nan=True)
    g.close()
