"""Pins the oracle's Gaussian head on the 19 shipped classifier parameter sets."""
import numpy as np

from conftest import class_samples
from oracle import gauss as ogauss


def test_sqrt_det_identity(classifiers):
    """_sqrt_def_covs == det(inv_covs) ** -0.5  (what MDP stores; <= 4e-14 relative in the survey)."""
    assert len(classifiers) == 19
    for clf in classifiers:
        for ic, sd in zip(clf.inv_covs, clf._sqrt_def_covs):
            sign, logdet = np.linalg.slogdet(ic)
            assert sign > 0
            assert abs(np.exp(-0.5 * logdet) / sd - 1.0) < 1e-10, clf.name


def test_priors_and_shapes(classifiers):
    dims = sorted((c.input_dim, len(c.p)) for c in classifiers)
    # D x C of the shipped heads
    assert (9, 10) in dims and (20, 50) in dims and (4, 39) in dims and (5, 2) in dims and (12, 50) in dims
    for clf in classifiers:
        assert abs(sum(clf.p) - 1.0) < 1e-9
        assert clf.avg_labels.shape == (len(clf.p),)


def test_regression_at_class_means_two_class_heads(classifiers):
    """Posterior-weighted avg_labels evaluated at the class means reproduce the labels (exactly for the
    well-separated 2-class heads; SURVEY.md App. B.2: max error 1.6e-13 .. 1.4e-3)."""
    two = [c for c in classifiers if len(c.p) == 2]
    assert len(two) == 4
    for clf in two:
        reg = ogauss.regression(clf, np.asarray(clf.means), clf.avg_labels)
        assert np.abs(reg - clf.avg_labels).max() < 2e-3, clf.name


def test_regression_is_posterior_mean(classifiers):
    rng = np.random.default_rng(0)
    for clf in classifiers[:6]:
        x = class_samples(clf, 3, rng)
        P = ogauss.class_probabilities(clf, x)
        assert np.allclose(P.sum(axis=1), 1.0)
        v, s = ogauss.regression(clf, x, clf.avg_labels, estimate_std=True)
        assert np.allclose(v, P @ clf.avg_labels)
        assert np.allclose(s ** 2, (P * (clf.avg_labels[None, :] - v[:, None]) ** 2).sum(axis=1))
        lab = ogauss.label(clf, x)
        assert lab == [clf.labels[k] for k in P.argmax(axis=1)]
        # the numerically safe posterior agrees wherever the direct formula has not underflowed
        Ps, _ = ogauss.log_domain_posterior(clf, x)
        assert np.allclose(P, Ps, atol=1e-12)


def test_underflow_radius_gives_nan(classifiers):
    """All class likelihoods underflow beyond ~35-38 Mahalanobis units -> 0/0 = NaN."""
    clf = classifiers[1]
    mu = np.asarray(clf.means[0])
    cov = np.linalg.inv(clf.inv_covs[0])
    w, V = np.linalg.eigh((cov + cov.T) / 2)
    direction = V[:, -1] * np.sqrt(w[-1])          # one Mahalanobis unit along the widest axis of class 0
    near = mu + 3.0 * direction
    far = mu + 4000.0 * direction
    out = ogauss.regression(clf, np.stack([near, far]), clf.avg_labels)
    assert np.isfinite(out[0]) and np.isnan(out[1])
    _, logq = ogauss.log_domain_posterior(clf, np.stack([near, far]))
    assert logq[1].max() < -745.0 < logq[0].max()
