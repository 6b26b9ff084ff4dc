"""Oracle (test infrastructure): ``mdp.nodes.GaussianClassifier`` + cuicuilco's ``.regression`` patch.

Call sites: ``FaceDetectUpdated.py:709-719``, ``face_analysis.py:1068-1071,1261-1287``.  mdp / cuicuilco
are not vendored; the formula is pinned by known-answer identities on the 19 shipped parameter sets
(SURVEY.md 8c / Appendix B.2): ``det(inv_covs[c]) ** -0.5 == _sqrt_def_covs[c]`` and
``regression(means) ~ avg_labels`` -- see ``tests/test_oracle_gauss.py``.

The class loop, the ``(2 pi) ** (-D / 2) / sqrt_det`` constant, the un-normalised ``exp`` and the final
division are evaluated exactly in this order in float64, so the reference's underflow behaviour
(all classes underflow -> 0/0 -> NaN) is reproduced, not "fixed".
"""
import numpy as np


def _params(clf):
    means = np.asarray([np.asarray(m, dtype=np.float64) for m in clf.means])
    inv_covs = np.asarray([np.asarray(m, dtype=np.float64) for m in clf.inv_covs])
    sqrt_det = np.asarray([float(v) for v in clf._sqrt_def_covs], dtype=np.float64)
    p = np.asarray([float(v) for v in clf.p], dtype=np.float64)
    return means, inv_covs, sqrt_det, p


def class_probabilities(clf, x):
    x = np.asarray(x, dtype=np.float64)
    means, inv_covs, sqrt_det, p = _params(clf)
    C, D = means.shape
    if x.ndim != 2 or x.shape[1] != D:
        raise ValueError("GaussianClassifier: x has dimension %s, should be %d" % (x.shape[1:], D))
    prob = np.zeros((x.shape[0], C))
    with np.errstate(under="ignore", invalid="ignore", divide="ignore"):
        for c in range(C):
            x_mn = x - means[c][np.newaxis, :]
            exponent = 0.5 * ((x_mn @ inv_covs[c]) * x_mn).sum(axis=1)
            constant = (2.0 * np.pi) ** (-D / 2.0) / sqrt_det[c]
            prob[:, c] = constant * np.exp(-exponent)
            prob[:, c] *= p[c]
        tot = prob.sum(axis=1)[:, np.newaxis]
        return prob / tot


def regression(clf, x, avg_labels, estimate_std=False):
    """cuicuilco GaussianRegression: posterior-weighted mean label (and its std)."""
    avg_labels = np.asarray(avg_labels, dtype=np.float64)
    prob = class_probabilities(clf, x)
    value = prob @ avg_labels
    if estimate_std:
        diff = (avg_labels.reshape(1, -1) - value.reshape(-1, 1)) ** 2
        with np.errstate(invalid="ignore"):
            std = np.sqrt((prob * diff).sum(axis=1))
        return value, std
    return value


def label(clf, x):
    """MAP label (``GaussianClassifier._label``): ``labels[argmax_c P(c|x)]``."""
    prob = class_probabilities(clf, x)
    winner = prob.argmax(axis=-1)
    return [clf.labels[w] for w in winner]


def log_domain_posterior(clf, x):
    """Numerically safe posterior (never NaN) -- used by tests to decide which rows of the
    reference formula are in the underflow regime; not part of the reference behaviour."""
    x = np.asarray(x, dtype=np.float64)
    means, inv_covs, sqrt_det, p = _params(clf)
    C, D = means.shape
    logq = np.zeros((x.shape[0], C))
    for c in range(C):
        x_mn = x - means[c][np.newaxis, :]
        logq[:, c] = (np.log(p[c]) - 0.5 * D * np.log(2 * np.pi) - np.log(sqrt_det[c])
                      - 0.5 * ((x_mn @ inv_covs[c]) * x_mn).sum(axis=1))
    m = logq.max(axis=1, keepdims=True)
    w = np.exp(logq - m)
    return w / w.sum(axis=1, keepdims=True), logq
