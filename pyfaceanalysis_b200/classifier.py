"""GPU Gaussian classifier: drop-in for the ``classifiers[i]`` entries of the reference.

Reference use (``FaceDetectUpdated.py:709-719``, ``face_analysis.py:1068-1071,1261-1287``)::

    reg_num_signals = classifiers[i].input_dim
    avg_labels = classifiers[i].avg_labels
    reg_out = classifiers[i].regression(sl[:, 0:reg_num_signals], avg_labels)
    value, std = classifiers[i].regression(x, avg_labels, estimate_std=True)

i.e. ``mdp.nodes.GaussianClassifier`` with cuicuilco's ``regression`` patch and the extra ``avg_labels``
attribute.  Built from the unpickled object (``SavedClassifiers/*.pckl`` load with
``pickles.load_obj``); all arithmetic runs in ``csrc/gauss.cu``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class GpuGaussianClassifier(object):
    def __init__(self, clf, device=0):
        self.means = np.ascontiguousarray([np.asarray(m, dtype=np.float64) for m in clf.means])
        self.inv_covs = np.ascontiguousarray([np.asarray(m, dtype=np.float64) for m in clf.inv_covs])
        self._sqrt_def_covs = np.ascontiguousarray([float(v) for v in clf._sqrt_def_covs], dtype=np.float64)
        self.p = np.ascontiguousarray([float(v) for v in clf.p], dtype=np.float64)
        self.labels = list(clf.labels)
        al = getattr(clf, "avg_labels", None)
        self.avg_labels = None if al is None else np.asarray(al, dtype=np.float64)
        C_, D = self.means.shape
        if self.inv_covs.shape != (C_, D, D) or self._sqrt_def_covs.shape != (C_,) or self.p.shape != (C_,):
            raise ValueError("inconsistent GaussianClassifier parameters")
        self.input_dim = int(getattr(clf, "_input_dim", None) or getattr(clf, "input_dim", D))
        if self.input_dim != D:
            raise ValueError("classifier input_dim %d but means have dimension %d" % (self.input_dim, D))
        self.output_dim = self.input_dim
        self.device = int(device)
        self._handle = C.c_void_p()
        _lib.check(_lib.load().hgsfa_gauss_create(_lib.ptr(self.means), _lib.ptr(self.inv_covs),
                                                  _lib.ptr(self._sqrt_def_covs), _lib.ptr(self.p), C_, D,
                                                  self.device, C.byref(self._handle)))

    def _prep(self, x):
        x = np.asarray(x)
        if x.ndim != 2:
            raise ValueError("x has rank %d, should be 2" % x.ndim)
        if x.shape[1] != self.input_dim:
            raise ValueError("x has dimension %d, should be %d" % (x.shape[1], self.input_dim))
        if x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
        if x.strides[1] != x.itemsize or x.strides[0] % x.itemsize:
            x = np.ascontiguousarray(x)
        ld = x.strides[0] // x.itemsize if x.shape[0] > 1 else x.shape[1]
        return x, ld

    def _run(self, x, avg_labels=None, want_value=False, want_std=False, want_winner=False, want_probs=False):
        x, ld = self._prep(x)
        n = x.shape[0]
        value = np.empty(n) if want_value else None
        std = np.empty(n) if want_std else None
        winner = np.empty(n, dtype=np.int32) if want_winner else None
        probs = np.empty((n, len(self.p))) if want_probs else None
        if avg_labels is not None:
            avg_labels = np.ascontiguousarray(avg_labels, dtype=np.float64)
            if avg_labels.shape != (len(self.p),):
                raise ValueError("avg_labels has shape %s, should be (%d,)" % (avg_labels.shape, len(self.p)))
        if n:
            _lib.check(_lib.load().hgsfa_gauss_regress(self._handle, _lib.ptr(x), _lib.dtype_code(x.dtype), n, ld,
                                                       _lib.ptr(avg_labels), _lib.ptr(value), _lib.ptr(std),
                                                       _lib.ptr(winner), _lib.ptr(probs), None))
        return value, std, winner, probs

    # ---- mdp / cuicuilco surface ------------------------------------------------------------
    def regression(self, x, avg_labels=None, estimate_std=False):
        if avg_labels is None:
            avg_labels = self.avg_labels
        value, std, _, _ = self._run(x, avg_labels, want_value=True, want_std=estimate_std)
        return (value, std) if estimate_std else value

    def class_probabilities(self, x):
        return self._run(x, want_probs=True)[3]

    prob = class_probabilities

    def label(self, x):
        winner = self._run(x, want_winner=True)[2]
        return [self.labels[w] for w in winner]

    def execute(self, x):   # ClassifierNode.execute == label in MDP
        return self.label(x)

    @property
    def handle(self):
        return self._handle

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            _lib.load().hgsfa_gauss_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
