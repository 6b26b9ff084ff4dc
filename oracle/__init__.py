"""CPU oracle for the HiGSFA sliding-window hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

float64 numpy restatement of the reference path (AlbertoEsc/PyFaceAnalysis):
window grid -> EXTENT crop -> MDP/cuicuilco flow ``execute`` -> GaussianClassifier regression ->
cascade controller.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package; nothing under
``pyfaceanalysis_b200/`` does.

Pinning status:

* grid / controller : restated line by line from ``face_analysis.py:575-669, 803-887`` (source in tree).
* crop              : pinned against Pillow 12.2 ``Image.transform(EXTENT, NEAREST|BILINEAR|BICUBIC)`` run live in
                      the tests plus the fixtures under ``tests/golden/`` (Pillow's C core is what the
                      reference calls).
* Gaussian head     : pinned against the 19 shipped ``SavedClassifiers/*.pckl`` parameter sets through
                      known-answer identities (``tests/golden/classifiers.npz``, made by
                      ``tools/make_golden.py``).
* flow nodes        : **PARITY UNPINNED** -- ``mdp`` and ``cuicuilco`` (commit 9bfd242...) are not vendored
                      in the reference, are not installable here, the reference ships no tests or golden
                      vectors, and the ``SavedNetworks/*.pckl`` flows were stripped.  The node semantics
                      below restate the published MDP / cuicuilco algorithms and become normative for
                      this build.
"""
