"""Host-side purge (reference face_analysis.py:186-221): vectorised form == the oracle's loop form."""
import numpy as np

from oracle import controller as ctl
from pyfaceanalysis_b200.cascade import purge_detections


def test_purge_matches_oracle_loop():
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 7, 40):
        det = np.zeros((n, 10))
        c = rng.uniform(50, 400, size=(n, 2))
        c[n // 2:] = c[:n - n // 2] + rng.normal(0, 2.0, size=(n - n // 2, 2))     # near-duplicates
        half = rng.uniform(10, 40, size=n)
        det[:, 0:2], det[:, 2:4] = c - half[:, None], c + half[:, None]
        det[:, 4] = rng.uniform(-20, 20, size=n)
        det[:, 5:7] = c + np.stack([-0.37 * half, -0.25 * half], axis=1)
        det[:, 7:9] = c + np.stack([0.37 * half, -0.25 * half], axis=1)
        det[:, 9] = rng.uniform(0, 0.5, size=n)
        got = purge_detections(det)
        ref = np.array(ctl.purge([r for r in det])).reshape(-1, 10) if n else np.zeros((0, 10))
        assert got.shape == ref.shape and np.array_equal(got, ref)


def test_label_strings():
    import pytest
    from pyfaceanalysis_b200.attributes import map_real_gender_labels_to_strings, map_real_race_labels_to_strings
    assert map_real_gender_labels_to_strings([-1.0, 0.0, 0.3, 1.0]) == ["Male", "Male", "Female", "Female"]
    assert map_real_gender_labels_to_strings([-0.5, 0.5], long_text=False) == ["M", "F"]
    assert map_real_race_labels_to_strings([-2.0, 0.0, 1.7]) == ["Black", "Black", "White"]
    with pytest.raises(Exception, match="Unrecognized label"):
        map_real_gender_labels_to_strings([1.5])
    with pytest.raises(Exception, match="Unrecognized label"):
        map_real_race_labels_to_strings([10.0])


def test_clock_sampler_window(tmp_path):
    """bench.py's nvidia-smi parser: only samples inside the timed region count; none inside -> samples under load."""
    import datetime
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    t = 1_800_000_000.0

    def line(dt, mhz, cap):
        ts = datetime.datetime.fromtimestamp(t + dt).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        return "%s, %d, 1965, 700.0, 0x4, Not Active, Not Active, Not Active, %s\n" % (ts, mhz, cap)

    class Dead(object):
        def terminate(self):
            pass

        def wait(self, timeout=None):
            pass

    for lines, expect in (([line(-1.0, 500, "Not Active"), line(0.1, 1900, "Active"), line(0.3, 1950, "Active"), line(2.0, 600, "Not Active")],
                           (1925.0, 2, ["sw_power_cap"], "timed region")),
                          ([line(-0.5, 1930, "Not Active"), line(-0.3, 300, "Not Active")], (1930.0, 1, [], None))):
        path = tmp_path / "smi.csv"
        path.write_text("".join(lines))
        s = bench.ClockSampler(0)
        s.proc, s.path = Dead(), str(path)
        out = s.stop(t, t + 0.5)
        assert out["sm_mhz"] == expect[0] and out["samples"] == expect[1] and out["reasons"] == expect[2]
        assert out["sm_max_mhz"] == 1965.0 and (expect[3] is None or out["window"] == expect[3])


def test_group_confidences_matches_the_reference_loop():
    """Eye stage bookkeeping (FaceDetectUpdated.py:1011-1017, 1036-1041): within a (image, scale) group the survivors take
    the confidences of the group's FIRST faces, whichever faces were dropped."""
    from pyfaceanalysis_b200.cascade import group_confidences
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 7, 64, 500):
        cf = rng.random(n)
        ok = rng.random(n) > 0.3
        im = np.sort(rng.integers(0, 5, n))
        sc = np.zeros(n, dtype=np.int64)
        for k in np.unique(im):                     # scales ascending inside an image, as the pyramid enumerates them
            sc[im == k] = np.sort(rng.integers(0, 4, int((im == k).sum())))
        grp = np.stack([im, sc], axis=1)
        want = np.empty(int(ok.sum()))
        pos = start = 0
        while start < n:                            # the reference's per-group slices, written as a loop
            stop = start
            while stop < n and (grp[stop] == grp[start]).all():
                stop += 1
            k_ok = int(ok[start:stop].sum())
            want[pos:pos + k_ok] = cf[start:start + k_ok]
            pos += k_ok
            start = stop
        got = group_confidences(cf, ok, grp)
        assert got.shape == want.shape and np.array_equal(got, want)
