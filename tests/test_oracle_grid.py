"""Window-grid enumeration: counts and bit patterns of the reference arithmetic (face_analysis.py:575-669)."""
import numpy as np

from oracle import grid as ogrid


def test_counts_match_survey(grid_golden):
    # SURVEY.md 8d / BASELINE.md section 3
    assert sum(grid_golden["tns_group_0.1"]["counts"]) == 1308
    assert grid_golden["tns_group_0.1"]["counts"] == [560, 315, 192, 108, 63, 35, 20, 9, 4, 2]
    assert sum(grid_golden["fhd_0.05_prescaled"]["counts"]) == 7452
    assert sum(grid_golden["fhd_0.05"]["counts"]) == 7395
    assert sum(grid_golden["uhd_0.02_prescaled"]["counts"]) == 21182
    assert sum(grid_golden["uhd_0.02"]["counts"]) == 48089


def test_oracle_reproduces_golden_bits(grid_golden, pipeline):
    for name, g in grid_golden.items():
        wins = ogrid.enumerate_windows(g["width"], g["height"], pipeline["net"], g["smallest_face"])
        assert [len(c) for _, c, _ in wins] == g["counts"], name
        assert [float(s).hex() for s, _, _ in wins] == g["sampling_values"], name
        assert [float(v).hex() for v in wins[0][1][1]] == g["first_box"]
        assert [float(v).hex() for v in wins[-1][1][-1]] == g["last_box"]


def test_grid_layout_y_outer_x_inner(pipeline):
    s, coords, geo = ogrid.enumerate_windows(1000, 750, pipeline["net"], 0.1)[0]
    nx, ny = geo["n_x"], geo["n_y"]
    assert len(coords) == nx * ny
    assert coords[0, 0] == 0.0 and coords[0, 1] == 0.0
    assert np.all(coords[:nx, 1] == coords[0, 1])               # first row: constant y
    assert np.all(np.diff(coords[:nx, 0]) > 0)
    assert coords[nx, 0] == coords[0, 0] and coords[nx, 1] > coords[0, 1]
    pw = 64 * s
    assert np.allclose(coords[:, 2] - coords[:, 0], pw - 1) and np.allclose(coords[:, 3] - coords[:, 1], pw - 1)
    assert coords[nx - 1, 0] == 1000 - pw                        # linspace forces the last element to `stop`


def test_prescale(pipeline):
    assert ogrid.prescaled_size(3648, 2736)[:2] == (1000, 750)
    assert ogrid.prescaled_size(1920, 1080)[:2] == (1000, 562)
    assert ogrid.prescaled_size(800, 600) == (800, 600, 1.0)
