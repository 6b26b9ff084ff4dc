"""GPU parity of window extraction (csrc/crop.cu): bit-exact against Pillow's own output (golden
fixtures and live) for angle 0, and against the oracle's definition for rotated windows."""
import numpy as np
import pytest

from oracle import crop as ocrop

pytestmark = pytest.mark.gpu


def test_nearest_golden_bit_exact(crop_golden):
    from pyfaceanalysis_b200 import extract_subimages, NEAREST
    img, boxes = crop_golden["image"], crop_golden["boxes"]
    got = extract_subimages(img, boxes, None, (64, 64), NEAREST, np.uint8)
    assert np.array_equal(got.reshape(-1, 64, 64), crop_golden["nearest"])
    got64 = extract_subimages(img, boxes, None, (64, 64), NEAREST, np.float64)
    assert got64.dtype == np.float64 and np.array_equal(got64, got.astype(np.float64))


def test_bilinear_golden_bit_exact(crop_golden):
    from pyfaceanalysis_b200 import extract_subimages, BILINEAR
    img, boxes = crop_golden["image"], crop_golden["boxes"]
    got = extract_subimages(img, boxes, None, (64, 64), BILINEAR, np.uint8)
    assert np.array_equal(got.reshape(-1, 64, 64), crop_golden["bilinear"])


def test_detection_grid_windows_vs_pillow(pipeline):
    """Every window of the reference grid for a 1000x750 image (config 1 geometry), incl. the
    overhanging ones at the largest scales, against Pillow run live."""
    from PIL import Image
    from oracle import grid as ogrid
    from pyfaceanalysis_b200 import extract_subimages, NEAREST
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (750, 1000), dtype=np.uint8)
    pim = Image.fromarray(img, "L")
    coords = np.concatenate([c for _, c, _ in ogrid.enumerate_windows(1000, 750, pipeline["net"], 0.1)])
    assert len(coords) == 1308
    got = extract_subimages(img, coords, np.zeros(len(coords)), (64, 64), NEAREST, np.uint8)
    for k in range(0, len(coords), 7):
        ref = np.asarray(pim.transform((64, 64), Image.EXTENT, tuple(coords[k]), Image.NEAREST))
        assert np.array_equal(got[k].reshape(64, 64), ref), k
    ref_all = ocrop.extract_subimages(img, coords).astype(np.uint8)
    assert np.array_equal(got, ref_all)


def test_bicubic_bit_exact(crop_golden):
    """BICUBIC (listed as an alternative by the reference, FaceDetectUpdated.py:125): device == oracle, which is
    pinned against Pillow 12.2 in tests/test_oracle_crop.py; tiled and row-major outputs, overhanging boxes."""
    from pyfaceanalysis_b200 import extract_subimages, BICUBIC
    img, boxes = crop_golden["image"], crop_golden["boxes"]
    got = extract_subimages(img, boxes, None, (64, 64), BICUBIC, np.uint8)
    ref = ocrop.extract_subimages(img, boxes, None, (64, 64), BICUBIC).astype(np.uint8)
    assert np.array_equal(got, ref)
    rng = np.random.default_rng(8)
    small = rng.integers(0, 256, (37, 41), dtype=np.uint8)
    b = np.array([[-5.5, -3.2, 30.1, 28.9], [10.0, 12.0, 60.0, 50.0], [0.0, 0.0, 41.0, 37.0], [39.5, 35.5, 44.0, 40.0]])
    got = extract_subimages(small, b, None, (32, 24), BICUBIC, np.float64)
    ref = ocrop.extract_subimages(small, b, None, (32, 24), BICUBIC)
    assert np.array_equal(got, ref)


def test_rotated_windows_match_oracle():
    from pyfaceanalysis_b200 import extract_subimages, NEAREST, BILINEAR, BICUBIC
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (300, 400), dtype=np.uint8)
    n = 60
    s = rng.uniform(20, 200, n)
    x0 = rng.uniform(-30, 380, n)
    y0 = rng.uniform(-30, 280, n)
    coords = np.stack([x0, y0, x0 + s - 1, y0 + s - 1], axis=1)
    angles = rng.uniform(-25, 25, n)
    angles[::5] = 0.0
    for interp in (NEAREST, BILINEAR, BICUBIC):
        got = extract_subimages(img, coords, angles, (64, 64), interp, np.uint8)
        ref = ocrop.extract_subimages(img, coords, angles, (64, 64), interp).astype(np.uint8)
        mism = np.mean(got != ref)
        # sin/cos differ by <= 1 ulp between libm and the device: a sample on a pixel boundary may flip
        assert mism <= 1e-4, (interp, mism)


def test_load_network_subimages_signature():
    from pyfaceanalysis_b200 import load_network_subimages
    rng = np.random.default_rng(5)
    imgs = [rng.integers(0, 256, (120, 160), dtype=np.uint8)]
    coords = np.array([[3.2, 4.1, 80.7, 81.6], [-10.0, -5.0, 53.0, 58.0]])
    out = load_network_subimages(imgs, np.zeros(2, dtype=int), coords, np.zeros(2), 64, 64, 0, False)
    assert out.shape == (2, 4096) and out.dtype == np.float64
    assert np.array_equal(out, ocrop.extract_subimages(imgs[0], coords))
    assert load_network_subimages(imgs, np.zeros(0, dtype=int), np.zeros((0, 4)), np.zeros(0), 64, 64, 0, False).shape == (0, 0)
