"""Pins the oracle's EXTENT resampler against Pillow (live and the committed golden vectors)."""
import numpy as np
import pytest
from PIL import Image

from oracle import crop as ocrop


def test_golden_vectors(crop_golden):
    img, boxes = crop_golden["image"], crop_golden["boxes"]
    for k, b in enumerate(boxes):
        assert np.array_equal(ocrop.extent_nearest(img, b), crop_golden["nearest"][k]), k
        assert np.array_equal(ocrop.extent_bilinear(img, b), crop_golden["bilinear"][k]), k


def test_live_pillow_random_and_adversarial():
    rng = np.random.default_rng(7)
    H, W = 211, 333
    img = rng.integers(0, 256, (H, W), dtype=np.uint8)
    pim = Image.fromarray(img, "L")
    boxes = []
    for _ in range(150):
        s = rng.uniform(8, 400)
        x0, y0 = rng.uniform(-80, W), rng.uniform(-80, H)
        boxes.append((x0, y0, x0 + s - 1, y0 + s * rng.uniform(0.6, 1.4) - 1))
    for x0 in (0.9, 0.7, 1.3, 2.1, 0.3, 10.9):        # accumulate form != multiply form here
        for a in (0.2, 0.6, 1.4, 0.3, 2.2, 0.7):
            boxes.append((x0, x0, x0 + 64 * a, x0 + 64 * a))
    for b in boxes:
        assert np.array_equal(ocrop.extent_nearest(img, b),
                              np.asarray(pim.transform((64, 64), Image.EXTENT, b, Image.NEAREST)))
        assert np.array_equal(ocrop.extent_bilinear(img, b),
                              np.asarray(pim.transform((64, 64), Image.EXTENT, b, Image.BILINEAR)))
        assert np.array_equal(ocrop.extent_bicubic(img, b),
                              np.asarray(pim.transform((64, 64), Image.EXTENT, b, Image.BICUBIC)))


def test_multiply_form_would_be_wrong():
    """The reason the kernel carries the sequential double accumulation."""
    diff = 0
    for x0 in (0.9, 0.7, 1.3, 2.1, 0.3, 10.9):
        for a in (0.2, 0.6, 1.4, 0.3, 2.2, 0.7):
            acc = ocrop.nearest_index_table(x0, x0 + 64 * a, 64, 10 ** 6)
            mul = np.array([int(x0 + a * (c + 0.5)) for c in range(64)])
            diff += int((acc != mul).any())
    assert diff > 0


def test_rotation_zero_angle_consistency():
    rng = np.random.default_rng(8)
    img = rng.integers(0, 256, (120, 160), dtype=np.uint8)
    box = (10.3, 7.9, 90.1, 88.2)
    # a rotation by exactly 0 degrees of the generic path samples the multiply-form positions
    r = ocrop.extent_rotated(img, box, 0.0)
    assert r.shape == (64, 64)
    full = ocrop.extract_subimages(img, [box, box], np.array([0.0, 12.5]))
    assert full.shape == (2, 4096) and full.dtype == np.float64
    assert np.array_equal(full[0].reshape(64, 64), ocrop.extent_nearest(img, box))
    assert not np.array_equal(full[0], full[1])
