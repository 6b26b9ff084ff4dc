"""The oracle's age-stage crop against Pillow run live: the three resampling steps of ``normalize_image``
(``face_normalization_tools.py:274-324``) composed with Pillow's own calls must give the oracle's bytes."""
import numpy as np
import pytest

from oracle import normalize as onorm


def _image(seed, H=180, W=240):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    return (128 + 70 * np.sin(xx / 7.0) * np.cos(yy / 11.0) + 35 * rng.standard_normal((H, W))).clip(0, 255).astype(np.uint8)


def _pillow_normalize(img, eyes):
    from PIL import Image
    H, W = img.shape
    g = onorm.geometry(eyes, W, H)
    im = Image.fromarray(img, "L")
    first = im.transform(g["crop_size"], Image.EXTENT, g["crop_coordinates"])
    rotated = first.rotate(g["rotation_angle"], Image.BICUBIC)        # = rotate_improved on a current Pillow (oracle header)
    return np.asarray(rotated.transform(onorm.OUT_SIZE, Image.EXTENT, g["transform_coords"], Image.BICUBIC))


@pytest.mark.parametrize("eyes", [(90.3, 70.2, 131.8, 74.9), (60.0, 100.0, 100.0, 100.0), (150.5, 40.25, 200.75, 22.5),
                                  (10.0, 20.0, 60.0, 35.0), (100.0, 60.0, 124.0, 58.0)])
def test_normalize_image_matches_pillow_composition(eyes):
    img = _image(3)
    got = onorm.normalize_image(img, eyes)
    ref = _pillow_normalize(img, eyes)
    assert got.shape == (260, 256) and got.dtype == np.uint8
    assert np.array_equal(got, ref)


def test_rotate_restatement_and_geometry():
    from PIL import Image
    img = _image(5, 61, 77)
    for ang in (0.0, 13.37, -7.5, 90.0, 180.0, 359.2):
        ref = np.asarray(Image.fromarray(img, "L").rotate(ang, Image.BICUBIC))
        assert np.array_equal(onorm.rotate_bicubic(img, ang), ref), ang
    g = onorm.geometry((90.3, 70.2, 131.8, 74.9), 240, 180)
    cw, ch = g["crop_size"]
    assert cw % 2 == 1 and ch % 2 == 1                                   # "Result is always integer and odd"
    x0, y0, x1, y1 = g["crop_coordinates"]
    assert x0 == int(x0) and y0 == int(y0) and x1 - x0 == cw and y1 - y0 == ch
    assert g["rotation_angle"] == pytest.approx(np.degrees(np.arctan2(74.9 - 70.2, 131.8 - 90.3)))
    with pytest.raises(ValueError):
        onorm.geometry((131.8, 70.2, 90.3, 74.9), 240, 180)              # the reference exits on swapped eyes
    box = onorm.age_box()
    assert box == pytest.approx((36.8, 31.96, 219.2, 214.36))


def test_age_subimages_shape_and_contrast():
    img = _image(7)
    det = np.array([[0, 0, 0, 0, 0, 90.3, 70.2, 131.8, 74.9, 0.1], [0, 0, 0, 0, 0, 60.0, 100.0, 100.0, 100.0, 0.2]])
    p = onorm.age_subimages(img, det)
    assert p.shape == (2, 9216)
    assert np.allclose(p.mean(axis=1), 0.0, atol=1e-9) and np.allclose(p.std(axis=1), 0.16, rtol=1e-6)
