"""oracle/lab.py pins: textbook CIELAB values, OpenCV's float Lab, the lab2rgb round trip.
scikit-image itself is not installed (parity of the LAB conversion is unpinned against it)."""
import numpy as np
import pytest

from oracle import lab as olab


def test_textbook_values():
	# SURVEY.md §8a-6 probe values (skimage constants, D65 / 2 degree observer)
	rgb = np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 255], [0, 0, 0]], dtype=np.uint8)
	lab = olab.rgb2lab(rgb)
	exp = np.array([[53.2406, 80.0923, 67.2028], [87.7351, -86.1830, 83.1797], [32.2957, 79.1856, -107.8573],
	                [100.0, -0.0025, 0.0047], [0.0, 0.0, 0.0]])
	assert np.abs(lab - exp).max() < 1.5e-3


def test_against_opencv_float_lab():
	cv = pytest.importorskip("cv2")
	rng = np.random.default_rng(1)
	rgb = rng.integers(0, 256, (4096, 1, 3), dtype=np.uint8)
	ref = cv.cvtColor(rgb.astype(np.float32) / 255.0, cv.COLOR_RGB2Lab).reshape(-1, 3)
	got = olab.rgb2lab(rgb).reshape(-1, 3)
	assert np.abs(got - ref).max() < 0.5  # sanity only: OpenCV 4.13 interpolates a coarse fp32 LUT here


def test_roundtrip_and_lut():
	rng = np.random.default_rng(2)
	rgb = rng.integers(0, 256, (5000, 3), dtype=np.uint8)
	back = olab.lab2rgb(olab.rgb2lab(rgb))
	assert np.abs(back * 255 - rgb).max() < 1e-9
	lut = olab.linear_lut256()
	assert lut.shape == (256,) and lut[0] == 0.0 and abs(lut[255] - 1.0) < 1e-15 and np.all(np.diff(lut) > 0)
	# the table is exactly the per-element curve
	v = olab.srgb_linearize(olab.srgb_u8_to_unit(rgb))
	assert np.array_equal(v, lut[rgb])


def test_product_host_helpers_agree_with_oracle():
	"""image_segmenter_b200._colorspace (palette-sized host helpers of the product) == oracle."""
	from image_segmenter_b200 import _colorspace as cs

	rng = np.random.default_rng(3)
	rgb = rng.integers(0, 256, (3000, 3), dtype=np.uint8)
	assert np.array_equal(cs.linear_lut256(), olab.linear_lut256())
	a, b = cs.rgb2lab_small(rgb), olab.rgb2lab(rgb)
	assert np.abs(a - b).max() < 1e-12
	lab = b + rng.normal(0, 3, b.shape)
	assert np.abs(cs.lab2rgb_small(lab) - olab.lab2rgb(lab)).max() < 1e-12
