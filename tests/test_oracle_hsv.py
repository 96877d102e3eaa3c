"""oracle/hsv.py pinned against cv2.cvtColor(COLOR_RGB2HSV) over ALL 2^24 colours (the routine the
reference calls at color_simplify.py:947, 1097-1098)."""
import numpy as np
import pytest

from oracle import hsv as ohsv


def test_all_colours_match_opencv():
	cv = pytest.importorskip("cv2")
	for r0 in range(0, 256, 32):
		r, g, b = np.meshgrid(np.arange(r0, r0 + 32, dtype=np.uint8), np.arange(256, dtype=np.uint8),
		                      np.arange(256, dtype=np.uint8), indexing="ij")
		rgb = np.stack([r, g, b], axis=-1).reshape(-1, 1, 3)
		ref = cv.cvtColor(rgb, cv.COLOR_RGB2HSV).reshape(-1, 3)
		got = ohsv.rgb_to_hsv_u8(rgb.reshape(-1, 3))
		assert np.array_equal(ref, got)


def test_weighted_features_and_product_luts():
	from image_segmenter_b200 import _colorspace as cs

	hsv = np.stack([np.arange(180, dtype=np.uint8), np.arange(180, dtype=np.uint8) + 70,
	                np.arange(180, dtype=np.uint8) + 40], axis=1)
	f = ohsv.hsv_weighted_features(hsv)
	assert f.dtype == np.float64
	lut = cs.hsv_feature_luts()
	via_lut = np.stack([lut[c][hsv[:, c]] for c in range(3)], axis=1)
	# the kernel computes in fp32: the table is the fp32 rounding of the reference's fp64 feature
	assert np.abs(via_lut.astype(np.float64) - f).max() <= 2.0 ** -24 * 2.0
	assert np.array_equal(cs.rgb2hsv_u8_small(hsv), ohsv.rgb_to_hsv_u8(hsv))
