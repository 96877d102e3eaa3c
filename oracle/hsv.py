"""RGB -> HSV on uint8 exactly as OpenCV's 8-bit path — restatement of cv2.cvtColor(COLOR_RGB2HSV).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference calls cv2.cvtColor on uint8 pixels
at app/processing/color_simplify.py:947 and :1097-1098; OpenCV (requirements.txt:
`opencv-python>=4.8`, unpinned; image has opencv-python-headless 4.13.0) implements it in
imgproc/src/color_hsv.simd.hpp (RGB2HSV_b) with 12-bit fixed-point reciprocal tables.  Pinned by
tests/test_oracle_hsv.py against cv2 itself over all 2^24 colours.
"""
from __future__ import annotations

import numpy as np

HSV_SHIFT = 12


def _tables():
	i = np.arange(1, 256, dtype=np.float64)
	sdiv = np.zeros(256, dtype=np.int64)
	hdiv = np.zeros(256, dtype=np.int64)
	# saturate_cast<int>(double) == cvRound: round half to even, which np.rint also does
	sdiv[1:] = np.rint((255 << HSV_SHIFT) / i).astype(np.int64)
	hdiv[1:] = np.rint((180 << HSV_SHIFT) / (6.0 * i)).astype(np.int64)
	return sdiv, hdiv


SDIV, HDIV180 = _tables()


def rgb_to_hsv_u8(rgb: np.ndarray) -> np.ndarray:
	"""rgb: (..., 3) uint8 -> (..., 3) uint8 with H in [0,179], S,V in [0,255]."""
	rgb = np.asarray(rgb, dtype=np.uint8)
	r = rgb[..., 0].astype(np.int64)
	g = rgb[..., 1].astype(np.int64)
	b = rgb[..., 2].astype(np.int64)
	v = np.maximum(np.maximum(r, g), b)
	vmin = np.minimum(np.minimum(r, g), b)
	diff = v - vmin
	s = (diff * SDIV[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
	# hue numerator: v==r -> g-b ; else v==g -> b-r+2diff ; else r-g+4diff (first match wins)
	h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
	h = (h * HDIV180[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT  # arithmetic shift, as in C
	h = np.where(h < 0, h + 180, h)
	out = np.stack([h, s, v], axis=-1)
	return np.clip(out, 0, 255).astype(np.uint8)


def hsv_weighted_features(hsv_u8: np.ndarray) -> np.ndarray:
	"""Feature vector of simplify_colors_hsv_clustering (color_simplify.py:969-981):
	float32(h/179, s/255, v/255) * float64[2.0, 1.5, 1.0]  (result float64)."""
	hsv_u8 = np.asarray(hsv_u8)
	f = hsv_u8.copy().astype(np.float32)
	f[:, 0] = hsv_u8[:, 0] / 179.0
	f[:, 1:] = hsv_u8[:, 1:] / 255.0
	return f * np.array([2.0, 1.5, 1.0])
