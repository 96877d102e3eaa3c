"""Host-side colour-space helpers for PALETTE-SIZED arrays (<= a few thousand colours).

The per-pixel conversions run in CUDA (csrc/lab.cu); what stays on the host is what the reference
also does on a handful of colours: LAB of the <= 10 000 sampled unique colours before the palette
fit, LAB -> RGB of the K fitted centres, HSV / LAB of a user palette
(app/processing/color_simplify.py:470, 658, 681, 1091, 1098).  Float64 NumPy, the published
algorithm of skimage.color.rgb2lab / lab2rgb (illuminant D65, observer 2) and of OpenCV's 8-bit
RGB2HSV.  `linear_lut256()` is also the table handed to the K1/K4 kernels.
"""
from __future__ import annotations

import numpy as np

_M = np.array([[0.412453, 0.357580, 0.180423],
               [0.212671, 0.715160, 0.072169],
               [0.019334, 0.119193, 0.950227]], dtype=np.float64)
_M_INV = np.linalg.inv(_M)
_WHITE = np.array([0.95047, 1.0, 1.08883], dtype=np.float64)


def _linearize(v):
	v = np.array(v, dtype=np.float64, copy=True)
	hi = v > 0.04045
	v[hi] = np.power((v[hi] + 0.055) / 1.055, 2.4)
	v[~hi] /= 12.92
	return v


def linear_lut256() -> np.ndarray:
	"""sRGB-linearised value of each uint8 level (img_as_float multiplies by 1/255)."""
	return _linearize(np.multiply(np.arange(256, dtype=np.uint8), 1.0 / 255.0, dtype=np.float64))


def rgb2lab_small(rgb_u8: np.ndarray) -> np.ndarray:
	"""(N,3) uint8 -> (N,3) float64 CIELAB."""
	lin = linear_lut256()[np.asarray(rgb_u8, dtype=np.uint8)]
	t = (lin @ _M.T) / _WHITE
	hi = t > 0.008856
	f = np.where(hi, np.cbrt(np.where(hi, t, 1.0)), 7.787 * t + 16.0 / 116.0)
	return np.stack([116.0 * f[..., 1] - 16.0, 500.0 * (f[..., 0] - f[..., 1]), 200.0 * (f[..., 1] - f[..., 2])], axis=-1)


def lab2rgb_small(lab: np.ndarray) -> np.ndarray:
	"""(N,3) float CIELAB -> (N,3) float64 sRGB in [0,1], clipped."""
	lab = np.asarray(lab, dtype=np.float64)
	fy = (lab[..., 0] + 16.0) / 116.0
	fx = lab[..., 1] / 500.0 + fy
	fz = np.maximum(fy - lab[..., 2] / 200.0, 0.0)
	f = np.stack([fx, fy, fz], axis=-1)
	hi = f > 0.2068966
	xyz = np.where(hi, np.power(f, 3.0), (f - 16.0 / 116.0) / 7.787) * _WHITE
	v = xyz @ _M_INV.T
	hi = v > 0.0031308
	v = np.where(hi, 1.055 * np.power(np.where(hi, v, 1.0), 1 / 2.4) - 0.055, v * 12.92)
	return np.clip(v, 0, 1)


def rgb2hsv_u8_small(rgb_u8: np.ndarray) -> np.ndarray:
	"""(N,3) uint8 -> (N,3) uint8 HSV with OpenCV's 8-bit fixed-point arithmetic (H in [0,179])."""
	i = np.arange(1, 256, dtype=np.float64)
	sdiv = np.zeros(256, dtype=np.int64)
	hdiv = np.zeros(256, dtype=np.int64)
	sdiv[1:] = np.rint((255 << 12) / i)
	hdiv[1:] = np.rint((180 << 12) / (6.0 * i))
	c = np.asarray(rgb_u8, dtype=np.uint8).astype(np.int64)
	r, g, b = c[..., 0], c[..., 1], c[..., 2]
	v = c.max(axis=-1)
	diff = v - c.min(axis=-1)
	s = (diff * sdiv[v] + 2048) >> 12
	h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
	h = (h * hdiv[diff] + 2048) >> 12
	h = np.where(h < 0, h + 180, h)
	return np.clip(np.stack([h, s, v], axis=-1), 0, 255).astype(np.uint8)


def hsv_feature_luts() -> np.ndarray:
	"""3 x 256 fp32 tables of the weighted HSV features of simplify_colors_hsv_clustering
	(color_simplify.py:969-981): float32(h/179)*2.0, float32(s/255)*1.5, float32(v/255)*1.0 —
	evaluated in float64 as the reference does, then rounded to the fp32 the kernel computes in."""
	i = np.arange(256, dtype=np.uint8)
	h = (i / 179.0).astype(np.float32) * 2.0
	s = (i / 255.0).astype(np.float32) * 1.5
	v = (i / 255.0).astype(np.float32) * 1.0
	return np.stack([h, s, v]).astype(np.float32)
