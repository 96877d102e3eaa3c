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


# The two helpers below keep skimage's array SHAPES and operation order (the reference calls
# rgb2lab / lab2rgb on (N,1,3) arrays, color_simplify.py:470, 658, 681): the K fitted centres are
# truncated to uint8 afterwards, and a centre that is an exact colour comes back as 110.99999999999999
# or 111.00000000000001 depending on how the 3x3 product is rounded — so the product of the host
# helper has to round like the library's.


def rgb2lab_small(rgb_u8: np.ndarray) -> np.ndarray:
	"""(N,3) uint8 -> (N,3) float64 CIELAB (rgb2xyz -> xyz2lab, D65 / 2 degree observer)."""
	arr = linear_lut256()[np.asarray(rgb_u8, dtype=np.uint8).reshape(-1, 1, 3)]
	xyz = arr @ _M.T
	t = xyz / _WHITE
	hi = t > 0.008856
	f = np.empty_like(t)
	f[hi] = np.cbrt(t[hi])
	f[~hi] = 7.787 * t[~hi] + 16.0 / 116.0
	fx, fy, fz = f[..., 0], f[..., 1], f[..., 2]
	return np.stack([116.0 * fy - 16.0, 500.0 * (fx - fy), 200.0 * (fy - fz)], axis=-1).reshape(-1, 3)


def lab2rgb_small(lab: np.ndarray) -> np.ndarray:
	"""(N,3) float CIELAB -> (N,3) float64 sRGB in [0,1], clipped (lab2xyz -> xyz2rgb)."""
	lab = np.asarray(lab, dtype=np.float64).reshape(-1, 1, 3)
	L, a, b = lab[..., 0], lab[..., 1], lab[..., 2]
	fy = (L + 16.0) / 116.0
	fx = a / 500.0 + fy
	fz = fy - b / 200.0
	fz = np.where(fz < 0, 0.0, fz)
	f = np.stack([fx, fy, fz], axis=-1)
	hi = f > 0.2068966
	out = np.empty_like(f)
	out[hi] = np.power(f[hi], 3.0)
	out[~hi] = (f[~hi] - 16.0 / 116.0) / 7.787
	out *= _WHITE
	v = out @ _M_INV.T
	hi = v > 0.0031308
	v[hi] = 1.055 * np.power(v[hi], 1 / 2.4) - 0.055
	v[~hi] *= 12.92
	np.clip(v, 0, 1, out=v)
	return v.reshape(-1, 3)


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


def hsv_feature_luts64() -> np.ndarray:
	"""The same three tables exactly as the reference's float64 feature values
	(float32(h/179), float32(s/255), float32(v/255) widened to float64, times [2.0, 1.5, 1.0]) — used where
	reference precision matters (k-means++ seeding, initial centres)."""
	i = np.arange(256, dtype=np.uint8)
	h = (i / 179.0).astype(np.float32).astype(np.float64) * 2.0
	s = (i / 255.0).astype(np.float32).astype(np.float64) * 1.5
	v = (i / 255.0).astype(np.float32).astype(np.float64) * 1.0
	return np.stack([h, s, v])
