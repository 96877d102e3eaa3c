"""sRGB <-> CIELAB in NumPy float64 — restatement of scikit-image's rgb2lab / lab2rgb.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED AGAINST SKIMAGE ITSELF: the
reference calls `skimage.color.rgb2lab` / `lab2rgb` (app/processing/color_simplify.py:470, 540,
658, 681, 688, 757, 1090-1091) but scikit-image (requirements.txt: `scikit-image>=0.22`,
unpinned) is not installed in this image and cannot be fetched.  This file restates the published
algorithm of `skimage/color/colorconv.py` (0.22-0.25: rgb2xyz, xyz2lab, lab2xyz, xyz2rgb with
illuminant D65, observer 2) and is checked against textbook CIELAB values and OpenCV's float
COLOR_RGB2Lab (tests/test_oracle_lab.py).
"""
from __future__ import annotations

import numpy as np

# skimage.color.colorconv.xyz_from_rgb (sRGB primaries, D65)
XYZ_FROM_RGB = np.array([[0.412453, 0.357580, 0.180423],
                         [0.212671, 0.715160, 0.072169],
                         [0.019334, 0.119193, 0.950227]], dtype=np.float64)
RGB_FROM_XYZ = np.linalg.inv(XYZ_FROM_RGB)
# skimage xyz_tristimulus_values(illuminant="D65", observer="2")
WHITE_D65 = np.array([0.95047, 1.0, 1.08883], dtype=np.float64)


def srgb_u8_to_unit(u8: np.ndarray) -> np.ndarray:
	"""skimage img_as_float on uint8: multiply by 1/255 in float64 (util/dtype.py `_convert`)."""
	return np.multiply(u8, 1.0 / 255.0, dtype=np.float64)


def srgb_linearize(v: np.ndarray) -> np.ndarray:
	"""rgb2xyz's companding step: v > 0.04045 ? ((v+0.055)/1.055)**2.4 : v/12.92."""
	v = np.array(v, dtype=np.float64, copy=True)
	mask = v > 0.04045
	v[mask] = np.power((v[mask] + 0.055) / 1.055, 2.4)
	v[~mask] /= 12.92
	return v


def linear_lut256() -> np.ndarray:
	"""Linearised value of every uint8 level.  The companding curve is a pure function of one
	byte, so this 256-entry table is exact; the product builds the same table on the host and
	hands it to the K1/K4 kernels."""
	return srgb_linearize(srgb_u8_to_unit(np.arange(256, dtype=np.uint8)))


def rgb2lab(rgb: np.ndarray) -> np.ndarray:
	"""rgb: (..., 3) uint8 or float in [0,1] -> (..., 3) float64 L,a,b."""
	rgb = np.asarray(rgb)
	arr = srgb_u8_to_unit(rgb) if rgb.dtype == np.uint8 else rgb.astype(np.float64)
	arr = srgb_linearize(arr)
	xyz = arr @ XYZ_FROM_RGB.T
	t = xyz / WHITE_D65
	mask = t > 0.008856
	f = np.empty_like(t)
	f[mask] = np.cbrt(t[mask])
	f[~mask] = 7.787 * t[~mask] + 16.0 / 116.0
	fx, fy, fz = f[..., 0], f[..., 1], f[..., 2]
	L = 116.0 * fy - 16.0
	a = 500.0 * (fx - fy)
	b = 200.0 * (fy - fz)
	return np.stack([L, a, b], axis=-1)


def lab2rgb(lab: np.ndarray) -> np.ndarray:
	"""lab: (..., 3) float -> (..., 3) float64 sRGB in [0,1] (clipped), as skimage lab2rgb."""
	lab = np.asarray(lab, dtype=np.float64)
	L, a, b = lab[..., 0], lab[..., 1], lab[..., 2]
	fy = (L + 16.0) / 116.0
	fx = a / 500.0 + fy
	fz = fy - b / 200.0
	fz = np.where(fz < 0, 0.0, fz)  # skimage clamps negative z (with a warning)
	f = np.stack([fx, fy, fz], axis=-1)
	mask = f > 0.2068966
	out = np.empty_like(f)
	out[mask] = np.power(f[mask], 3.0)
	out[~mask] = (f[~mask] - 16.0 / 116.0) / 7.787
	out *= WHITE_D65
	arr = out @ RGB_FROM_XYZ.T
	mask = arr > 0.0031308
	arr[mask] = 1.055 * np.power(arr[mask], 1 / 2.4) - 0.055
	arr[~mask] *= 12.92
	np.clip(arr, 0, 1, out=arr)
	return arr


def install_skimage_stub() -> None:
	"""Make `from skimage import color` resolve to this module so the UNMODIFIED reference
	functions that need skimage can run in the authoring container (oracle/make_golden.py)."""
	import sys
	import types

	if "skimage" in sys.modules and not getattr(sys.modules["skimage"], "_oracle_stub", False):
		return  # a real scikit-image is present: use it
	pkg = types.ModuleType("skimage")
	pkg._oracle_stub = True
	col = types.ModuleType("skimage.color")
	col.rgb2lab = rgb2lab
	col.lab2rgb = lab2rgb
	pkg.color = col
	sys.modules["skimage"] = pkg
	sys.modules["skimage.color"] = col
