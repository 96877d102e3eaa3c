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


def test_against_independent_high_precision_evaluation():
	"""The published rgb2lab formulas (same constants: sRGB companding, the 0.412453... matrix, D65 / 2 degree
	white, 0.008856 / 7.787 branch) evaluated pixel by pixel with 50-digit `decimal` arithmetic: an independent
	code path (no NumPy, no shared helper) that the vectorised float64 oracle must match to rounding level —
	a transcription or operation-order slip in oracle/lab.py would show up here."""
	from decimal import Decimal as D, getcontext

	getcontext().prec = 50
	M = [[D("0.412453"), D("0.357580"), D("0.180423")], [D("0.212671"), D("0.715160"), D("0.072169")],
	     [D("0.019334"), D("0.119193"), D("0.950227")]]
	W = [D("0.95047"), D(1), D("1.08883")]

	def ln(x):
		return x.ln()

	def powd(x, e):
		return (ln(x) * e).exp()

	def lab_of(rgb):
		lin = []
		for c in rgb:
			v = D(int(c)) / D(255)
			lin.append(powd((v + D("0.055")) / D("1.055"), D("2.4")) if v > D("0.04045") else v / D("12.92"))
		f = []
		for i in range(3):
			t = sum(M[i][j] * lin[j] for j in range(3)) / W[i]
			f.append(powd(t, D(1) / D(3)) if t > D("0.008856") else D("7.787") * t + D(16) / D(116))
		return [float(D(116) * f[1] - D(16)), float(D(500) * (f[0] - f[1])), float(D(200) * (f[1] - f[2]))]

	rng = np.random.default_rng(5)
	rgb = np.vstack([rng.integers(0, 256, (300, 3)), [[0, 0, 0], [255, 255, 255], [1, 1, 1], [10, 10, 10], [11, 11, 11],
	                 [255, 0, 0], [0, 255, 0], [0, 0, 255], [2, 0, 1], [128, 128, 128]]]).astype(np.uint8)
	ref = np.array([lab_of(p) for p in rgb])
	got = olab.rgb2lab(rgb)
	# float64 evaluation: |error| a few ulp of the largest intermediate (500 * f ~ 500)
	assert np.abs(got - ref).max() < 5e-12
