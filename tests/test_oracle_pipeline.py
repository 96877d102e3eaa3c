"""oracle/pipeline.py (whole entry points) against the fixtures produced by the UNMODIFIED
reference module (tests/golden/reference_entry_points.npz, made by oracle/make_golden.py)."""
import warnings

import numpy as np
import pytest

from oracle import pipeline as op

IMAGES = ("blobby", "uniform", "fewcolors")


def _eq(g, tag, out, pal):
	assert np.array_equal(out, g[f"{tag}__rgba"]), tag
	ref_pal = g[f"{tag}__palette"]
	assert np.array_equal(np.asarray(pal), ref_pal), tag
	assert np.asarray(pal).dtype == ref_pal.dtype, tag


@pytest.mark.parametrize("name", IMAGES)
def test_integer_paths_bit_exact(golden, name):
	img = golden[f"in_{name}"]
	for k in (8, 16, 100):
		_eq(golden, f"{name}__median_cut_{k}", *op.median_cut(img, k))
	for k in (6, 16, 256):
		_eq(golden, f"{name}__octree_{k}", *op.median_cut(img, k, power_of_two=False))
	for k in (2, 8, 16, 256):
		_eq(golden, f"{name}__threshold_{k}", *op.threshold(img, k))
	_eq(golden, f"{name}__threshold_8_noalpha", *op.threshold(img, 8, preserve_alpha=False))


@pytest.mark.parametrize("name", IMAGES)
def test_statistics(golden, name):
	st = op.statistics(golden[f"in_{name}"])
	ref = golden[f"{name}__stats"]
	assert st["total_unique_colors"] == int(ref[0]) and st["non_transparent_pixels"] == int(ref[1])
	assert np.allclose(st["rgb_mean"], ref[2:5], rtol=1e-12) and np.allclose(st["rgb_std"], ref[5:8], rtol=1e-12)


@pytest.mark.parametrize("name", IMAGES)
def test_kmeans_reference_quirk_and_palette(golden, name):
	img = golden[f"in_{name}"]
	for k in (5, 16):
		with warnings.catch_warnings():
			warnings.simplefilter("ignore")
			out, pal = op.kmeans_rgb(img, k, intended_remap=False)
		_eq(golden, f"{name}__kmeans_{k}", out, pal)
		assert not out[:, :, :3].any()  # the reference's remap is a no-op (color_simplify.py:90)


@pytest.mark.parametrize("name", IMAGES)
def test_custom_palette(golden, name):
	img, cp = golden[f"in_{name}"], golden["custom_palette_in"]
	for metric in ("rgb", "hsv", "lab"):
		_eq(golden, f"{name}__custom_{metric}", *op.custom_palette(img, cp, True, metric))
	_eq(golden, f"{name}__custom_lab_noalpha", *op.custom_palette(img, cp, False, "lab"))


@pytest.mark.parametrize("name", IMAGES)
def test_perceptual_and_hsv(golden, name):
	img = golden[f"in_{name}"]
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		np.random.seed(7)
		_eq(golden, f"{name}__perceptual_fast_6", *op.perceptual_fast(img, 6))
		np.random.seed(7)
		_eq(golden, f"{name}__perceptual_5", *op.perceptual(img, 5, max_samples=2000))
		_eq(golden, f"{name}__hsv_6", *op.hsv_clustering(img, 6))
