"""oracle/mediancut.py pinned against Pillow's own Image.quantize(method=MEDIANCUT) — the call the
reference makes at color_simplify.py:145 and :201.  Palettes and index maps must be bit-identical."""
import numpy as np
import pytest

from oracle import mediancut as omc

Image = pytest.importorskip("PIL.Image")


def pillow_quantize(rgb, k):
	im = Image.fromarray(rgb).quantize(colors=k, method=Image.Quantize.MEDIANCUT)
	pal = np.array(im.getpalette()).reshape(-1, 3)
	idx = np.array(im)
	return pal, idx


def blobby(seed, h=60, w=70, ncol=7, sigma=10):
	rng = np.random.default_rng(seed)
	cent = rng.integers(0, 256, (ncol, 3))
	which = rng.integers(0, ncol, (h, w))
	return np.clip(cent[which] + rng.normal(0, sigma, (h, w, 3)), 0, 255).astype(np.uint8)


CASES = [(s, k) for s in range(6) for k in (2, 5, 16, 64, 256)]


@pytest.mark.parametrize("seed,k", CASES)
def test_generic_images(seed, k):
	rgb = blobby(seed) if seed % 2 == 0 else np.random.default_rng(seed).integers(0, 256, (48, 52, 3), dtype=np.uint8)
	pal_ref, idx_ref = pillow_quantize(rgb, k)
	pal, idx = omc.quantize(rgb, k)
	assert np.array_equal(pal, pal_ref[:len(pal)])
	assert np.array_equal(idx, idx_ref)


@pytest.mark.parametrize("seed", range(12))
def test_equal_population_heap_order(seed):
	"""Boxes with equal pixel counts: the pop order is decided by Pillow's array heap."""
	rng = np.random.default_rng(100 + seed)
	ncol = int(rng.integers(3, 12))
	cols = rng.integers(0, 256, (ncol, 3), dtype=np.uint8)
	rgb = np.repeat(cols, 8, axis=0).reshape(ncol, 8, 3)
	k = int(rng.integers(2, ncol + 2))
	pal_ref, idx_ref = pillow_quantize(rgb, k)
	pal, idx = omc.quantize(rgb, k)
	assert np.array_equal(pal, pal_ref[:len(pal)])
	assert np.array_equal(idx, idx_ref)


def test_large_scale_shift():
	"""> 65536 distinct colours forces a cell shift (create_pixel_hash rescales)."""
	rng = np.random.default_rng(7)
	rgb = rng.integers(0, 256, (320, 320, 3), dtype=np.uint8)
	shift, cells, counts, keys = omc.histogram_cells(rgb)
	assert shift >= 1 and len(cells) <= 65536 and counts.sum() == 320 * 320
	pal_ref, idx_ref = pillow_quantize(rgb, 32)
	pal, idx = omc.quantize(rgb, 32)
	assert np.array_equal(pal, pal_ref[:len(pal)])
	assert np.array_equal(idx, idx_ref)


def test_tie_heavy_map():
	"""Palette entries equidistant from many pixels: Pillow's scan-order tie rule."""
	g = np.arange(0, 256, 5, dtype=np.uint8)
	rgb = np.stack(np.meshgrid(g, g[:20], indexing="ij"), axis=-1)
	rgb = np.concatenate([rgb, rgb[..., :1]], axis=-1).astype(np.uint8)
	for k in (4, 8):
		pal_ref, idx_ref = pillow_quantize(rgb, k)
		pal, idx = omc.quantize(rgb, k)
		assert np.array_equal(pal, pal_ref[:len(pal)])
		assert np.array_equal(idx, idx_ref)
