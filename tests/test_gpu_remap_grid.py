"""K4 grid-filtered paths (csrc/lab.cu; RGB / LAB metric, K >= 4): remap_grid_kernel (three-phase tiles, 2 MP
up) and remap_lut_build_kernel + remap_lut_kernel (per-colour table of the mixed cells, 16 MP up; policies 1
and 2 of cs_remap_set_policy force either at any size): the labels and the output image must equal the direct
kernel's — the fp64 first minimum of
pairwise_distances_argmin_min (app/processing/color_simplify.py:543-557, 691-705, 1106-1121) — for every
one of the 2^24 colours."""
import numpy as np
import pytest

from oracle import kmeans as okm
from oracle import lab as olab

from gpu_util import engine, to_dev

pytestmark = pytest.mark.gpu


def _all_colours():
	v = np.arange(1 << 24, dtype=np.uint32)
	img = np.empty((1 << 24, 4), np.uint8)
	img[:, 0], img[:, 1], img[:, 2] = v & 255, (v >> 8) & 255, v >> 16
	img[:, 3] = 255
	return img


def _direct(e, d_img, sp, feats, pal, preserve_alpha):
	"""The direct (per-pixel, all K) kernel: a 4-byte-offset view of a padded copy is not 16-byte aligned,
	which the grid path refuses."""
	import torch

	n = d_img.shape[0]
	buf = torch.empty((n + 1, 4), dtype=torch.uint8, device=e.dev)
	buf[1:] = d_img
	return e.assign_remap(buf[1:], sp, feats, pal, preserve_alpha, want_labels=True)


@pytest.fixture(autouse=True)
def _default_policy():
	yield
	engine().set_remap_policy(0)


@pytest.mark.parametrize("policy", [1, 2])
@pytest.mark.parametrize("space,K,seed", [("lab", 16, 0), ("lab", 5, 1), ("lab", 64, 2), ("lab", 256, 3), ("rgb", 16, 4), ("rgb", 200, 5)])
def test_grid_remap_equals_direct_kernel_on_all_colours(space, K, seed, policy):
	from image_segmenter_b200 import _ffi

	e = engine()
	e.set_remap_policy(policy)
	rng = np.random.default_rng(seed)
	pal = rng.integers(0, 256, (K, 3), dtype=np.uint8)
	if K >= 8:
		pal[K // 2] = pal[1]  # duplicate entry: lowest index wins
		pal[K - 1] = np.clip(pal[2].astype(int) + [1, 0, -1], 0, 255)  # near-duplicate: a thin cell
	feats, sp = (olab.rgb2lab(pal), _ffi.CS_SPACE_LAB) if space == "lab" else (pal.astype(np.float64), _ffi.CS_SPACE_RGB)
	img = _all_colours()
	img[::7, 3] = 0      # transparent pixels keep RGB 0, label 255
	img[3::11, 3] = 100  # preserve_alpha=False: alpha <= 128 -> 0
	d = to_dev(img)
	for pa in (True, False):
		out_g, lab_g = e.assign_remap(d, sp, feats, pal, pa, want_labels=True)
		out_d, lab_d = _direct(e, d, sp, feats, pal, pa)
		assert bool((lab_g == lab_d).all()), f"{int((lab_g != lab_d).sum())} labels differ"
		assert bool((out_g == out_d).all())
	# and against the oracle on a strided sample of the colours (sklearn ArgKmin restated in oracle/kmeans.py)
	sub = np.arange(0, 1 << 24, 37)
	sub = sub[img[sub, 3] > 0]
	f = olab.rgb2lab(img[sub, :3]) if space == "lab" else img[sub, :3].astype(np.float64)
	ref, _ = okm.argmin_min(f, feats)
	got = lab_g.cpu().numpy()[sub]
	mism = np.nonzero(got != ref)[0]
	if len(mism):  # only fp64 rounding-level ties between the GEMM form and the direct form
		b, s2 = okm.near_tie_gap(f[mism], feats)
		assert (s2 - b <= 1e-9 * np.maximum(1.0, s2)).all()


@pytest.mark.parametrize("policy", [1, 2])
def test_grid_remap_ragged_size_and_centres_off_the_palette(policy):
	"""n not a multiple of the tile or of 4; float centres that are not palette colours (perceptual_fast's
	fitted LAB centres).  Policy 1: the three-phase tiles, 2: the colour table."""
	from image_segmenter_b200 import _ffi

	e = engine()
	e.set_remap_policy(policy)
	rng = np.random.default_rng(9)
	n = (1 << 21) + 4099
	img = np.empty((n, 4), np.uint8)
	img[:, :3] = rng.integers(0, 256, (n, 3), dtype=np.uint8)
	img[:, 3] = rng.choice([0, 1, 128, 129, 255], n)
	cen = olab.rgb2lab(rng.integers(0, 256, (16, 3), dtype=np.uint8)) + rng.normal(0, 0.3, (16, 3))
	pal = rng.integers(0, 256, (16, 3), dtype=np.uint8)
	d = to_dev(img)
	out_g, lab_g = e.assign_remap(d, _ffi.CS_SPACE_LAB, cen, pal, True, want_labels=True)
	out_d, lab_d = _direct(e, d, _ffi.CS_SPACE_LAB, cen, pal, True)
	assert bool((lab_g == lab_d).all()) and bool((out_g == out_d).all())


def test_default_policy_by_size_gives_the_same_image():
	"""Policy 0 at 16 MP (colour table) and the direct kernel (policy -1) on an image with few colours and a
	4-colour palette."""
	from image_segmenter_b200 import _ffi

	e = engine()
	rng = np.random.default_rng(21)
	n = (1 << 24) + 3
	cols = rng.integers(0, 256, (300, 4), dtype=np.uint8)
	img = cols[rng.integers(0, 300, n)]
	pal = rng.integers(0, 256, (4, 3), dtype=np.uint8)
	d = to_dev(img)
	out0, lab0 = e.assign_remap(d, _ffi.CS_SPACE_LAB, olab.rgb2lab(pal), pal, False, want_labels=True)
	e.set_remap_policy(-1)
	out1, lab1 = e.assign_remap(d, _ffi.CS_SPACE_LAB, olab.rgb2lab(pal), pal, False, want_labels=True)
	assert bool((lab0 == lab1).all()) and bool((out0 == out1).all())
