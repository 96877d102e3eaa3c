"""Grid-filtered exact assignment (csrc/lloyd.cu: grid_build_kernel + the GRID Lloyd kernel; enabled by
cs_lloyd_set_feature_box for CS_LLOYD_EXACT_TIES launches with 4 <= K <= 64 on >= 2^18 pixels): labels must
equal the fp64 first minimum (sklearn/cluster/_k_means_lloyd.pyx:205-213) exactly as the full walk's do —
for any centres and for any input, including pixels outside the box the caller vouched for."""
import numpy as np
import pytest

from oracle import kmeans as okm

from gpu_util import engine, lab_like, lloyd_step, planes_of
from test_gpu_lloyd import _check_step

pytestmark = pytest.mark.gpu

N = (1 << 18) + 4103  # eligible (>= 2^18) and ragged: a partial tile and a < 4-pixel tail


def _lab_box():
	from image_segmenter_b200 import _ffi

	return _ffi.CS_LAB_BOX


def _same_as_full_walk(X32, C, box):
	p = planes_of(X32)
	g = lloyd_step(p, len(X32), C, exact=True, fused=True, box=box)
	f = lloyd_step(p, len(X32), C, exact=True, fused=True, box=None)
	assert np.array_equal(g["labels"], f["labels"])
	assert np.array_equal(g["counts"], f["counts"])
	assert (g["guard"] == 77).all()
	assert np.allclose(g["sums"], f["sums"], rtol=2e-6, atol=1e-2)  # fp32 slots: the pixel -> tile split differs
	return g


@pytest.mark.parametrize("K", [4, 5, 8, 13, 16, 17, 32, 47, 64])
def test_grid_labels_equal_oracle_uniform_lab(K):
	rng = np.random.default_rng(K)
	X32 = lab_like(rng, N)
	C = X32[rng.choice(N, K, replace=False)].astype(np.float64)
	_check_step(X32, C, exact=True, fused=True, box=_lab_box())
	_same_as_full_walk(X32, C, _lab_box())


def test_grid_on_real_lab_of_srgb():
	from oracle import lab as olab

	rng = np.random.default_rng(3)
	X32 = olab.rgb2lab(rng.integers(0, 256, (N, 3), dtype=np.uint8)).astype(np.float32)
	C = X32[rng.choice(N, 16, replace=False)].astype(np.float64)
	_check_step(X32, C, exact=True, fused=True, box=_lab_box())
	_same_as_full_walk(X32, C, _lab_box())


def test_grid_crowded_centres_overflow_pool_and_fp64_cells():
	"""16 centres inside a radius-3 ball: the cells around it list more than four (pool) or more than eight
	(fp64 evaluation) candidates; most pixels sit right there."""
	rng = np.random.default_rng(11)
	C = np.array([50.0, 0.0, 0.0]) + rng.normal(0, 1.5, (16, 3))
	X32 = (np.array([50.0, 0.0, 0.0]) + rng.normal(0, 6.0, (N, 3))).astype(np.float32)
	X32[: N // 8] = lab_like(rng, N // 8)
	_check_step(X32, C, exact=True, fused=True, box=_lab_box())
	_same_as_full_walk(X32, C, _lab_box())


def test_grid_crowded_centres_k64():
	rng = np.random.default_rng(12)
	C = np.vstack([np.array([50.0, 0.0, 0.0]) + rng.normal(0, 2.0, (40, 3)), lab_like(rng, 24).astype(np.float64)])
	X32 = (np.array([50.0, 0.0, 0.0]) + rng.normal(0, 8.0, (N, 3))).astype(np.float32)
	X32[: N // 4] = lab_like(rng, N // 4)
	_check_step(X32, C, exact=True, fused=True, box=_lab_box())
	_same_as_full_walk(X32, C, _lab_box())


def test_grid_box_too_small_is_only_slower():
	"""Pixels outside the box fall into border cells that extend to infinity: still exact."""
	rng = np.random.default_rng(5)
	X32 = lab_like(rng, N)
	X32[:1000] *= 40.0  # far outside any LAB box
	C = X32[rng.choice(N, 16, replace=False)].astype(np.float64)
	box = ((40.0, -10.0, -10.0), (60.0, 10.0, 10.0))
	x2 = float((X32.astype(np.float64) ** 2).sum(1).max()) * 1.001
	p = planes_of(X32)
	g = lloyd_step(p, N, C, exact=True, fused=True, box=box, x2max=x2)
	ref = okm.assign_labels(X32.astype(np.float64), C)
	mism = np.nonzero(g["labels"] != ref)[0]
	if len(mism):
		best, second = okm.near_tie_gap(X32[mism].astype(np.float64), C)
		assert (second - best <= 1e-9 * np.maximum(1.0, best)).all()


@pytest.mark.parametrize("box", [((0.0, 0.0, -108.5), (100.5, 0.0, 95.0)), ((0.0, -87.0, -108.5), (0.0, 99.0, 95.0)),
                                 ((5.0, 5.0, 5.0), (5.0, 5.0, 5.0))])
def test_grid_degenerate_box(box):
	rng = np.random.default_rng(7)
	X32 = lab_like(rng, N)
	C = X32[rng.choice(N, 9, replace=False)].astype(np.float64)
	_same_as_full_walk(X32, C, box)


def test_grid_duplicate_centres_and_exact_ties():
	"""First minimum on exact ties (duplicate centres; lattice points midway between two centres)."""
	C = np.array([[10.0, 0, 0], [20.0, 0, 0], [10.0, 0, 0], [30.0, 5, 5], [20.0, 0, 0], [90.0, 50, -50]])
	base = np.array([[10, 0, 0], [15, 0, 0], [20, 0, 0], [25, 2.5, 2.5], [60, 27.5, -22.5]], dtype=np.float32)
	X32 = np.tile(base, (N // 5 + 1, 1))[:N]
	g = _same_as_full_walk(X32, C, _lab_box())
	assert np.array_equal(g["labels"][:5], np.array([0, 0, 1, 1, 3], np.uint8))


def test_grid_fit_trajectory_matches_full_walk_and_oracle():
	from image_segmenter_b200.engine import KMeansGPU

	rng = np.random.default_rng(21)
	X32 = lab_like(rng, N)
	C0 = X32[rng.choice(N, 16, replace=False)].astype(np.float64)
	e = engine()
	p = planes_of(X32)
	from image_segmenter_b200 import _ffi

	fg = KMeansGPU(e, "f32", N, planes=p, box=_ffi.CS_LAB_BOX, grid_policy=1).fit_single(C0, max_iter=7, tol=-1.0)
	ff = KMeansGPU(e, "f32", N, planes=p, box=None).fit_single(C0, max_iter=7, tol=-1.0)
	assert fg.n_iter == ff.n_iter == 7
	assert np.allclose(fg.centers, ff.centers, rtol=1e-6, atol=1e-6)
	lab_o, _, cen_o, _ = okm.kmeans_single_lloyd(X32.astype(np.float64), C0, max_iter=7, tol=-1.0)
	assert np.allclose(fg.centers, cen_o, rtol=1e-4, atol=1e-5)
	assert (fg.labels[:N].cpu().numpy() != lab_o.astype(np.uint8)).mean() < 1e-4  # centres differ at 1e-7: near ties may flip


def test_default_policy_grid_from_ten_megapixels_matches_walk():
	"""Policy 0 (what KMeansGPU / make_gpu_lloyd run): at K = 16 the library takes the grid path from 10^7 pixels
	(csrc/lloyd.cu grid_eligible).  10.5 MP of real sRGB -> CIELAB pixels: six iterations through the default
	path give the labels of six iterations through the full walk, byte for byte."""
	import torch

	from image_segmenter_b200 import _ffi
	from image_segmenter_b200.sharded import make_gpu_lloyd

	e = engine()
	n = 10_500_003
	g = torch.Generator(device=e.dev)
	g.manual_seed(11)
	rgba = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device=e.dev, generator=g)
	planes = e.rgba_to_lab(rgba)
	del rgba
	idx = torch.from_numpy(np.random.default_rng(2).choice(n, 16, replace=False)).to(e.dev)
	C0 = np.ascontiguousarray(planes[:, idx].T.double().cpu().numpy())
	got = {}
	for mode, box in (("default", _ffi.CS_LAB_BOX), ("walk", None)):
		lab = torch.full((n + 4,), 77, dtype=torch.uint8, device=e.dev)
		drv = make_gpu_lloyd(e, planes, n, 16, labels=lab, exact=True, box=box)
		drv.set_centers(C0)
		for _ in range(6):
			drv.iterate()
		torch.cuda.synchronize()
		got[mode] = (lab.cpu().numpy(), drv.c[drv.cur].cpu().numpy())
	assert np.array_equal(got["default"][0][:n], got["walk"][0][:n])
	assert (got["default"][0][n:] == 77).all() and (got["walk"][0][n:] == 77).all()
	assert np.allclose(got["default"][1], got["walk"][1], rtol=0, atol=1e-5)
