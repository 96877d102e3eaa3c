"""The C-ABI library: it builds, loads, and exports every symbol include/colorsimplify.h declares;
argument errors and the no-device error are reported, not swallowed.  No device compute here —
only the host-side entry point (cs_median_cut_boxes) is executed, against the oracle and Pillow."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
	from image_segmenter_b200 import _ffi, build

	build.build()
	return _ffi.load_library()


def header_symbols():
	txt = (ROOT / "include" / "colorsimplify.h").read_text()
	txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
	return sorted(set(re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound(lib):
	from image_segmenter_b200 import _ffi

	syms = header_symbols()
	assert len(syms) >= 35
	for s in syms:
		assert hasattr(lib, s), f"{s} declared in the header but not exported"
	assert sorted(_ffi.SIGNATURES) == syms, "ctypes SIGNATURES and the header disagree"
	assert lib.cs_abi_version() == 1


def test_argument_counts_match_header():
	from image_segmenter_b200 import _ffi

	txt = (ROOT / "include" / "colorsimplify.h").read_text()
	txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
	for name, args in re.findall(r"\b(cs_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt):
		n = 0 if args.strip() in ("", "void") else len(args.split(","))
		assert len(_ffi.SIGNATURES[name]) == n, name


def test_no_device_is_loud(lib):
	import torch

	if torch.cuda.is_available():
		pytest.skip("a GPU is present")
	h = C.c_void_p()
	rc = lib.cs_ctx_create(0, C.byref(h))
	assert rc != 0 and not h.value
	assert b"no CUDA device" in lib.cs_last_error() or b"failed" in lib.cs_last_error()
	from image_segmenter_b200 import _ffi, color_simplify as cs

	img = np.zeros((4, 4, 4), np.uint8)
	img[..., 3] = 255
	for fn in (cs.simplify_colors_kmeans, cs.simplify_colors_median_cut, cs.simplify_colors_threshold,
	           cs.get_color_statistics, cs.simplify_colors_perceptual_fast):
		with pytest.raises(_ffi.ColorSimplifyError):
			fn(img)


def test_value_errors_match_reference_messages():
	from image_segmenter_b200 import color_simplify as cs

	bad = [np.zeros((4, 4, 3), np.uint8), np.zeros((4, 4, 4), np.float32), np.zeros((4, 4), np.uint8)]
	fns = [cs.simplify_colors_kmeans, cs.simplify_colors_median_cut, cs.simplify_colors_octree,
	       cs.simplify_colors_threshold, cs.get_color_statistics, cs.simplify_colors_perceptual,
	       cs.simplify_colors_perceptual_fast, cs.simplify_colors_adaptive_distance, cs.simplify_colors_hsv_clustering]
	for fn in fns:
		for b in bad:
			with pytest.raises(ValueError, match="rgba must be HxWx4 uint8"):
				fn(b)
	ok = np.zeros((2, 2, 4), np.uint8)
	with pytest.raises(ValueError, match="custom_palette must be Nx3 uint8"):
		cs.simplify_colors_custom_palette(ok, np.zeros((3, 3), np.int64))
	with pytest.raises(ValueError, match="Custom palette requires palette parameter"):
		cs.simplify_colors_adaptive(ok, 8, True, "custom_palette")


def _cells(rgb):
	from oracle import mediancut as omc

	shift, cells, counts, keys = omc.histogram_cells(rgb)
	return shift, cells, counts.astype(np.uint32), keys.astype(np.uint32)


@pytest.mark.parametrize("seed,k", [(s, k) for s in range(5) for k in (2, 7, 16, 100, 256)])
def test_host_median_cut_boxes_matches_oracle(lib, seed, k):
	from oracle import mediancut as omc

	rng = np.random.default_rng(seed)
	if seed % 2:
		rgb = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
	else:
		cols = rng.integers(0, 256, (int(rng.integers(3, 14)), 3), dtype=np.uint8)
		rgb = np.repeat(cols, 6, axis=0).reshape(-1, 6, 3)  # equal populations: heap order matters
	shift, cells, counts, keys = _cells(rgb)
	exp = omc.median_cut_boxes(cells, counts.astype(np.int64), k)
	box = np.zeros(len(keys), np.uint16)
	nb = C.c_int(0)
	rc = lib.cs_median_cut_boxes(keys.ctypes.data, counts.ctypes.data, len(keys), shift, k, box.ctypes.data, C.byref(nb))
	assert rc == 0
	assert nb.value == int(exp.max()) + 1
	assert np.array_equal(box.astype(np.int64), exp)


def test_host_median_cut_bad_args(lib):
	nb = C.c_int(0)
	assert lib.cs_median_cut_boxes(None, None, 0, 0, 8, None, C.byref(nb)) != 0
	assert lib.cs_last_error()


def test_header_constants_match_ffi():
	"""Every integer / float #define of the header that the Python binding mirrors has the same value."""
	from image_segmenter_b200 import _ffi

	txt = (ROOT / "include" / "colorsimplify.h").read_text()
	defs = dict(re.findall(r"^#define\s+(CS_[A-Z0-9_]+)\s+(-?[0-9][0-9.eE+-]*)\s*$", txt, flags=re.M))
	assert {"CS_LLOYD_EXACT_TIES", "CS_LLOYD_CHAINED", "CS_LAB_NORM2_MAX"} <= set(defs)
	mirrored = [k for k in defs if hasattr(_ffi, k)]
	assert "CS_LLOYD_CHAINED" in mirrored and "CS_LLOYD_EXACT_TIES" in mirrored
	for k in mirrored:
		assert float(getattr(_ffi, k)) == float(defs[k]), k
	# flag bits are distinct
	assert _ffi.CS_LLOYD_EXACT_TIES & _ffi.CS_LLOYD_CHAINED == 0
