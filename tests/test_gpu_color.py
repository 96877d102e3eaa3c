"""K1 / K9 / K4 parity: colour conversions over ALL 2^24 colours and the fused nearest-centre
remap against the oracle."""
import numpy as np
import pytest

from oracle import hsv as ohsv
from oracle import kmeans as okm
from oracle import lab as olab
from oracle import pipeline as op

from gpu_util import blobby_rgba, engine, to_dev

pytestmark = pytest.mark.gpu


def all_colours_rgba():
	k = np.arange(1 << 24, dtype=np.uint32)
	return np.stack([(k >> 16) & 0xFF, (k >> 8) & 0xFF, k & 0xFF, np.full(1 << 24, 255)], axis=1).astype(np.uint8)


def test_lab_all_colours_within_1e4():
	"""fp32 LAB planes vs the fp64 oracle: abs(err) <= 1e-4 * max(1, |ref|) (SURVEY §8c vii) — the
	kernel evaluates in fp64 and rounds once, so the error is the fp32 rounding (~4e-6)."""
	e = engine()
	px = all_colours_rgba()
	planes = e.rgba_to_lab(to_dev(px)).cpu().numpy()[:, :1 << 24]
	step = 1 << 21
	worst = 0.0
	for s in range(0, 1 << 24, step):
		ref = olab.rgb2lab(px[s:s + step, :3])
		got = planes[:, s:s + step].T.astype(np.float64)
		err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
		worst = max(worst, float(err.max()))
		assert np.array_equal(got.astype(np.float32), ref.astype(np.float32)) or err.max() < 1e-5
	assert worst <= 1e-4


def test_lab_f64_rows_bit_close():
	e = engine()
	rng = np.random.default_rng(0)
	px = rng.integers(0, 256, (100001, 4), dtype=np.uint8)
	got = e.rgba_to_lab_f64(to_dev(px)).cpu().numpy()
	ref = olab.rgb2lab(px[:, :3])
	assert np.abs(got - ref).max() <= 1e-11  # cbrt / FMA-free fp64 evaluation: a few ulp


def test_lab_ragged_sizes():
	e = engine()
	rng = np.random.default_rng(1)
	for n in (1, 2, 3, 5, 1023, 4097):
		px = rng.integers(0, 256, (n, 4), dtype=np.uint8)
		got = e.rgba_to_lab(to_dev(px)).cpu().numpy()[:, :n].T
		ref = olab.rgb2lab(px[:, :3])
		assert np.abs(got - ref).max() < 1e-4


def test_hsv_all_colours_bit_exact():
	e = engine()
	px = all_colours_rgba()
	px[::7, 3] = 13  # alpha is carried through
	got = e.rgba_to_hsv(to_dev(px)).cpu().numpy()
	assert np.array_equal(got[:, 3], px[:, 3])
	step = 1 << 22
	for s in range(0, 1 << 24, step):
		assert np.array_equal(got[s:s + step, :3], ohsv.rgb_to_hsv_u8(px[s:s + step, :3]))


@pytest.mark.parametrize("space", ["rgb", "hsv", "lab"])
@pytest.mark.parametrize("preserve_alpha", [True, False])
def test_assign_remap_matches_oracle(space, preserve_alpha):
	from image_segmenter_b200 import _ffi

	e = engine()
	img = blobby_rgba(5, 203, 157)
	rng = np.random.default_rng(2)
	pal = rng.integers(0, 256, (23, 3), dtype=np.uint8)
	pal[11] = pal[4]  # duplicate entry: lowest index wins
	img[50:60, 50:60, :3] = pal[11]
	if space == "lab":
		feats, sp = olab.rgb2lab(pal), _ffi.CS_SPACE_LAB
	elif space == "hsv":
		feats, sp = ohsv.rgb_to_hsv_u8(pal).astype(np.float64), _ffi.CS_SPACE_HSV
	else:
		feats, sp = pal.astype(np.float64), _ffi.CS_SPACE_RGB
	out, lab = e.assign_remap(to_dev(img.reshape(-1, 4)), sp, feats, pal, preserve_alpha, want_labels=True)
	ref_out, ref_idx = op.assign_remap(img, feats, pal, space, preserve_alpha)
	got_idx = lab.cpu().numpy().reshape(img.shape[:2])
	mism = np.nonzero(got_idx != ref_idx)
	if len(mism[0]):  # only fp64 rounding-level ties between the GEMM form and the direct form
		px = img[mism][:, :3]
		f = olab.rgb2lab(px) if space == "lab" else (ohsv.rgb_to_hsv_u8(px) if space == "hsv" else px).astype(np.float64)
		b, s2 = okm.near_tie_gap(f, feats)
		assert (s2 - b <= 1e-9 * np.maximum(1.0, s2)).all()
		assert len(mism[0]) <= 4
	got = out.cpu().numpy().reshape(img.shape)
	same = got_idx == ref_idx
	assert np.array_equal(got[same], ref_out[same])
	assert np.array_equal(got[..., 3], ref_out[..., 3])


def test_perceptual_quirk_rgb_centres_in_lab_space():
	"""simplify_colors_perceptual compares LAB pixels with RGB-valued centres (:540-544): the kernel
	just takes the table it is given."""
	from image_segmenter_b200 import _ffi

	e = engine()
	img = blobby_rgba(6, 96, 80, alpha_holes=False)
	cen = np.array([[200, 30, 40], [20, 180, 90], [90, 90, 200], [128, 128, 128]], dtype=np.uint8)
	out, _ = e.assign_remap(to_dev(img.reshape(-1, 4)), _ffi.CS_SPACE_LAB, cen.astype(np.float64), cen, True)
	ref, _ = op.assign_remap(img, cen.astype(np.float64), cen, "lab", True)
	assert np.array_equal(out.cpu().numpy().reshape(img.shape), ref)
