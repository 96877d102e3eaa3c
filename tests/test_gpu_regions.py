"""analyze_regions on the device (one labelling for all colours) against the unmodified reference's
fixture and, on larger images, against the oracle (cv2 per colour, as the reference does)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import regions as oreg

from gpu_util import blobby_rgba
from regions_util import CASES, check_against_golden

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "reference_regions.npz"


@pytest.fixture(scope="module")
def rc():
	from image_segmenter_b200 import region_cleanup

	return region_cleanup


@pytest.mark.parametrize("name,conn", CASES)
def test_matches_reference_fixture(rc, name, conn):
	g = np.load(GOLD)
	check_against_golden(rc.analyze_regions(g[f"in_{name}"], 100, conn), g, f"{name}__c{conn}")


def _same(a, b):
	assert a["total_regions"] == b["total_regions"] and a["small_regions"] == b["small_regions"]
	assert a["largest_region_size"] == b["largest_region_size"] and a["smallest_region_size"] == b["smallest_region_size"]
	assert a["size_distribution"] == b["size_distribution"]
	assert a["region_sizes"] == b["region_sizes"]
	assert [tuple(int(v) for v in c) for c in a["region_colors"]] == [tuple(int(v) for v in c) for c in b["region_colors"]]
	for x, y in zip(a["all_regions"], b["all_regions"]):
		assert x["label"] == y["label"] and tuple(int(v) for v in x["bbox"]) == tuple(int(v) for v in y["bbox"])
	seen = set()
	for x, y in zip(a["all_regions"], b["all_regions"]):
		if x["color"] in seen:
			continue
		seen.add(x["color"])
		assert np.array_equal(x["labels"], y["labels"]) and np.array_equal(x["color_mask"], y["color_mask"])


@pytest.mark.parametrize("conn", [8, 4])
@pytest.mark.parametrize("shape,step", [((257, 331), 128), ((1080, 1920), 85), ((33, 2049), 128)])
def test_larger_images_vs_oracle(rc, shape, step, conn):
	img = blobby_rgba(41, *shape, sigma=14.0)
	img[:, :, :3] = (img[:, :, :3] // step) * step  # a posterised (colour-simplified) image with noisy borders
	_same(rc.analyze_regions(img, 100, conn), oreg.analyze_regions(img, 100, conn))


def test_random_noise_many_components(rc):
	rng = np.random.default_rng(3)
	img = np.zeros((300, 400, 4), np.uint8)
	img[:, :, :3] = rng.integers(0, 2, (300, 400, 3)) * 255  # 8 colours, salt-and-pepper
	img[:, :, 3] = (rng.random((300, 400)) < 0.9) * 255
	for conn in (8, 4):
		_same(rc.analyze_regions(img, 3, conn), oreg.analyze_regions(img, 3, conn))


def test_degenerate_and_errors(rc):
	img = np.zeros((5, 6, 4), np.uint8)
	r = rc.analyze_regions(img)
	assert r["total_regions"] == 0 and r["all_regions"] == [] and r["size_distribution"] == {}
	with pytest.raises(ValueError, match="rgba must be HxWx4 uint8"):
		rc.analyze_regions(img[:, :, :3])
	one = np.full((4, 4, 4), 255, np.uint8)
	r = rc.analyze_regions(one)
	assert r["total_regions"] == 1 and r["region_sizes"] == [16] and tuple(int(v) for v in r["all_regions"][0]["bbox"]) == (0, 0, 4, 4)
