"""oracle/pipeline.py on degenerate and ragged inputs (1x1, 3x5, mixed alpha, fully transparent, all dark,
only the fallback brightness threshold passes, two colours) against outputs of the UNMODIFIED reference
(tests/golden/reference_edge_cases.npz, made by `python -m oracle.make_golden edge`)."""
import warnings
from pathlib import Path

import numpy as np
import pytest

from oracle import pipeline as op

G = np.load(Path(__file__).resolve().parent / "golden" / "reference_edge_cases.npz", allow_pickle=False)
IMAGES = ("tiny", "onepx", "ragged", "transparent", "dark", "midbright", "twocolors")
CP = G["custom_palette_in"]

CALLS = {
	"kmeans_8": lambda im: op.kmeans_rgb(im, 8, intended_remap=False),
	"median_cut_8": lambda im: op.median_cut(im, 8),
	"octree_5": lambda im: op.median_cut(im, 5, power_of_two=False),
	"threshold_8": lambda im: op.threshold(im, 8),
	"threshold_8_noalpha": lambda im: op.threshold(im, 8, preserve_alpha=False),
	"hsv_4": lambda im: op.hsv_clustering(im, 4),
	"custom_rgb": lambda im: op.custom_palette(im, CP, True, "rgb"),
	"custom_lab": lambda im: op.custom_palette(im, CP, True, "lab"),
	"custom_hsv_noalpha": lambda im: op.custom_palette(im, CP, False, "hsv"),
	"perceptual_fast_4": lambda im: op.perceptual_fast(im, 4),
	"perceptual_3": lambda im: op.perceptual(im, 3, max_samples=2000),
}


@pytest.mark.parametrize("name", IMAGES)
@pytest.mark.parametrize("tag", sorted(CALLS))
def test_edge_case_matches_reference(name, tag):
	img = G[f"in_{name}"]
	key = f"{name}__{tag}"
	assert f"{key}__raises" not in G.files  # the reference raised nowhere on these inputs
	with warnings.catch_warnings():
		warnings.simplefilter("ignore")
		np.random.seed(7)
		out, pal = CALLS[tag](img)
	assert np.array_equal(out, G[f"{key}__rgba"]), key
	ref_pal = G[f"{key}__palette"]
	assert np.array_equal(np.asarray(pal), ref_pal), key
	assert np.asarray(pal).dtype == ref_pal.dtype, key
	# degenerate early-outs hand back the INPUT object
	assert (out is img) == bool(G[f"{key}__same_object"][0]), key


@pytest.mark.parametrize("name", IMAGES)
def test_edge_case_statistics(name):
	st = op.statistics(G[f"in_{name}"])
	ref = G[f"{name}__stats"]
	assert st["total_unique_colors"] == int(ref[0]) and st["non_transparent_pixels"] == int(ref[1])
	assert np.allclose(st["rgb_mean"], ref[2:5], rtol=1e-12, atol=1e-12)
	assert np.allclose(st["rgb_std"], ref[5:8], rtol=1e-12, atol=1e-12)
