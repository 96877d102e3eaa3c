"""Comparison of an analyze_regions result with the fixture of the unmodified reference."""
import numpy as np

DIST_KEYS = ("< 50", "50-99", "100-199", "200-499", "500+")


def check_against_golden(r, g, tag):
	assert [r["total_regions"], r["small_regions"], r["largest_region_size"], r["smallest_region_size"]] == list(g[f"{tag}__summary"])
	assert np.array_equal(np.array(r["region_colors"], dtype=np.uint8).reshape(-1, 3), g[f"{tag}__colors"])
	assert np.array_equal(np.array(r["region_sizes"], dtype=np.int64), g[f"{tag}__sizes"])
	assert np.array_equal(np.array([a["label"] for a in r["all_regions"]]), g[f"{tag}__labels"])
	assert np.array_equal(np.array([a["component_id"] for a in r["all_regions"]]), g[f"{tag}__labels"])
	assert np.array_equal(np.array([a["bbox"] for a in r["all_regions"]], dtype=np.int64).reshape(-1, 4), g[f"{tag}__bbox"])
	assert [r["size_distribution"].get(k, 0) for k in DIST_KEYS] == list(g[f"{tag}__dist"])
	ucol = np.unique(g[f"{tag}__colors"], axis=0)
	for j, c in enumerate(ucol):
		a = next(a for a in r["all_regions"] if tuple(int(v) for v in a["color"]) == tuple(int(v) for v in c))
		assert a["labels"].dtype == np.int32 and a["color_mask"].dtype == np.uint8
		assert np.array_equal(a["labels"], g[f"{tag}__label_images"][j])
		assert np.array_equal(a["color_mask"], g[f"{tag}__mask_images"][j])
	for a in r["all_regions"]:
		assert isinstance(a["size"], int) and isinstance(a["label"], int) and isinstance(a["color"], tuple)


CASES = [(n, c) for n in ("blobby_t8", "blobby_mc8", "fewcolors", "uniform_t2") for c in (8, 4)]
