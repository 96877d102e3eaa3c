"""oracle/regions.py against the fixture of the unmodified reference's analyze_regions, and the ordering
rule of OpenCV's component numbers (which the CUDA path reproduces) against cv2 itself."""
from pathlib import Path

import numpy as np
import pytest

from oracle import regions as oreg

from regions_util import CASES, check_against_golden

GOLD = Path(__file__).resolve().parent / "golden" / "reference_regions.npz"


@pytest.mark.parametrize("name,conn", CASES)
def test_oracle_matches_reference_fixture(name, conn):
	g = np.load(GOLD)
	check_against_golden(oreg.analyze_regions(g[f"in_{name}"], 100, conn), g, f"{name}__c{conn}")


@pytest.mark.parametrize("conn", [4, 8])
@pytest.mark.parametrize("h,w,p", [(7, 9, 0.5), (64, 64, 0.3), (65, 63, 0.55), (301, 517, 0.45), (1024, 1024, 0.5)])
def test_opencv_component_order_rule(conn, h, w, p):
	cv = pytest.importorskip("cv2")
	rng = np.random.default_rng(h * w + conn)
	m = (rng.random((h, w)) < p).astype(np.uint8) * 255
	n, lab, stats, _ = cv.connectedComponentsWithStats(m, connectivity=conn)
	keys = oreg.component_order_keys(lab, n, conn)
	assert np.all(np.diff(keys) > 0)


def test_degenerate():
	img = np.zeros((5, 6, 4), np.uint8)
	r = oreg.analyze_regions(img)
	assert r["total_regions"] == 0 and r["all_regions"] == []
	with pytest.raises(ValueError, match="rgba must be HxWx4 uint8"):
		oreg.analyze_regions(img[:, :, :3])
