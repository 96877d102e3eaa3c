"""oracle/ — CPU restatement of the reference's colour-simplification hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `image_segmenter_b200/` imports this package; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs do,
and there only as the checker or as the timed CPU baseline — never as the product path.

Every function cites the reference file:line it follows (relative to the reference checkout
`app/processing/color_simplify.py`, or `sklearn/…`, `PIL/…`, `cv2` for the third-party wheel
the reference calls into).

How the oracle is pinned: the reference ships no tests, golden vectors or
fixtures for this path, so the pins are
  * outputs of the UNMODIFIED reference module imported from /root/reference in the authoring
    container (`oracle/make_golden.py`, fixtures committed under `tests/golden/`), and
  * the third-party routines the reference itself calls — scikit-learn 1.9.0, OpenCV 4.13.0,
    Pillow 12.2.0 — which are installed in the image (also on the GPU box) and are compared
    with each restatement directly in `tests/test_oracle_*.py`.
scikit-image (rgb2lab / lab2rgb) is NOT installed anywhere: `oracle/lab.py` restates its
published algorithm and is validated against textbook CIELAB values and OpenCV's float Lab
only — parity for the LAB conversion is therefore "unpinned against skimage itself".
"""
