"""image_segmenter_b200 — B200-native colour-simplification hot path.

Drop-in for `app/processing/color_simplify.py` of jeffreyperez1620/image_segmenter: same entry
points and arguments (`image_segmenter_b200.color_simplify`), every per-pixel step in
hand-written sm_100a CUDA kernels reached through the C ABI in `include/colorsimplify.h`.
There is no CPU fallback: importing the package is cheap, but any compute call raises
`RuntimeError` when the CUDA library or a B200 device is missing.
"""
__version__ = "0.1.0"
