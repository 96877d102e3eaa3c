"""cs_host_upload (csrc/hostio.cu): the threaded staged upload of pageable caller memory — the ordinary NumPy
array the reference's colour panel passes (app/ui/main_window.py:596-601) — must deliver every byte, for sizes
around the 8 MB staging buffers and the 512 KB work items, back to back (ring reuse), and through the public
entry points."""
import numpy as np
import pytest

from gpu_util import engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nbytes", [1, 4099, (512 << 10) + 1, 8 << 20, (8 << 20) + 4, (40 << 20) + 12345, 96 << 20])
def test_staged_upload_delivers_every_byte(nbytes):
	import torch

	e = engine()
	rng = np.random.default_rng(nbytes % 1000)
	for rep in range(2):  # the second call reuses staging buffers whose DMAs may still be in flight
		src = rng.integers(0, 256, nbytes, dtype=np.uint8)
		d = torch.zeros(nbytes + 16, dtype=torch.uint8, device=e.dev)
		e._call("cs_host_upload", src.ctypes.data, nbytes, d.data_ptr())
		src_copy = src.copy()
		src[:] = 0  # the call has consumed the source: overwriting it must not change what arrives
		got = d.cpu().numpy()
		assert (got[:nbytes] == src_copy).all()
		assert (got[nbytes:] == 0).all()  # nothing written past the end


def test_entry_point_takes_the_staged_path_for_a_pageable_image():
	from image_segmenter_b200 import color_simplify as cs

	e = engine()
	rng = np.random.default_rng(5)
	img = rng.integers(0, 256, (2048, 2048, 4), dtype=np.uint8)  # 16 MB, pageable
	img[..., 3] = 255
	assert img.nbytes >= e.STAGED_UPLOAD_MIN_BYTES
	d = e.upload_rgba(img)
	assert bool((d.cpu().numpy() == img.reshape(-1, 4)).all())
	out_a, pal_a = cs.simplify_colors_threshold(img, 8)
	pinned = e.to_host(d).reshape(img.shape)  # the same pixels in page-locked memory: the direct path
	out_b, pal_b = cs.simplify_colors_threshold(pinned, 8)
	assert (out_a == out_b).all() and (np.asarray(pal_a) == np.asarray(pal_b)).all()
