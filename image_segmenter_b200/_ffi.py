"""ctypes binding of libcolorsimplify.so (the C ABI in include/colorsimplify.h).

PyTorch tensors are only the buffer carrier: every call passes `tensor.data_ptr()` and the
current CUDA stream handle.  There is no fallback — a missing library or device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "_lib" / "libcolorsimplify.so"

CS_LLOYD_EXACT_TIES = 1
CS_LLOYD_CHAINED = 2
CS_SPACE_RGB, CS_SPACE_LAB, CS_SPACE_HSV = 0, 1, 2
CS_MAX_K = 256
CS_LAB_NORM2_MAX = 31400.0
# box of cs_rgba8_to_lab's output over all 2^24 sRGB colours, with slack (include/colorsimplify.h CS_LAB_BOX_*)
CS_LAB_BOX = ((0.0, -87.0, -108.5), (100.5, 99.0, 95.0))

_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64

# name -> argtypes (restype is always int unless noted); mirrors include/colorsimplify.h
SIGNATURES = {
	"cs_abi_version": [],
	"cs_last_error": [],
	"cs_ctx_create": [_i, C.POINTER(_vp)],
	"cs_ctx_destroy": [_vp],
	"cs_ctx_sm_count": [_vp],
	"cs_debug_scratch": [_vp, _vp],
	"cs_rgba8_to_lab": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp],
	"cs_rgba8_to_lab_f64": [_vp, _vp, _i64, _vp, _vp, _vp],
	"cs_lloyd_step_f32": [_vp, _vp, _vp, _vp, _i64, _vp, _i, _vp, _vp, _vp, _vp, C.c_double, _i, _vp],
	"cs_feature_norm2_max_f32": [_vp, _vp, _vp, _vp, _i64, _vp, _vp],
	"cs_lloyd_step_rgba8": [_vp, _vp, _i64, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp],
	"cs_lloyd_set_feature_box": [_vp, _vp, _vp],
	"cs_lloyd_set_grid_policy": [_vp, _i],
	"cs_lloyd_finalize": [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp],
	"cs_lloyd_iter_f32": [_vp, _vp, _vp, _vp, _i64, _vp, _i, _vp, _vp, _vp, _vp, _vp, C.c_double, _i, _vp],
	"cs_lloyd_run_f32": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i, _vp, _vp, _vp, C.c_double, _i, _i, _vp, _vp],
	"cs_lloyd_run_px8": [_vp, _vp, _i64, _vp, _i, _i, C.c_double, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp],
	"cs_mg_create": [_vp, _i, _i, _vp],
	"cs_mg_connect": [_vp, _vp],
	"cs_mg_error": [_vp, _vp],
	"cs_mg_destroy": [_vp],
	"cs_lloyd_iter_f32_mg": [_vp, _vp, _vp, _vp, _i64, _vp, _i, _vp, _vp, _vp, _vp, _vp, C.c_double, _i, _vp],
	"cs_lloyd_relocate_f32": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i, _vp, _vp, _vp],
	"cs_lloyd_farthest_f32": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i, C.c_uint64, _vp, _vp, _vp],
	"cs_lloyd_relocate_px8": [_vp, _vp, _i64, _vp, _vp, _vp, _i, _vp, _vp, _vp],
	"cs_lloyd_iter_rgba8": [_vp, _vp, _i64, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp],
	"cs_lloyd_iter_rgba8_batched": [_vp, _vp, _i64, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp],
	"cs_lloyd_step_px8lut": [_vp, _vp, _i64, _vp, _i, _i, C.c_double, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp],
	"cs_kpp_eval_batched": [_vp, _vp, _i64, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, C.POINTER(_i), _vp],
	"cs_kpp_update_batched": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp],
	"cs_nn_argmin_rows64": [_vp, _vp, _i64, _vp, _i64, _vp, _vp],
	"cs_kmeans_fit_rows64_small": [_vp, _vp, _i64, _vp, _i, _i, _i, C.c_double, _vp, _vp, _vp, _vp],
	"cs_lloyd_step_rows64": [_vp, _vp, _i64, _vp, _i, _vp, _vp, _vp],
	"cs_sum_by_label_rows64": [_vp, _vp, _i64, _vp, _i, _vp, _vp, _vp],
	"cs_kpp_locate_batched": [_vp, _vp, _i64, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp],
	"cs_kpp_pick_batched": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp],
	"cs_kpp_draw_batched": [_vp, _vp, _i64, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp],
	"cs_kpp_record_batched": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp],
	"cs_sum_by_label_rgba8": [_vp, _vp, _vp, _i64, _vp, _i, _i, _i, _vp, _vp],
	"cs_merge_labels_u8": [_vp, _vp, _vp, _i64, _vp, _i, _i, _vp, _vp],
	"cs_label_cooccurrence_u8": [_vp, _vp, _vp, _i64, _vp, _i, _i, _vp, _vp],
	"cs_select_compact_px8": [_vp, _vp, _i64, _i, _i, _vp, _vp, _i64, _vp, _vp],
	"cs_channel_hist_px8": [_vp, _vp, _i64, _i, _i, _vp, _vp],
	"cs_gather_px8": [_vp, _vp, _i64, _vp, _i64, _vp, _vp],
	"cs_assign_remap_rgba8": [_vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp],
	"cs_remap_set_policy": [_vp, _i],
	"cs_remap_labels_rgba8": [_vp, _vp, _vp, _i64, _vp, _i, _i, _vp, _i, _i, _vp, _vp],
	"cs_hist_rgb24": [_vp, _vp, _i64, _vp, _vp],
	"cs_hist_fold": [_vp, _vp, _i, _vp, _vp, _vp],
	"cs_hist_compact": [_vp, _vp, _i64, _vp, _vp, C.c_uint32, _vp, _vp],
	"cs_median_cut_boxes": [_vp, _vp, C.c_uint32, _i, _i, _vp, C.POINTER(_i)],
	"cs_box_sums": [_vp, _vp, _vp, _i, _i, _vp, _vp],
	"cs_palette_map_rgba8": [_vp, _vp, _i64, _vp, _i, _vp, _i, _i, _vp, _vp, _vp],
	"cs_posterize_rgba8": [_vp, _vp, _i64, _i, _i, _vp, _vp, _vp],
	"cs_stats_rgba8": [_vp, _vp, _i64, _vp, _vp, _vp],
	"cs_bitmap_popcount": [_vp, _vp, _i64, _vp, _vp],
	"cs_mask_stats_rgba8": [_vp, _vp, _i64, _i, _vp, _vp, _vp],
	"cs_mask_stats_hsv8": [_vp, _vp, _i64, _i, _vp, _vp, _vp],
	"cs_rgba8_to_hsv8": [_vp, _vp, _i64, _vp, _vp],
	"cs_ccl_label": [_vp, _vp, _i, _i, _i, _vp, _vp],
	"cs_ccl_roots": [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp],
	"cs_ccl_stats": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
	"cs_ccl_extract": [_vp, _vp, _vp, _i64, _vp, _vp, _i, _vp, _vp, _vp],
	"cs_host_register": [_vp, C.c_size_t],
	"cs_host_unregister": [_vp],
	"cs_host_upload": [_vp, _vp, C.c_size_t, _vp, _vp],
	"cs_host_lab_kmeans": [_vp, _vp, _i64, _vp, _vp, _i, _i, C.c_double, _i, _vp, C.POINTER(_i),
	                       C.POINTER(C.c_double)],
}

_lib = None


class ColorSimplifyError(RuntimeError):
	pass


def load_library() -> C.CDLL:
	"""dlopen the in-tree library and bind every symbol of the header.  Raises if absent."""
	global _lib
	if _lib is not None:
		return _lib
	path = Path(os.environ.get("COLORSIMPLIFY_LIB", LIB_PATH))
	if not path.exists():
		raise ColorSimplifyError(
			f"{path} is missing: build it with `python -m image_segmenter_b200.build` "
			"(there is no CPU fallback)")
	lib = C.CDLL(str(path))
	for name, argtypes in SIGNATURES.items():
		fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
		fn.argtypes = argtypes
		fn.restype = C.c_char_p if name == "cs_last_error" else C.c_int
	if lib.cs_abi_version() != 1:
		raise ColorSimplifyError("libcolorsimplify ABI version mismatch")
	_lib = lib
	return lib


def check(rc: int, what: str) -> None:
	if rc != 0:
		msg = load_library().cs_last_error()
		raise ColorSimplifyError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


class Context:
	"""Owns one cs_ctx (per-device scratch).  Single-threaded, like the reference's caller."""

	def __init__(self, device: int = 0):
		import torch

		if not torch.cuda.is_available():
			raise ColorSimplifyError(
				"no CUDA device: image_segmenter_b200 runs on B200 (sm_100a) only and has no CPU fallback")
		self.lib = load_library()
		self.device = int(device)
		h = _vp()
		check(self.lib.cs_ctx_create(self.device, C.byref(h)), "cs_ctx_create")
		self.handle = h
		self.torch_device = torch.device("cuda", self.device)

	def close(self):
		if getattr(self, "handle", None):
			self.lib.cs_ctx_destroy(self.handle)
			self.handle = None

	def __del__(self):
		try:
			self.close()
		except Exception:
			pass

	@property
	def sm_count(self) -> int:
		return self.lib.cs_ctx_sm_count(self.handle)

	def stream(self) -> int:
		"""Handle of torch's current stream on this device (the raw accessor: ~0.3 us instead of the ~10 us it takes
		to build a torch.cuda.Stream object — an entry point of the small-image configs makes ~150 calls)."""
		import torch

		raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
		if raw is not None:
			return raw(self.device)
		return torch.cuda.current_stream(self.torch_device).cuda_stream

	def call(self, name: str, *args) -> None:
		check(getattr(self.lib, name)(self.handle, *args), name)


_contexts: dict[int, Context] = {}


def get_context(device: int | None = None) -> Context:
	import torch

	if device is None:
		device = torch.cuda.current_device() if torch.cuda.is_available() else 0
	ctx = _contexts.get(device)
	if ctx is None:
		ctx = _contexts[device] = Context(device)
	return ctx


def ptr(t) -> int:
	"""Device (or host) address of a tensor / numpy array; None -> NULL."""
	if t is None:
		return None
	if hasattr(t, "data_ptr"):
		return t.data_ptr()
	return t.ctypes.data
