"""In-tree build of libcolorsimplify.so (nvcc, sm_100a only).

`python -m image_segmenter_b200.build` compiles every .cu / .cpp under csrc/ into
`image_segmenter_b200/_lib/libcolorsimplify.so`.  nvcc cross-compiles without a GPU, so this
runs in the authoring container; the built .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "_lib"
OBJDIR = LIBDIR / "obj"
LIB = LIBDIR / "libcolorsimplify.so"
ROOT = PKG.parent

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "-I", str(ROOT / "include")]


def _sources():
	return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def _digest(src: Path, extra: list[str]) -> str:
	h = hashlib.sha256()
	h.update(src.read_bytes())
	for hdr in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((ROOT / "include").glob("*.h"))):
		h.update(hdr.read_bytes())
	h.update(" ".join(extra).encode())
	return h.hexdigest()


def _compile(src: Path, defines: list[str], verbose: bool) -> Path:
	obj = OBJDIR / (src.stem + ".o")
	stamp = OBJDIR / (src.stem + ".sha")
	dig = _digest(src, defines)
	if obj.exists() and stamp.exists() and stamp.read_text() == dig:
		return obj
	cmd = [NVCC, *ARCH, *COMMON, *defines, "-c", str(src), "-o", str(obj)]
	if src.suffix == ".cpp":
		cmd.insert(1, "-x")
		cmd.insert(2, "cu")
	if verbose:
		print(" ".join(cmd), flush=True)
	subprocess.run(cmd, check=True)
	stamp.write_text(dig)
	return obj


def build(verbose: bool = False, tuning_variants: bool | None = None) -> Path:
	"""Compile (if stale) and link libcolorsimplify.so; returns its path."""
	if tuning_variants is None:
		tuning_variants = os.environ.get("CS_TUNING_VARIANTS", "0") == "1"
	defines = ["-DCS_TUNING_VARIANTS"] if tuning_variants else []
	defines += ["-D" + d for d in os.environ.get("CS_EXTRA_DEFINES", "").split() if d]  # development experiments
	if os.environ.get("CS_PHASE_TIMING", "0") == "1":
		defines.append("-DCS_PHASE_TIMING")  # development: globaltimer stamps of block 0 (tools/launch_overhead.py)
	OBJDIR.mkdir(parents=True, exist_ok=True)
	srcs = _sources()
	with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
		objs = list(ex.map(lambda s: _compile(s, defines, verbose), srcs))
	newest = max(o.stat().st_mtime for o in objs)
	if (not LIB.exists()) or LIB.stat().st_mtime < newest:
		cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs)]
		if verbose:
			print(" ".join(cmd), flush=True)
		subprocess.run(cmd, check=True)
	return LIB


if __name__ == "__main__":
	p = build(verbose=True, tuning_variants=("--tuning" in sys.argv) or None)
	print(p)
