#!/bin/bash
# Development: build variants of libcolorsimplify.so that differ in lloyd.cu only (different -D defines, or another
# source file) into image_segmenter_b200/_lib/variants/<name>.so; select one with COLORSIMPLIFY_LIB=<path>.
# usage: tools/build_variants.sh name[:src.cu][:"-DX=1 -DY=2"] ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OBJ=$ROOT/image_segmenter_b200/_lib/obj
OUT=$ROOT/image_segmenter_b200/_lib/variants
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OTHERS=$(ls "$OBJ"/*.o | grep -v '/lloyd.o$')
for spec in "$@"; do
	IFS=':' read -r name src defs <<<"$spec"
	src=${src:-$ROOT/image_segmenter_b200/csrc/lloyd.cu}
	(
		$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I "$ROOT/include" \
			-I "$ROOT/image_segmenter_b200/csrc" $defs -c "$src" -o "$OUT/$name.lloyd.o" 2>/dev/null
		$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/$name.so" $OTHERS "$OUT/$name.lloyd.o"
		echo "built $OUT/$name.so"
	) &
done
wait
