#!/bin/sh
# builds the microbenchmarks next to their sources (binaries are git-ignored; they travel to the GPU box with gpurun)
set -e
cd "$(dirname "$0")"
for f in pipes pipes2 core fence_peer; do
	nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o "$f" "$f.cu"
done
