#!/bin/bash
# builds the small measurement tools in-tree (binaries are git-ignored, they travel to the GPU box with gpurun)
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peak.bin fp64_peak.cu
