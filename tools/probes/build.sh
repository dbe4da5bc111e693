#!/bin/bash
# builds the round-2 probes into tools/probes/_bin/ (git-ignored, shipped to the GPU box by gpurun)
set -e
cd "$(dirname "$0")/../.."
mkdir -p tools/probes/_bin
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I rlpyt_b200/csrc"
nvcc $FLAGS -o tools/probes/_bin/shift_probe tools/probes/tcgen05_shift_probe.cu
nvcc $FLAGS -o tools/probes/_bin/conv1_v2 tools/probes/conv1_v2_probe.cu
nvcc $FLAGS -o tools/probes/_bin/conv1_wgrad_v2 tools/probes/conv1_wgrad_v2_probe.cu
ls -la tools/probes/_bin
