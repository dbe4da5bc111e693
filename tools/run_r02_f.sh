#!/bin/bash
set -x
OUT=gpurun_out/r02_f
mkdir -p $OUT
timeout 100 tools/probes/_bin/conv2_s2d 8 > $OUT/mn_probe.txt 2>&1; cat $OUT/mn_probe.txt
timeout 1200 python -m pytest tests -m gpu -q --durations=15 > $OUT/pytest_gpu.txt 2>&1; tail -60 $OUT/pytest_gpu.txt
