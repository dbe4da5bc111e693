#!/bin/bash
# ncu --set full captures of the round-2 kernels on the probe binaries (N = 8192); .ncu-rep files come back in gpurun_out/
set -x
OUT=gpurun_out/r02_ncu
mkdir -p $OUT
NCU="ncu --set full --import-source on --clock-control none"
timeout 300 $NCU -k regex:conv1_i8_fwd_kernel -s 8 -c 1 -o $OUT/conv1_i8_fwd tools/probes/_bin/conv1_i8 8 > $OUT/a.log 2>&1; tail -2 $OUT/a.log
timeout 300 $NCU -k regex:"conv1_i8_wgrad_kernel|absmax_kernel|wgrad_i8_reduce" -s 9 -c 3 -o $OUT/conv1_i8_wgrad tools/probes/_bin/conv1_i8 32 > $OUT/b.log 2>&1; tail -2 $OUT/b.log
timeout 300 $NCU -k regex:"conv2_s2d_fwd_kernel" -s 33 -c 1 -o $OUT/conv2_s2d_fwd tools/probes/_bin/conv2_s2d 1 > $OUT/c.log 2>&1; tail -2 $OUT/c.log
timeout 300 $NCU -k regex:"conv2_s2d_dgrad_kernel" -s 31 -c 1 -o $OUT/conv2_s2d_dgrad tools/probes/_bin/conv2_s2d 2 > $OUT/d.log 2>&1; tail -2 $OUT/d.log
timeout 300 $NCU -k regex:"conv2_s2d_wgrad_kernel" -s 6 -c 1 -o $OUT/conv2_s2d_wgrad tools/probes/_bin/conv2_s2d 4 > $OUT/e.log 2>&1; tail -2 $OUT/e.log
ls -la $OUT
