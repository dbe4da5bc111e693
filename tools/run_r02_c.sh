#!/bin/bash
# Round 2, third GPU call: conv2 v2 probes + sampler configurations (clean per-step profile).
set -x
OUT=gpurun_out/r02_c
mkdir -p $OUT
timeout 300 tools/probes/_bin/conv2_s2d 3 > $OUT/conv2_probe.txt 2>&1; cat $OUT/conv2_probe.txt
run() { name=$1; shift; RLPYT_B200_SAMPLER_PROFILE=1 timeout 300 "$@" python bench.py --no-cpu-baseline --steps 8 --warmup 3 $EXTRA > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python -c "import json;d=json.load(open('$OUT/bench_$name.json'));e=d['e2e'];print('$name', d['config']['env_workers_per_rank'], round(d['ms_per_step'],1), round(e['ms_per_step'],1), round(e['sampling_ms_per_step_rank0'],1), e['sampler_profile_us_per_env_step_rank0'])"; }
EXTRA="" run std7 env
EXTRA="--smt-workers" run std14smt env
EXTRA="" run alt7 env RLPYT_B200_BENCH_SAMPLER=alternating
EXTRA="--smt-workers" run alt14smt env RLPYT_B200_BENCH_SAMPLER=alternating
EXTRA="" run std7_hybrid env RLPYT_B200_SAMPLER_SYNC=hybrid
EXTRA="" run alt7_hybrid env RLPYT_B200_BENCH_SAMPLER=alternating RLPYT_B200_SAMPLER_SYNC=hybrid
