#!/bin/bash
# Round 2, fifth GPU call: conv2 wgrad probe, full suite with durations, sampler step breakdown, bench.
set -x
OUT=gpurun_out/r02_e
mkdir -p $OUT
timeout 200 tools/probes/_bin/conv2_s2d 4 > $OUT/conv2_wgrad_probe.txt 2>&1; cat $OUT/conv2_wgrad_probe.txt
timeout 100 tools/probes/_bin/conv1_i8 16 > $OUT/conv1_wgrad_probe.txt 2>&1; grep "us per" $OUT/conv1_wgrad_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=25 > $OUT/pytest_gpu.txt 2>&1; tail -45 $OUT/pytest_gpu.txt
timeout 200 python tools/sampler_step_breakdown.py > $OUT/sampler_step_breakdown.json 2> $OUT/ssb.err; cat $OUT/sampler_step_breakdown.json; tail -3 $OUT/ssb.err
RLPYT_B200_SAMPLER_PROFILE=1 timeout 600 python bench.py --no-cpu-baseline --steps 8 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -3 $OUT/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_e/bench_n1.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',d['e2e'])
for k in d.get('step_kernels',[]): print(k['kernel'][:60], round(k['us_per_launch'],1), round(k['frac'],3), round(k['share_of_step'],3))
PY
