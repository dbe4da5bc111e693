#!/bin/bash
set -x
OUT=gpurun_out/r02_g
mkdir -p $OUT
timeout 100 tools/probes/_bin/conv2_s2d 12 > $OUT/mn_wgrad_probe.txt 2>&1; cat $OUT/mn_wgrad_probe.txt
timeout 900 python -m pytest tests -m gpu -q -rf --tb=short --durations=8 > $OUT/pytest_gpu.txt 2>&1; tail -120 $OUT/pytest_gpu.txt | cut -c1-250
RLPYT_B200_SAMPLER_PROFILE=1 timeout 400 python bench.py --no-cpu-baseline --steps 8 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -3 $OUT/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_g/bench_n1.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',d['e2e'])
PY
