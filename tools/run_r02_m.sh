#!/bin/bash
# round 2, GPU call m: per-worker uploads in the samplers, conv2 weight gradient with merged bx taps
OUT=gpurun_out/r02m
mkdir -p $OUT
S=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_async.py tests/test_gpu_runner_dropin.py tests/test_gpu_recurrent.py tests/test_gpu_gemm.py tests/test_gpu_ppo.py -q -m gpu > $OUT/tests.log 2>&1; echo "tests rc=$? secs=$(( $(date +%s)-S ))" >> $OUT/tests.log; grep -v "Warning\|warnings\|^  \|^$" $OUT/tests.log | tail -30
RLPYT_B200_SAMPLER_PROFILE=1 timeout 300 python bench.py --steps 6 --warmup 3 > $OUT/bench_prof.json 2> $OUT/bench_prof.err; tail -c 500 $OUT/bench_prof.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02m/bench_prof.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
    for k in d.get('step_kernels',[]): print(k['kernel'][:60], round(k['us_per_launch'],1), round(k['frac'],3), round(k['share_of_step'],3))
PY
