#!/bin/bash
# round 2, GPU call w: fixes after call v (foreign gradients under graph replay, sentinel drain) - the training and sampler suites + bench
OUT=gpurun_out/r02w
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_recurrent.py tests/test_gpu_dqn.py tests/test_gpu_r2d1.py tests/test_gpu_runner_dropin.py tests/test_gpu_async.py tests/test_gpu_sampler.py -q -m gpu > $OUT/tests.log 2>&1; echo "rc=$?" >> $OUT/tests.log; grep -v "Warning\|warnings\|^  \|^$\|^tests/" $OUT/tests.log | tail -25
timeout 300 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 300 $OUT/bench.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02w/bench.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e'].get('sampling_ms_per_step_rank0'))
PY
