#!/bin/bash
# round 2, GPU call k: async runner, bulk replay extraction, fused priority update, graphed PPO minibatch
OUT=gpurun_out/r02k
mkdir -p $OUT
S=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_async.py tests/test_gpu_replay.py tests/test_gpu_replay_sequence.py tests/test_gpu_ppo.py tests/test_gpu_dqn.py tests/test_gpu_r2d1.py tests/test_gpu_runner_dropin.py tests/test_gpu_sampler.py -q -m gpu > $OUT/tests.log 2>&1; echo "tests rc=$? secs=$(( $(date +%s)-S ))" >> $OUT/tests.log; grep -v "Warning\|warnings\|^  \|^$" $OUT/tests.log | tail -40
S=$(date +%s)
timeout 300 python bench.py --workload replay > $OUT/replay.json 2> $OUT/replay.err; echo "replay rc=$? secs=$(( $(date +%s)-S ))"; tail -c 600 $OUT/replay.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02k/replay.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ('value','ms_per_step')}, d.get('phases_us'), d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])
PY
S=$(date +%s)
timeout 600 python bench.py --steps 6 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? secs=$(( $(date +%s)-S ))"; tail -c 800 $OUT/bench.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02k/bench.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e'].get('sampling_ms_per_step_rank0'))
PY
