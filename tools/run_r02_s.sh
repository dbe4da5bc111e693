#!/bin/bash
# round 2, GPU call s: heads v2 check, 1-GPU quick bench
OUT=gpurun_out/r02s
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_heads.py tests/test_gpu_ppo.py -q -m gpu 2>&1 | grep -v "Warning\|warnings\|^  \|^$\|^tests/" | tail -12
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; tail -c 300 $OUT/bench.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02s/bench.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e'].get('sampling_ms_per_step_rank0'))
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_ppo_iter.csv python tools/ncu_target.py ppo > $OUT/ncu_launch.log 2>&1; grep -c "gpu__time" $OUT/launches_ppo_iter.csv
