#!/bin/bash
# round 2, GPU call u: vectorised collector writes + arena allocations: tests, sampler A/B, replay workload
OUT=gpurun_out/r02u
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_replay.py tests/test_gpu_replay_sequence.py tests/test_gpu_dqn.py tests/test_gpu_async.py tests/test_gpu_runner_dropin.py -q -m gpu 2>&1 | grep -v "Warning\|warnings\|^  \|^$\|^tests/" | tail -10
for cfg in "0 once" "1 yield" "0 once" "1 yield"; do
set -- $cfg
RLPYT_B200_SAMPLER_CHUNKED=$1 RLPYT_B200_SAMPLER_POLL=$2 RLPYT_B200_SAMPLER_PROFILE=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench_$1_$2.json 2> $OUT/bench_$1_$2.err; python - $1 $2 <<'PY'
import json,sys
l=[x for x in open('gpurun_out/r02u/bench_%s_%s.json' % (sys.argv[1], sys.argv[2])) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); p=d['e2e']['sampler_profile_us_per_env_step_rank0']
    print('chunked', sys.argv[1], sys.argv[2], round(d['value']), round(d['e2e']['value']), round(d['e2e']['sampling_ms_per_step_rank0'],1), {k: round(v,1) for k,v in p.items()})
PY
done
timeout 300 python bench.py --workload replay > $OUT/bench_workload_replay.json 2> $OUT/replay.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02u/bench_workload_replay.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['value'], d.get('phases_us'), d['roofline']['frac'])
PY
