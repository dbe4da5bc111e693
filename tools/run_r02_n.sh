#!/bin/bash
# round 2, GPU call n: why did per-worker uploads slow the sampler down - copy probe + A/B of the switch
OUT=gpurun_out/r02n
mkdir -p $OUT
timeout 120 python tools/upload_chunk_probe.py > $OUT/upload_probe.json 2> $OUT/upload_probe.err; cat $OUT/upload_probe.json; tail -3 $OUT/upload_probe.err
for c in 0 1; do
RLPYT_B200_SAMPLER_CHUNKED=$c RLPYT_B200_SAMPLER_PROFILE=1 timeout 300 python bench.py --steps 5 --warmup 3 > $OUT/bench_c$c.json 2> $OUT/bench_c$c.err; python - $c <<'PY'
import json,sys
l=[x for x in open('gpurun_out/r02n/bench_c%s.json' % sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('chunked', sys.argv[1], {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['sampling_ms_per_step_rank0'], d['e2e']['sampler_profile_us_per_env_step_rank0'])
PY
done
