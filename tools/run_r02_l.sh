#!/bin/bash
# round 2, GPU call l: async test fix, sampler step profile (where do the 410 us of a device step go)
OUT=gpurun_out/r02l
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_async.py -q -m gpu -x 2>&1 | grep -v "Warning\|^  \|^$" | tail -8
RLPYT_B200_SAMPLER_PROFILE=1 timeout 300 python bench.py --steps 5 --warmup 3 > $OUT/bench_prof.json 2> $OUT/bench_prof.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02l/bench_prof.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
PY
timeout 120 python tools/sampler_halfstep_probe.py > $OUT/halfstep.json 2> $OUT/halfstep.err; cat $OUT/halfstep.json; tail -3 $OUT/halfstep.err
