#!/bin/bash
# round 2, GPU call j (session 3 start): verify HEAD on a fresh box - full gpu suite, smoke, default bench line
OUT=gpurun_out/r02j
mkdir -p $OUT
S=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/tests.log 2>&1; echo "tests rc=$? secs=$(( $(date +%s)-S ))" >> $OUT/tests.log; tail -25 $OUT/tests.log
S=$(date +%s)
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; echo "smoke secs=$(( $(date +%s)-S ))"
S=$(date +%s)
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? secs=$(( $(date +%s)-S ))"; tail -c 800 $OUT/bench.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02j/bench.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e'].get('sampling_ms_per_step_rank0'))
    for k in d.get('step_kernels',[]): print(k['kernel'][:60], round(k['us_per_launch'],1), round(k['frac'],3), round(k['share_of_step'],3))
PY
