#!/bin/bash
# round 2, GPU call i: TMEM-operand GEMM - tests, kernel comparison, learner bench
OUT=gpurun_out/r02i
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu -k "gemm or linear" > $OUT/tests.log 2>&1; echo "tests rc=$?" >> $OUT/tests.log; tail -5 $OUT/tests.log
timeout 200 python tools/gemm_compare.py > $OUT/gemm_compare.jsonl 2> $OUT/gemm_compare.err; cat $OUT/gemm_compare.jsonl; tail -3 $OUT/gemm_compare.err
timeout 600 python bench.py --steps 6 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -c 1500 $OUT/bench.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02i/bench.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e'].get('sampling_ms_per_step_rank0'))
    for k in d.get('step_kernels',[]): print(k['kernel'][:60], round(k['us_per_launch'],1), round(k['frac'],3), round(k['share_of_step'],3))
PY
