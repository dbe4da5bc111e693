#!/bin/bash
# round 2, GPU call r: fused policy/value heads, channel maxima hand-over, ReLU fusion opt-in; bench + launch list
OUT=gpurun_out/r02r
mkdir -p $OUT
S=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_heads.py tests/test_gpu_ppo.py tests/test_gpu_gemm.py tests/test_gpu_recurrent.py tests/test_gpu_loss.py tests/test_gpu_runner_dropin.py -q -m gpu > $OUT/tests.log 2>&1; echo "tests rc=$? secs=$(( $(date +%s)-S ))" >> $OUT/tests.log; grep -v "Warning\|warnings\|^  \|^$\|^tests/" $OUT/tests.log | tail -25
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; tail -c 300 $OUT/bench.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02r/bench.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e'].get('sampling_ms_per_step_rank0'))
    for k in d.get('step_kernels',[]): print(k['kernel'][:60], round(k['us_per_launch'],1), round(k['frac'],3), round(k['share_of_step'],3))
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_ppo_iter.csv python tools/ncu_target.py ppo > $OUT/ncu_launch.log 2>&1; echo "launch list rc=$?"; wc -l $OUT/launches_ppo_iter.csv
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"pg_heads" -c 3 -o $OUT/pg_heads python tools/ncu_target.py ppo > $OUT/a.log 2>&1; tail -1 $OUT/a.log
