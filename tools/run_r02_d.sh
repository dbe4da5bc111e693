#!/bin/bash
# Round 2, fourth GPU call: full GPU suite with conv2 s2d + alternating tests, bench line, launch list, workloads.
set -x
OUT=gpurun_out/r02_d
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; tail -6 $OUT/pytest_gpu.txt
RLPYT_B200_SAMPLER_PROFILE=1 timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -3 $OUT/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_d/bench_n1.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',d['e2e'])
print('cpu_baseline',d.get('cpu_baseline'))
for k in d.get('step_kernels',[]): print(k['kernel'][:60], round(k['us_per_launch'],1), round(k['frac'],3), round(k['share_of_step'],3))
PY
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/r02_launches_ppo_iter_v2.csv python tools/ncu_target.py ppo > /dev/null 2>&1
tail -2 $OUT/r02_launches_ppo_iter_v2.csv
for w in gae replay dqn; do timeout 600 python bench.py --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err; tail -2 $OUT/bench_$w.err; cut -c1-600 $OUT/bench_$w.json; done
