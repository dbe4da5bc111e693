#!/bin/bash
# Round 2, second GPU call: full GPU test suite on the i8 conv1 path, the bench line, sampler variants, launch list.
set -x
OUT=gpurun_out/r02_b
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; tail -5 $OUT/pytest_gpu.txt
RLPYT_B200_SAMPLER_PROFILE=1 timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -3 $OUT/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_b/bench_n1.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'])
print('cpu_baseline',d.get('cpu_baseline'))
for k in d.get('step_kernels',[]): print(k['kernel'][:60], round(k['us_per_launch'],1), round(k['frac'],3), round(k['share_of_step'],3))
PY
for cfg in "--smt-workers" "--workers 14 --smt-workers"; do
  RLPYT_B200_SAMPLER_PROFILE=1 timeout 300 python bench.py --no-cpu-baseline $cfg > $OUT/bench_tmp.json 2>/dev/null
  python -c "import json;d=json.load(open('$OUT/bench_tmp.json'));print('$cfg', d['config']['env_workers_per_rank'], d['e2e'])"
done
RLPYT_B200_BENCH_SAMPLER=alternating RLPYT_B200_SAMPLER_PROFILE=1 timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_alt.json 2>/dev/null
python -c "import json;d=json.load(open('$OUT/bench_alt.json'));print('alternating', d['e2e'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -2 $OUT/bench_ref.err; cut -c1-1800 $OUT/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/r02_launches_ppo_iter.csv python tools/ncu_target.py ppo > /dev/null 2>&1
tail -2 $OUT/r02_launches_ppo_iter.csv
