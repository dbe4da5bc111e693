#!/bin/bash
# round 2, GPU call q: fused ReLU backward tests, bench, ncu launch list + full captures of the new kernels
OUT=gpurun_out/r02q
mkdir -p $OUT
S=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_gemm.py tests/test_gpu_recurrent.py -q -m gpu 2>&1 | grep -v "Warning\|warnings\|^  \|^$" | tail -8; echo "tests secs=$(( $(date +%s)-S ))"
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; tail -c 300 $OUT/bench.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02q/bench.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e'].get('sampling_ms_per_step_rank0'))
    for k in d.get('step_kernels',[]): print(k['kernel'][:60], round(k['us_per_launch'],1), round(k['frac'],3), round(k['share_of_step'],3))
PY
S=$(date +%s)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_ppo_iter.csv python tools/ncu_target.py ppo > $OUT/ncu_launch.log 2>&1; echo "launch list rc=$? secs=$(( $(date +%s)-S ))"; wc -l $OUT/launches_ppo_iter.csv
NCU="ncu --set full --import-source on --clock-control none --profile-from-start off"
S=$(date +%s)
timeout 300 $NCU -k regex:"conv2_s2d_wgrad_kernel" -c 1 -o $OUT/conv2_s2d_wgrad python tools/ncu_target.py convbwd > $OUT/a.log 2>&1; tail -1 $OUT/a.log
timeout 300 $NCU -k regex:"gemm_ts_kernel" -c 3 -o $OUT/gemm_ts python tools/ncu_target.py ppo > $OUT/b.log 2>&1; tail -1 $OUT/b.log
timeout 300 $NCU -k regex:"replay_extract_bulk_kernel|sumtree_batch_leaves_kernel|sumtree_propagate_kernel|sumtree_find_kernel" -c 8 -o $OUT/replay_kernels python tools/ncu_target.py replay > $OUT/c.log 2>&1; tail -1 $OUT/c.log
echo "full captures secs=$(( $(date +%s)-S ))"; ls -la $OUT | tail -8
