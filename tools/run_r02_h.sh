#!/bin/bash
set -x
OUT=gpurun_out/r02_h
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_recurrent.py tests/test_gpu_runner_dropin.py tests/test_gpu_gemm.py -m gpu -q -rf --tb=short > $OUT/pytest_sel.txt 2>&1; tail -40 $OUT/pytest_sel.txt | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; tail -5 $OUT/bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_h/bench_n2.json'))
print('N=2 value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',d['e2e'], 'identical', d.get('params_identical_across_ranks'), d['config'].get('worker_cpus_rank0'))
PY
timeout 400 python bench.py --steps 8 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -3 $OUT/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_h/bench_n1.json'))
print('N=1 value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',d['e2e'])
print(d.get('cpu_baseline')); print(d.get('gae_ppo_loss'))
for k in d.get('step_kernels',[]): print(k['kernel'][:60], round(k['us_per_launch'],1), round(k['frac'],3), round(k['share_of_step'],3))
PY
