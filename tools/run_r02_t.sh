#!/bin/bash
# round 2, GPU call t: replay workload with the large-batch asymptote, ncu --set full of the final conv1 wgrad / conv2 dgrad / heads
OUT=gpurun_out/r02t
mkdir -p $OUT
timeout 300 python bench.py --workload replay > $OUT/bench_workload_replay.json 2> $OUT/replay.err; tail -c 300 $OUT/replay.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02t/bench_workload_replay.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['value'], d.get('phases_us'), d['roofline']['frac'], d['roofline'].get('at_batch_8192'))
PY
NCU="ncu --set full --import-source on --clock-control none --profile-from-start off"
timeout 300 $NCU -k regex:"conv1_i8_wgrad_kernel|wgrad_i8_reduce_kernel|conv2_s2d_dgrad_kernel|pg_heads" -c 6 -o $OUT/final_kernels python tools/ncu_target.py ppo > $OUT/a.log 2>&1; tail -1 $OUT/a.log
