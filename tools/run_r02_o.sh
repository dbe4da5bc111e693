#!/bin/bash
# round 2, GPU call o: A/B of the alternating master's polling on ONE box
OUT=gpurun_out/r02o
mkdir -p $OUT
lscpu | grep -i "model name\|^CPU(s)\|Thread\|NUMA node0" | head -5
cat /sys/devices/system/cpu/cpu0/topology/thread_siblings_list /sys/devices/system/cpu/cpu1/topology/thread_siblings_list
for cfg in "0 once" "1 once" "1 spin" "1 yield" "0 once" "1 once"; do
set -- $cfg
RLPYT_B200_SAMPLER_CHUNKED=$1 RLPYT_B200_SAMPLER_POLL=$2 RLPYT_B200_SAMPLER_PROFILE=1 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $OUT/bench_$1_$2.json 2> $OUT/bench_$1_$2.err; python - $1 $2 <<'PY'
import json,sys
l=[x for x in open('gpurun_out/r02o/bench_%s_%s.json' % (sys.argv[1], sys.argv[2])) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); p=d['e2e']['sampler_profile_us_per_env_step_rank0']
    print('chunked', sys.argv[1], sys.argv[2], round(d['e2e']['value']), round(d['e2e']['sampling_ms_per_step_rank0'],1), {k: round(v,1) for k,v in p.items()})
PY
done
