#!/bin/bash
# round 2, GPU call p: agent.step's first layer streaming the step buffer over PCIe itself (zero-copy sampler steps)
OUT=gpurun_out/r02p
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sampler.py -q -m gpu -x 2>&1 | grep -v "Warning\|warnings\|^  \|^$" | tail -12
for z in 0 1 0 1; do
RLPYT_B200_SAMPLER_ZEROCOPY=$z RLPYT_B200_SAMPLER_PROFILE=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench_z$z.json 2> $OUT/bench_z$z.err; tail -c 300 $OUT/bench_z$z.err; python - $z <<'PY'
import json,sys
l=[x for x in open('gpurun_out/r02p/bench_z%s.json' % sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); p=d['e2e']['sampler_profile_us_per_env_step_rank0']
    print('zerocopy', sys.argv[1], round(d['value']), round(d['e2e']['value']), round(d['e2e']['sampling_ms_per_step_rank0'],1), {k: round(v,1) for k,v in p.items()})
PY
done
RLPYT_B200_SAMPLER_ZEROCOPY=1 timeout 120 python tools/sampler_halfstep_probe.py > $OUT/halfstep_zc.json 2> $OUT/halfstep_zc.err; cat $OUT/halfstep_zc.json; tail -2 $OUT/halfstep_zc.err
