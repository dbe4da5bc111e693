#!/bin/bash
# round 2, final GPU call: what the driver runs at round end (full gpu suite, smoke, default bench, reference arm) + the
# workload lines for BASELINE.json configs[1] and [3]; everything lands under gpurun_out/r02final
OUT=gpurun_out/r02final
mkdir -p $OUT
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > $OUT/tests.log 2>&1; echo "tests rc=$? secs=$(( $(date +%s)-S ))" >> $OUT/tests.log; grep -v "Warning\|warnings\|^  \|^$" $OUT/tests.log | tail -22
S=$(date +%s)
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; echo "smoke secs=$(( $(date +%s)-S ))"
S=$(date +%s)
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? secs=$(( $(date +%s)-S ))"; tail -c 400 $OUT/bench.err
S=$(date +%s)
timeout 600 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm rc=$? secs=$(( $(date +%s)-S ))"; tail -c 300 $OUT/bench_reference.err
for w in gae replay dqn; do
S=$(date +%s)
timeout 300 python bench.py --workload $w > $OUT/bench_workload_$w.json 2> $OUT/bench_workload_$w.err; echo "workload $w rc=$? secs=$(( $(date +%s)-S ))"; tail -c 200 $OUT/bench_workload_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02final/bench*.json')):
    l=[x for x in open(f) if x.startswith('{')]
    if not l: print(f, 'NO LINE'); continue
    d=json.loads(l[-1])
    print(f.split('/')[-1], {k:d.get(k) for k in ('value','unit','ms_per_step','gpu_launches')}, (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'))
    for k in d.get('step_kernels',[]): print('   ', k['kernel'][:60], round(k['us_per_launch'],1), round(k['frac'],3), round(k['share_of_step'],3))
PY
