set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; tail -2 gpurun_out/bench_r1j.err; cut -c1-900 gpurun_out/bench_r1j.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01_launches_ppo_iter_v6.csv python tools/ncu_target.py ppo > /dev/null 2>&1
