set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -2 gpurun_out/bench_r1f.err; cut -c1-900 gpurun_out/bench_r1f.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01_launches_ppo_iter_v3.csv python tools/ncu_target.py ppo > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_fwd_tc|gemm_tf32x3|conv1_wgrad_kernel" -c 8 -o gpurun_out/r01_tc_kernels_full python tools/ncu_target.py ppo > gpurun_out/ncu_tc_full.log 2>&1; tail -1 gpurun_out/ncu_tc_full.log
python tools/microbench.py conv gemm > gpurun_out/r01_microbench_conv_gemm.jsonl 2>&1; tail -3 gpurun_out/r01_microbench_conv_gemm.jsonl | cut -c1-300
