set -x
python tools/microbench.py replay > gpurun_out/r01_microbench_replay.jsonl 2>&1; tail -7 gpurun_out/r01_microbench_replay.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01_launches_ppo_iter_v2.csv python tools/ncu_target.py ppo > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv1_fwd|conv1_wgrad_kernel|gemm_tf32x3|gather_rows_vec16|pg_loss_kernel|clip_adam" -c 8 -o gpurun_out/r01_ppo_kernels_full python tools/ncu_target.py ppo > gpurun_out/ncu_ppo_full.log 2>&1; tail -2 gpurun_out/ncu_ppo_full.log
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"sumtree|replay_extract|is_weights|pow_f32" -c 12 -o gpurun_out/r01_replay_kernels_full python tools/ncu_target.py replay > gpurun_out/ncu_replay_full.log 2>&1; tail -2 gpurun_out/ncu_replay_full.log
ls -la gpurun_out | tail -8
