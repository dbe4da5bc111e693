#!/bin/bash
# First GPU call of round 2: everything that was prepared at the end of round 1 without a GPU.
#   gpurun --timeout 1500 -- 'bash tools/run_round2_first.sh'
# Outputs land in gpurun_out/r02_first/ (merged back by gpurun).
set -x
OUT=gpurun_out/r02_first
mkdir -p $OUT
[ -x tools/probes/_bin/shift_probe ] || bash tools/probes/build.sh > $OUT/probe_build.log 2>&1
# 1. hardware semantics + v2 conv probes (tools/probes/README.md)
timeout 120 tools/probes/_bin/shift_probe > $OUT/shift_probe.txt 2>&1; tail -8 $OUT/shift_probe.txt
for bo in 0 1; do timeout 120 tools/probes/_bin/conv1_v2 8192 $bo > $OUT/conv1_v2_bo$bo.txt 2>&1; cat $OUT/conv1_v2_bo$bo.txt; done
timeout 120 tools/probes/_bin/conv1_wgrad_v2 8192 > $OUT/conv1_wgrad_v2.txt 2>&1; cat $OUT/conv1_wgrad_v2.txt
# 2. sampler handshake modes: e2e of the default bench with each (value is unaffected)
for mode in futex hybrid spin; do
  RLPYT_B200_SAMPLER_SYNC=$mode timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_sync_$mode.json 2> $OUT/bench_sync_$mode.err
  python -c "import json;d=json.load(open('$OUT/bench_sync_$mode.json'));print('$mode', d['value'], d['e2e'])"
done
# 2b. alternating sampler (master side never run on a GPU yet)
RLPYT_B200_BENCH_SAMPLER=alternating timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_alternating.json 2> $OUT/bench_alternating.err
tail -3 $OUT/bench_alternating.err; python -c "import json;d=json.load(open('$OUT/bench_alternating.json'));print('alternating', d['value'], d['e2e'])"
RLPYT_B200_SAMPLER_SYNC=hybrid timeout 600 python -m pytest tests/test_gpu_sampler.py -x -q 2>&1 | tail -3
RLPYT_B200_SAMPLER_SYNC=spin timeout 600 python -m pytest tests/test_gpu_sampler.py -x -q 2>&1 | tail -3
# 3. DQN learner benchmark (parity-green in round 1, unmeasured)
timeout 600 python tools/bench_dqn.py > $OUT/bench_dqn.json 2> $OUT/bench_dqn.err; tail -2 $OUT/bench_dqn.err; cut -c1-700 $OUT/bench_dqn.json
# 4. evaluate_agent (added at the end of round 1, CPU-tested only): one evaluation through the GpuSampler
timeout 300 python - <<'PY' 2>&1 | tail -3
import torch
from rlpyt_b200.agents.pg.atari import AtariFfAgent
from rlpyt_b200.envs.synthetic import SyntheticAtariEnv
from rlpyt_b200.samplers.parallel.gpu.sampler import GpuSampler
kw = dict(image_shape=(4, 84, 84), n_actions=6, p_done=0.05)
s = GpuSampler(EnvCls=SyntheticAtariEnv, env_kwargs=kw, batch_T=8, batch_B=8, max_decorrelation_steps=0,
               eval_n_envs=4, eval_max_steps=400, eval_max_trajectories=10)
agent = AtariFfAgent()
s.initialize(agent, affinity=dict(cuda_idx=0, workers_cpus=[0, 1]), seed=0, bootstrap_value=True)
infos = s.evaluate_agent(itr=1)
print("evaluate_agent:", len(infos), "trajectories, mean length", sum(i.Length for i in infos) / max(1, len(infos)))
s.obtain_samples(0)
s.shutdown()
PY
