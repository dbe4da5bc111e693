"""Alternating GPU sampler (mirror of ``rlpyt/samplers/parallel/gpu/alternating_sampler.py:8-85`` with the
step protocol of ``AlternatingActionServer``, ``rlpyt/samplers/parallel/gpu/action_server.py:123-173``;
SURVEY.md section 8(f) row 3): the workers form two groups, each owning half of the environments; while one
group steps its envs the master uploads, acts and records for the other group, so the GPU part of a step
(H2D + ``agent.step`` + D2H, ~0.35 ms) leaves the critical path when env stepping takes at least as long.

Built from the pieces of ``GpuSampler``: the same forked ``sampling_process`` workers and collectors, and
one ``DeviceRollout`` step engine per half operating on B-axis views of the same ``[T,B]`` HBM buffers and
of the pinned step buffer (each half replays its own per-step CUDA graphs).  Feed-forward agents only (the
recurrent alternating agents of the reference are outside the accelerated path).

Measured on the B200 host (profiles/r02_sampler_configs.txt, 8 cores per rank): with the two groups on sibling
hardware threads the env stepping disappears behind the device half-steps (59 ms per [128,256] batch against 76-79 ms
for the standard sampler); tests/test_gpu_sampler.py checks its batches against a host replay of the seeded envs and
against the reference GpuSampler's recorded batches, tests/test_sampler_protocol_cpu.py runs ``serve_actions``
against the real forked worker loop with a stand-in step engine.
"""
import os
import time

import numpy as np
import torch

from rlpyt_b200.samplers.parallel.gpu.sampler import GpuSampler
from rlpyt_b200.samplers.rollout import DeviceRollout


class AlternatingSampler(GpuSampler):

    alternating = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert self.batch_spec.B % 2 == 0, "Need even number for sampler batch_B."

    def initialize(self, agent, *args, **kwargs):
        if getattr(agent, "recurrent", False):
            assert getattr(agent, "alternating", False), "recurrent agents need the alternating state pair (AlternatingRecurrentAgentMixin)"
        else:
            agent.alternating = True       # alternating_sampler.py:37: a feed-forward agent only needs the flag
        examples = super().initialize(agent, *args, **kwargs)
        self._make_alternating_pairs()
        return examples

    def _get_n_envs_list(self, n_worker):
        """Even number of workers, the two halves mirror each other (alternating_sampler.py:64-80)."""
        B = self.batch_spec.B
        n_worker = min(n_worker, B) // 2 * 2
        if n_worker < 2:
            raise ValueError("AlternatingSampler needs at least 2 workers (affinity['workers_cpus'])")
        n_envs_list = [B // n_worker] * n_worker
        for w in range((B % n_worker) // 2):
            n_envs_list[w] += 1
            n_envs_list[w + n_worker // 2] += 1
        assert sum(n_envs_list) == B and sum(n_envs_list[:n_worker // 2]) == B // 2
        return n_envs_list

    def _make_alternating_pairs(self):
        half_w, B = self.n_worker // 2, self.batch_spec.B
        half_B = B // 2
        assert self.worker_slices[half_w].start == half_B
        self.halves = (slice(0, half_B), slice(half_B, B))
        self.obs_ready_pair = (self.sync.obs_ready[:half_w], self.sync.obs_ready[half_w:])
        self.act_ready_pair = (self.sync.act_ready[:half_w], self.sync.act_ready[half_w:])
        host = self.host
        self.rollouts = []
        for sl in self.halves:
            host_h = dict(step_np=host["step_np"][sl], step_pyt=host["step_pyt"][sl],
                          all_action=host["all_action"][:, sl], all_reward=host["all_reward"][:, sl],
                          pinned=host.get("pinned", False))
            ro = DeviceRollout(self.samples[:, sl], host_h, self.agent, self.device, stream=torch.cuda.Stream(self.device))
            ro.in_action.copy_(host_h["step_pyt"].action)
            w0 = 0 if sl.start == 0 else half_w
            ro.set_worker_chunks([slice(ws.start - sl.start, ws.stop - sl.start) for ws in self.worker_slices[w0:w0 + half_w]])
            self.rollouts.append(ro)

    def serve_actions(self, itr):
        """action_server.py:131-173 on the device step engines.  Each half owns a CUDA stream: while half A's
        ``agent.step`` runs, the master already takes half B's observations (their workers had A's whole turn to step)
        and starts their H2D copy, so per env step one of the two PCIe transfers leaves the critical path."""
        T = self.batch_spec.T
        wait_reset = not self.mid_batch_reset
        prof = self.profile if os.environ.get("RLPYT_B200_SAMPLER_PROFILE") == "1" else None
        clock = time.perf_counter
        if self.device.type == "cuda":
            current = torch.cuda.current_stream(self.device)
            for ro in self.rollouts:
                ro.side_stream.wait_stream(current)          # the learner's last update is ordered before this batch
        # Optionally (RLPYT_B200_SAMPLER_CHUNKED=1, RLPYT_B200_SAMPLER_POLL=spin|yield) each half's observations are uploaded
        # PER WORKER, as soon as that worker has signalled (rl_upload_async), with the master polling the other half's
        # workers while it waits for this half's agent.step.  (profiles/r02_sampler_halfstep.json: H2D of a half 100 us +
        # agent.step 94 us are otherwise serial per half - the other half's workers are never all done yet when the
        # master looks once right after launching agent.step.)
        # Measured on the B200 host (profiles/r02_sampler_poll_ab.txt): per-worker uploads + polling do take the H2D off the
        # device critical path (device part of a step 333 -> 268 us) but the env workers lose the same time - the two groups
        # sit on sibling hardware threads of the same 7 cores and now overlap more - so the defaults stay "one upload per
        # half, look once"; the switches remain for hosts with more cores per GPU.
        chunked = os.environ.get("RLPYT_B200_SAMPLER_CHUNKED", "0") == "1"   # 1: each worker's rows are uploaded as it signals
        poll_mode = os.environ.get("RLPYT_B200_SAMPLER_POLL", "once")        # once | spin | yield: how the master looks at the stepping half while agent.step runs
        pending = [list(range(len(p))) for p in self.obs_ready_pair]   # workers whose obs_ready for the half's next event is still to be taken
        uploaded = [False, False]
        half_w = len(self.obs_ready_pair[0])

        def poll(alt, k, block):
            """Take the obs_ready handshakes that are available (all of them when ``block``), uploading each worker's rows at once."""
            ro, sems, left = self.rollouts[alt], self.obs_ready_pair[alt], pending[alt]
            i = 0
            while i < len(left):
                w = left[i]
                if sems[w].acquire(block=block):
                    if chunked:
                        ro.upload_worker_rows(k, w)
                    left.pop(i)
                else:
                    i += 1
            return not left

        def finish_upload(alt, k):
            ro, sl = self.rollouts[alt], self.halves[alt]
            pending[alt] = list(range(half_w))
            done_now = ro.step_np.done
            if self.mid_batch_reset and np.any(done_now):
                for b in np.where(done_now)[0]:
                    self.agent.reset_one(idx=int(b) + sl.start)
            ro.upload_async(k, zero_inputs_on_done=True, obs_done=chunked)  # reward / done / agent inputs (+ observations if not chunked)

        for t in range(T):
            for alt in range(2):
                ro = self.rollouts[alt]
                t0 = clock() if prof is not None else 0.0
                if not uploaded[alt]:
                    poll(alt, t, True)                       # this half wrote obs(t), reward(t-1), done(t-1)
                    finish_upload(alt, t)
                uploaded[alt] = False
                t1 = clock() if prof is not None else 0.0
                ro.act_async(t, blank_done_rows=wait_reset)
                ta = clock() if prof is not None else 0.0
                other, t_other = alt ^ 1, (t if alt == 0 else t + 1)
                if t_other < T:
                    while not uploaded[other]:               # serve the other half's workers while this half's step runs
                        if poll(other, t_other, False):
                            finish_upload(other, t_other)
                            uploaded[other] = True
                        elif poll_mode == "once" or ro.act_done():
                            break
                        elif poll_mode == "yield":
                            os.sched_yield()
                tb = clock() if prof is not None else 0.0
                ro.wait()                                    # actions of this half are in the step buffer
                t2 = clock() if prof is not None else 0.0
                for s in self.act_ready_pair[alt]:
                    s.release()                              # this half steps while the other is served
                if prof is not None:
                    prof["wait_envs_s"] += t1 - t0
                    prof["device_step_s"] += t2 - t1
                    prof["release_s"] += clock() - t2
                    prof["act_launch_s"] = prof.get("act_launch_s", 0.0) + (ta - t1)
                    prof["other_upload_s"] = prof.get("other_upload_s", 0.0) + (tb - ta)
                    prof["stream_wait_s"] = prof.get("stream_wait_s", 0.0) + (t2 - tb)
                    prof["early_uploads"] = prof.get("early_uploads", 0) + (1 if uploaded[other] else 0)
                    prof["steps"] += 0.5                     # two half steps = one env step of the whole batch
        for alt in range(2):
            ro, sl = self.rollouts[alt], self.halves[alt]
            for s in self.obs_ready_pair[alt]:
                s.acquire()
            ro.finish()                                      # bootstrap value of this half
            ro.wait()                                        # the DMA out of the step buffer must have run before the host zeroes it below
            if np.any(ro.step_np.done):
                ended = np.where(ro.step_np.done)[0]
                ro.step_np.action[ended] = 0
                ro.step_np.reward[ended] = 0
                for b in ended:
                    self.agent.reset_one(idx=int(b) + sl.start)
                with torch.cuda.stream(ro.side_stream):
                    ro.zero_inputs_where_done()
            toggle = getattr(self.agent, "toggle_alt", None)  # value / reset do not advance the rnn state (action_server.py:168);
            if toggle is not None:                           # duck-typed feed-forward agents need not define it
                toggle()
        for ro in self.rollouts:
            ro.wait()
            ro.end_batch()
        for s in self.sync.obs_ready:
            assert not s.acquire(block=False)                # drained (action_server.py:170-173)
        for s in self.sync.act_ready:
            assert not s.acquire(block=False)
