"""Asynchronous samplers (mirror of ``rlpyt/samplers/async_/base.py:8-95``, ``async_/serial_sampler.py``,
``async_/gpu_sampler.py``, ``async_/alternating_sampler.py``): the sampler the
asynchronous runner drives - ``async_initialize`` returns a DOUBLE BUFFER of sample batches and the examples,
``initialize(affinity)`` brings the sampler up, ``obtain_samples(itr, db_idx)`` fills buffer ``db_idx`` and returns the
completed trajectory infos.

B200 design, behind those three methods.  The reference runs the sampler in a forked process that writes OS shared
memory, because its replay buffer and its optimizer live in other processes.  Here everything that consumes samples is
in HBM of the one GPU this process owns, so the asynchronous sampler is a THREAD of that process with its own CUDA
stream(s): the batch is collected by the same step engines as in synchronous mode (CPU env workers -> pinned step
buffer -> ``observation[t]`` in HBM, ``agent.step`` as one CUDA graph per step), then published to
``double_buffer[db_idx]`` by device-to-device copies (a [T,B] batch moves at HBM rate: microseconds, against the
memory-copier PROCESSES of the reference), with CUDA events handing it to the consumer's stream.  The sampler acts with
its own copy of the parameters (``agent.async_twin()``), refreshed between batches from the optimizer's staging copy
(``recv_shared_memory``), exactly the reference's hand-off (agents/base.py:218-243) without leaving the device.
"""
import torch

from rlpyt_b200.samplers.buffer import build_samples_buffer, get_example_outputs
from rlpyt_b200.samplers.parallel.gpu.alternating_sampler import AlternatingSampler
from rlpyt_b200.samplers.parallel.gpu.sampler import GpuSampler
from rlpyt_b200.samplers.serial.sampler import SerialSampler
from rlpyt_b200.utils.seed import make_seed


class AsyncSamplerMixin:

    async_ = True

    # ---- master (runner) side -----------------------------------------------------------------------------------
    def async_initialize(self, agent, bootstrap_value=False, traj_info_kwargs=None, seed=None, device=None):
        """async_/base.py:18-48: initialize the agent from an example environment, pre-allocate the double buffer
        (in HBM) and return it with the examples."""
        self.seed = make_seed() if seed is None else seed
        if device is None:
            if not torch.cuda.is_available():
                from rlpyt_b200 import _lib
                raise _lib.B200LibraryError("rlpyt_b200 asynchronous samplers keep their batches in HBM: a CUDA device is required")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        B = self.batch_spec.B
        env = self.EnvCls(**self.env_kwargs)
        agent.initialize(env.spaces, share_memory=True, global_B=B, env_ranks=list(range(B)))
        examples = get_example_outputs(agent, env)
        self.double_buffer = tuple(
            build_samples_buffer(agent, env, self.batch_spec, bootstrap_value, device=self.device, examples=examples)[0]
            for _ in range(2))
        env.close()
        if traj_info_kwargs:
            for k, v in traj_info_kwargs.items():
                setattr(self.TrajInfoCls, "_" + k, v)
        self.examples = examples
        self.agent = agent
        self._bootstrap_value = bootstrap_value
        self._published = [None, None]       # event: buffer i holds a complete batch (recorded on the sampler's stream)
        self._consumed = [None, None]        # event: the consumer has read buffer i  (recorded on the consumer's stream)
        return self.double_buffer, examples

    # ---- sampler side ---------------------------------------------------------------------------------------------
    def initialize(self, affinity):
        """async_/serial_sampler.py:31-77 / async_/gpu_sampler.py:40-70.  Called by the runner BEFORE it starts its
        threads (the parallel samplers fork their env workers here), after the optimizer's agent went to the device:
        the sampler gets its own parameter copy on the same GPU."""
        master = self.agent
        twin = master.async_twin()
        self._agent_preinitialized = True
        affinity = dict(affinity or {})
        affinity.setdefault("cuda_idx", self.device.index)
        super().initialize(twin, affinity=affinity, seed=self.seed, bootstrap_value=self._bootstrap_value)
        self.master_agent = master
        self.agent = twin
        for ro in ([self.rollout] if hasattr(self, "rollout") else []) + list(getattr(self, "rollouts", [])):
            ro.capture_error_mode = "thread_local"           # the optimizer thread keeps issuing CUDA calls while a step graph is captured

    def obtain_samples(self, itr, db_idx):
        """async_/serial_sampler.py:79-91 / async_/base.py:62-74: take new parameters if the optimizer published any,
        collect one batch, publish it in ``double_buffer[db_idx]``; returns the completed trajectory infos."""
        self.agent.recv_shared_memory()
        samples, traj_infos = super().obtain_samples(itr)
        stream = torch.cuda.current_stream(self.device)
        if self._consumed[db_idx] is not None:
            stream.wait_event(self._consumed[db_idx])          # the copier's reads of the previous use of this buffer
        self.double_buffer[db_idx][:] = samples                 # device-to-device, leaf by leaf (env_info: host)
        self._published[db_idx] = stream.record_event()
        return traj_infos

    def evaluate_agent(self, itr):
        """async_/serial_sampler.py:93-98."""
        self.agent.recv_shared_memory()
        return super().evaluate_agent(itr)

    # ---- consumer side (the runner's copier thread) ---------------------------------------------------------------
    def acquire_batch(self, db_idx, stream=None):
        """Make ``stream`` (default: the calling thread's current stream) wait until buffer ``db_idx`` is complete."""
        if self._published[db_idx] is not None:
            (stream or torch.cuda.current_stream(self.device)).wait_event(self._published[db_idx])
        return self.double_buffer[db_idx]

    def release_batch(self, db_idx, stream=None):
        """The consumer's reads of buffer ``db_idx`` have been enqueued on ``stream``: the sampler may overwrite it
        once they have run."""
        self._consumed[db_idx] = (stream or torch.cuda.current_stream(self.device)).record_event()


class AsyncSerialSampler(AsyncSamplerMixin, SerialSampler):
    """async_/serial_sampler.py:12-98."""


class AsyncGpuSampler(AsyncSamplerMixin, GpuSampler):
    """async_/gpu_sampler.py:18-120 (one GPU; the sampler shares the optimizer's B200)."""


class AsyncAlternatingSampler(AsyncSamplerMixin, AlternatingSampler):
    """async_/alternating_sampler.py:11-100."""
