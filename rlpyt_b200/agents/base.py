"""Agent base class: glue between sampler, network and algorithm (host-side mirror of
``rlpyt/agents/base.py:17-246`` plus the recurrent-state mixins of :252-371; the asynchronous-mode methods
``async_cpu`` / ``send_shared_memory`` / ``recv_shared_memory`` of :138-160,218-243 hand parameters from the
optimizer's agent to the sampler's through a staging copy in HBM, see ``async_twin``).

B200 design differences, all behind the unchanged method names:
* network outputs stay on the device (the reference copies them to the CPU, agents/pg/
  categorical.py:25,42); methods return tensors on the device their inputs came from, so a
  reference CPU sampler still gets CPU tensors back while the device-resident sampler and
  algorithms of this package never cross PCIe;
* ``data_parallel()`` does not wrap the model in DistributedDataParallel: parameters are
  broadcast from rank 0 once and gradients are summed by ONE all-reduce over the flat gradient
  buffer inside the optimizer step (``rlpyt_b200.algos.optim.FlatAdam``).
"""
import torch

from rlpyt_b200.utils.collections import namedarraytuple
from rlpyt_b200.models.utils import strip_ddp_state_dict

AgentInputs = namedarraytuple("AgentInputs", ["observation", "prev_action", "prev_reward"])
AgentStep = namedarraytuple("AgentStep", ["action", "agent_info"])


class BaseAgent:

    recurrent = False
    alternating = False

    def __init__(self, ModelCls=None, model_kwargs=None, initial_model_state_dict=None):
        self.ModelCls = ModelCls
        self.model_kwargs = dict() if model_kwargs is None else model_kwargs
        self.initial_model_state_dict = initial_model_state_dict
        self.model = None
        self.shared_model = None
        self.distribution = None
        self.device = torch.device("cpu")
        self._mode = None
        self.world_size = 1
        self._async = None             # parameter channel of asynchronous mode (async_twin)
        self._async_role = None

    def __call__(self, observation, prev_action, prev_reward):
        """Training forward pass (used by the algorithm)."""
        raise NotImplementedError

    def initialize(self, env_spaces, share_memory=False, **kwargs):
        """Build the model from the environment interface (rlpyt/agents/base.py:59-90)."""
        self.env_model_kwargs = self.make_env_to_model_kwargs(env_spaces)
        self.model = self.ModelCls(**self.env_model_kwargs, **self.model_kwargs)
        if share_memory:
            self.model.share_memory()
            self.shared_model = self.model
        if self.initial_model_state_dict is not None:
            self.model.load_state_dict(self.initial_model_state_dict)
        self.env_spaces = env_spaces
        self.share_memory = share_memory

    def make_env_to_model_kwargs(self, env_spaces):
        return {}

    def to_device(self, cuda_idx=None):
        """rlpyt/agents/base.py:99-116."""
        if cuda_idx is None:
            return
        if self.shared_model is not None:
            self.model = self.ModelCls(**self.env_model_kwargs, **self.model_kwargs)
            self.model.load_state_dict(self.shared_model.state_dict())
        self.device = torch.device("cuda", index=cuda_idx)
        self.model.to(self.device)

    def data_parallel(self):
        """Replicate parameters from rank 0 (what DDP's constructor does, agents/base.py:118-136);
        gradient averaging happens in the optimizer's single flat all-reduce."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            self.world_size = dist.get_world_size()
            for p in self.model.parameters():
                dist.broadcast(p.data, src=0)
            for b in self.model.buffers():
                dist.broadcast(b.data, src=0)
        return self.device.index

    def collector_initialize(self, global_B=1, env_ranks=None):
        pass

    @torch.no_grad()
    def step(self, observation, prev_action, prev_reward):
        raise NotImplementedError

    def reset(self):
        pass

    def reset_one(self, idx):
        pass

    def parameters(self):
        return self.model.parameters()

    def state_dict(self):
        return strip_ddp_state_dict(self.model.state_dict())

    def load_state_dict(self, state_dict):
        self.model.load_state_dict(strip_ddp_state_dict(state_dict))

    def train_mode(self, itr):
        self.model.train()
        self._mode = "train"

    def sample_mode(self, itr):
        self.model.eval()
        self._mode = "sample"

    def eval_mode(self, itr):
        self.model.eval()
        self._mode = "eval"

    def sync_shared_memory(self):
        if self.shared_model is not None and self.shared_model is not self.model:
            self.shared_model.load_state_dict(strip_ddp_state_dict(self.model.state_dict()))

    def toggle_alt(self):
        pass

    # ---- asynchronous mode (agents/base.py:138-160, 218-243) ------------------------------------------------------
    def async_cpu(self, share_memory=True):
        """The reference builds the sampler's separate CPU model here; on this path the sampler acts on the GPU with
        the parameter copy made by ``async_twin`` - nothing to do (kept so reference-style call sequences run)."""

    def _async_tensors(self):
        """The tensors the sampler needs from the optimizer: parameters and buffers of the acting model (agents with
        further acting networks extend this list; target networks are the optimizer's business)."""
        return list(self.model.state_dict().values())

    def async_twin(self):
        """-> the sampler's agent of asynchronous mode: a deep copy of this (optimizer-side) agent with its own
        parameters on the same device, its own distribution / recurrent state / mode flag, linked to this agent by a
        parameter channel: ``self.send_shared_memory()`` copies the trained parameters into a STAGING copy in HBM
        (under the write lock, on the caller's stream), ``twin.recv_shared_memory()`` copies staging -> the twin's
        parameters between batches if something new was sent (under the read lock, on the caller's stream).  Two
        device-to-device copies of a few MB replace the reference's GPU -> shared-memory -> sampler-model round trip;
        the staging copy is what lets the sampler finish a batch on one consistent parameter set while the optimizer
        keeps stepping."""
        import copy
        from rlpyt_b200.utils.synchronize import RWLock, StreamFence
        assert getattr(self, "_async", None) is None, "async_twin() may be called once"
        twin = copy.deepcopy(self)
        channel = dict(staging=[torch.empty_like(t) for t in self._async_tensors()], rw_lock=RWLock(),
                       fence=StreamFence(self.device if self.device.type == "cuda" else None), send_count=0)
        self._async = twin._async = channel
        self._async_role, twin._async_role = "optimizer", "sampler"
        twin._recv_count = 0
        return twin

    def send_shared_memory(self):
        """agents/base.py:218-229 (optimizer side)."""
        ch = getattr(self, "_async", None)
        if ch is None or self._async_role != "optimizer":
            return
        with ch["rw_lock"].write_lock:
            ch["fence"].before_write()
            with torch.no_grad():
                torch._foreach_copy_(ch["staging"], [t.detach() for t in self._async_tensors()])
            ch["fence"].after_write()
            ch["send_count"] += 1

    def recv_shared_memory(self):
        """agents/base.py:231-243 (sampler side)."""
        ch = getattr(self, "_async", None)
        if ch is None or self._async_role != "sampler":
            return
        with ch["rw_lock"]:
            if self._recv_count < ch["send_count"]:
                ch["fence"].before_read()
                with torch.no_grad():
                    torch._foreach_copy_([t.detach() for t in self._async_tensors()], ch["staging"])
                ch["fence"].after_read()
                self._recv_count = ch["send_count"]


class RecurrentAgentMixin:
    """Keeps the recurrent state between ``step`` calls so the sampler stays agnostic (mirror of
    ``rlpyt/agents/base.py:252-306``; use as ``class MyAgent(RecurrentAgentMixin, MyAgentBase)``).  The state is a
    namedarraytuple of ``[N,B,H]`` tensors (cuDNN layout) that lives on the agent's device; ``None`` means zeros."""

    recurrent = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._prev_rnn_state = None
        self._sample_rnn_state = None      # parked while training / evaluating

    def reset(self):
        self._prev_rnn_state = None

    def reset_one(self, idx):
        if self._prev_rnn_state is not None:
            self._prev_rnn_state[:, idx] = 0      # every leaf, column idx

    def advance_rnn_state(self, new_rnn_state):
        """Called by the agent at the end of ``step``."""
        self._prev_rnn_state = new_rnn_state

    @property
    def prev_rnn_state(self):
        return self._prev_rnn_state

    def train_mode(self, itr):
        if self._mode == "sample":
            self._sample_rnn_state = self._prev_rnn_state
        self._prev_rnn_state = None
        super().train_mode(itr)

    def sample_mode(self, itr):
        if self._mode != "sample":
            self._prev_rnn_state = self._sample_rnn_state
        super().sample_mode(itr)

    def eval_mode(self, itr):
        if self._mode == "sample":
            self._sample_rnn_state = self._prev_rnn_state
        self._prev_rnn_state = None
        super().eval_mode(itr)


class AlternatingRecurrentAgentMixin:
    """Two recurrent states, swapped by ``advance_rnn_state`` - for the alternating samplers, where two groups of
    environments take turns stepping (mirror of ``rlpyt/agents/base.py:309-371``)."""

    recurrent = True
    alternating = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._alt = 0
        self._prev_rnn_state = None
        self._prev_rnn_state_pair = [None, None]
        self._sample_rnn_state_pair = [None, None]

    def reset(self):
        self._prev_rnn_state_pair = [None, None]
        self._prev_rnn_state = None
        self._alt = 0

    # NOTE: like the reference's mixin (agents/base.py:309-371) this one has no ``reset_one``: an alternating
    # recurrent agent keeps its state across episode ends inside a batch (BaseAgent.reset_one is a no-op).

    def advance_rnn_state(self, new_rnn_state):
        self._prev_rnn_state_pair[self._alt] = new_rnn_state
        self._alt ^= 1
        self._prev_rnn_state = self._prev_rnn_state_pair[self._alt]

    @property
    def prev_rnn_state(self):
        return self._prev_rnn_state

    def _park(self):
        if self._mode == "sample":
            self._sample_rnn_state_pair = self._prev_rnn_state_pair
        self._prev_rnn_state_pair = [None, None]
        self._prev_rnn_state = None
        self._alt = 0

    def train_mode(self, itr):
        self._park()
        super().train_mode(itr)

    def eval_mode(self, itr):
        self._park()
        super().eval_mode(itr)

    def sample_mode(self, itr):
        if self._mode != "sample":
            self._prev_rnn_state_pair = self._sample_rnn_state_pair
            self._alt = 0
            self._prev_rnn_state = self._prev_rnn_state_pair[0]
        super().sample_mode(itr)

    def get_alt(self):
        return self._alt

    def toggle_alt(self):
        self._alt ^= 1
        self._prev_rnn_state = self._prev_rnn_state_pair[self._alt]
