"""Agent base class: glue between sampler, network and algorithm (host-side mirror of
``rlpyt/agents/base.py:17-246``; synchronous path only - the async/alternating machinery of the
reference is out of scope).

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
