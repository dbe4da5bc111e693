"""Proximal Policy Optimisation on device-resident samples (mirror of
``rlpyt/algos/pg/ppo.py:16-154``: same constructor, ``initialize``, ``optimize_agent``, ``loss``).

Where the time went in the reference and what replaces it:
* ``process_returns``: T-step Python loop on torch-CPU tensors -> GAE scan kernel;
* ``loss_inputs[T_idxs, B_idxs]`` (:99-100): a fancy-index per field, observations first copied
  to the device (:72) -> one 16-byte-vectorised row-gather kernel for the observations
  (231 MB/minibatch) + one multi-field gather for the scalars, all from the resident [T,B] buffers;
* loss arithmetic on CPU with D2H of the network outputs -> one fused forward+gradient kernel;
* DDP all-reduce + clip_grad_norm_ + Adam -> ``FlatAdam.clip_and_step`` (1 NCCL call, 2 kernels);
* 4 ``.item()`` host syncs per update (:106-109) -> one D2H of all OptInfo rows per iteration.
Recurrent agents (ppo.py:84-100, 127-131) train on whole trajectories, minibatches over B: ``_optimize_recurrent``.
"""
import numpy as np
import torch

from rlpyt_b200 import _lib
from rlpyt_b200.agents.base import AgentInputs
from rlpyt_b200.algos.optim import FlatAdam
from rlpyt_b200.algos.pg import loss_ops
from rlpyt_b200.algos.pg.base import PolicyGradientAlgo, OptInfo
from rlpyt_b200.utils.buffer import buffer_method, buffer_to
from rlpyt_b200.utils.collections import namedarraytuple
from rlpyt_b200.utils.gather import gather_rows, gather_rows_multi, LazyRows
from rlpyt_b200.utils.misc import iterate_mb_idxs

LossInputs = namedarraytuple("LossInputs",
                             ["agent_inputs", "action", "return_", "advantage", "valid", "old_dist_info"])


class PPO(PolicyGradientAlgo):

    def __init__(self, discount=0.99, learning_rate=0.001, value_loss_coeff=1., entropy_loss_coeff=0.01,
                 OptimCls=FlatAdam, optim_kwargs=None, clip_grad_norm=1., initial_optim_state_dict=None,
                 gae_lambda=1, minibatches=4, epochs=4, ratio_clip=0.1, linear_lr_schedule=True,
                 normalize_advantage=False):
        self.discount = discount
        self.learning_rate = learning_rate
        self.value_loss_coeff = value_loss_coeff
        self.entropy_loss_coeff = entropy_loss_coeff
        self.OptimCls = OptimCls
        self.optim_kwargs = dict() if optim_kwargs is None else optim_kwargs
        self.clip_grad_norm = clip_grad_norm
        self.initial_optim_state_dict = initial_optim_state_dict
        self.gae_lambda = gae_lambda
        self.minibatches = minibatches
        self.epochs = epochs
        self.ratio_clip = ratio_clip
        self.linear_lr_schedule = linear_lr_schedule
        self.normalize_advantage = normalize_advantage

    def initialize(self, *args, **kwargs):
        """ppo.py:46-57."""
        super().initialize(*args, **kwargs)
        self._batch_size = self.batch_spec.size // self.minibatches  # for logging
        if self.linear_lr_schedule:
            self.lr_scheduler = torch.optim.lr_scheduler.LambdaLR(
                optimizer=self.optimizer, lr_lambda=lambda itr: (self.n_itr - itr) / self.n_itr)
            self._ratio_clip = self.ratio_clip

    def optimize_agent(self, itr, samples):
        """ppo.py:59-115.  Feed-forward agents: shuffled [T*B] minibatches through the fused row gathers below;
        recurrent agents: whole trajectories, minibatches over B only (``_optimize_recurrent``)."""
        if self.agent.recurrent:
            return self._optimize_recurrent(itr, samples)
        dev = self._device()
        obs = self._on_device(samples.env.observation)               # ppo.py:67-72 (no-op when resident)
        prev_action = self._on_device(samples.agent.prev_action)
        prev_reward = self._on_device(samples.env.prev_reward)
        if hasattr(self.agent, "update_obs_rms"):
            self.agent.update_obs_rms(obs)
        prof = getattr(self, "profile_events", None)   # bench.py: device time of the returns + loss kernels inside a real iteration
        ev = (lambda: None) if prof is None else (lambda: prof.append(_recorded_event()))
        ev()
        return_, advantage, valid = self.process_returns(samples)     # ppo.py:75
        ev()
        action = self._on_device(samples.agent.action)
        old_prob = self._on_device(samples.agent.agent_info.dist_info.prob)
        T, B = samples.env.reward.shape[:2]
        batch_size = T * B
        flat = lambda x: x.reshape((batch_size,) + tuple(x.shape[2:]))
        obs_f = flat(obs)
        small = [flat(prev_action), flat(prev_reward), flat(action), flat(return_), flat(advantage), flat(old_prob)]
        if valid is not None:
            small.append(flat(valid))
        mb_size = batch_size // self.minibatches                       # ppo.py:90-91
        n_updates = self.epochs * (batch_size // mb_size)
        stats = torch.zeros((n_updates, 9), dtype=torch.float32, device=dev)   # per update: the loss kernel's 8 scalars + gradNorm
        fused_opt = isinstance(self.optimizer, FlatAdam)
        lazy_obs = bool(getattr(self.agent.model, "accepts_lazy_rows", False)) and obs_f.dtype == torch.uint8
        # every minibatch's rows of this iteration in one upload; the shuffles are drawn in the reference's order (ppo.py:92-95)
        rows_np = np.stack([(idxs % T) * B + (idxs // T) for _ in range(self.epochs)
                            for idxs in iterate_mb_idxs(batch_size, mb_size, shuffle=True)])
        rows_all = torch.from_numpy(rows_np).to(dev, non_blocking=True)
        mbg = self._minibatch_graph(obs_f, small, valid is not None, mb_size, lazy_obs) if (fused_opt and prof is None) else None
        if mbg is not None:
            mbg.clip.fill_(float(self.ratio_clip))
        for u in range(n_updates):
            if mbg is not None:
                # gather -> forward -> fused loss -> backward into the flat gradient buffer: ONE graph launch
                mbg.rows.copy_(rows_all[u], non_blocking=True)
                mbg.graph.replay()
                _lib.launch_count += mbg.n_launch                      # this package's kernels inside the graph (bench.py's count)
                sc = mbg.sc
            else:
                rows = rows_all[u]
                self.optimizer.zero_grad()                             # ppo.py:96
                sc = self._minibatch_forward_backward(obs_f, small, valid is not None, rows, lazy_obs, self.ratio_clip, ev)
            if fused_opt:
                self.optimizer.clip_and_step(self.clip_grad_norm, out=stats[u, 8:9])   # [all-reduce ->] clip + Adam
            else:
                stats[u, 8] = torch.nn.utils.clip_grad_norm_(self.agent.parameters(), self.clip_grad_norm).reshape(())
                self.optimizer.step()
            stats[u, :8].copy_(sc, non_blocking=True)
            self.update_counter += 1
        self._ff_iterations = getattr(self, "_ff_iterations", 0) + 1
        if self.linear_lr_schedule:                                    # ppo.py:110-113
            self.lr_scheduler.step()
            self.ratio_clip = self._ratio_clip * (self.n_itr - itr) / self.n_itr
        host = stats.cpu().numpy().astype(np.float64)                  # the iteration's single D2H sync
        return OptInfo(loss=host[:, 0].tolist(), gradNorm=host[:, 8].tolist(),
                       entropy=host[:, 1].tolist(), perplexity=host[:, 2].tolist())

    def _minibatch_forward_backward(self, obs_f, small, has_valid, rows, lazy_obs, ratio_clip, ev=lambda: None):
        """One minibatch of ppo.py:94-101 up to ``loss.backward()``: row gathers from the resident [T*B] buffers,
        forward, fused loss (+ its gradient), backward into the flat gradient buffer.  Returns the loss kernel's
        8 scalars (device)."""
        # models that read rows in place (fused first layer) get the un-gathered view
        obs_mb = LazyRows(obs_f, rows) if lazy_obs else gather_rows(obs_f, rows)
        got = gather_rows_multi(small, rows)
        pa_mb, pr_mb, act_mb, ret_mb, adv_mb, oldp_mb = got[:6]
        valid_mb = got[6] if has_valid else None
        dist_info, value = self.agent(obs_mb, pa_mb, pr_mb)            # ppo.py:133
        ev()
        loss, sc = loss_ops.ppo_loss(dist_info.prob, value, oldp_mb, act_mb, ret_mb, adv_mb, valid_mb,
                                     ratio_clip, self.value_loss_coeff, self.entropy_loss_coeff)
        ev()
        loss.backward()                                                # ppo.py:101
        return sc

    MAX_MINIBATCH_GRAPHS = 2

    def _minibatch_graph(self, obs_f, small, has_valid, mb_size, lazy_obs):
        """The minibatch body above as ONE CUDA graph (ppo.py:94-101 is ~110 kernel launches of 3-250 us each; issued
        from Python + autograd they leave the GPU idle for 10-15 % of the iteration).  The graph reads the sampler's
        resident [T*B] buffers (fixed addresses: the sampler returns the same buffers every iteration), a static row
        index vector and the ratio clip as a device scalar (it follows the linear schedule), and leaves the gradient in
        the flat buffer; all-reduce and clip + Adam stay outside (their host scalars - learning rate, step count -
        change per update).  Captured after one eager iteration (lazy initialisation must not happen under capture);
        buffers at new addresses get a new graph, more than ``MAX_MINIBATCH_GRAPHS`` of them switch graphs off (a
        caller that hands in fresh tensors every iteration would otherwise re-capture forever).
        ``RLPYT_B200_LEARNER_GRAPH=0`` disables."""
        import os
        if os.environ.get("RLPYT_B200_LEARNER_GRAPH", "1") != "1" or getattr(self, "_ff_iterations", 0) < 1:
            return None
        graphs = self.__dict__.setdefault("_mb_graphs", {})
        if graphs is None:
            return None
        key = (obs_f.data_ptr(), tuple(obs_f.shape), tuple((t.data_ptr(), tuple(t.shape)) for t in small), has_valid,
               mb_size, lazy_obs, tuple(p.data_ptr() for p in self.agent.parameters()))
        mbg = graphs.get(key)
        if mbg is not None:
            return mbg
        if len(graphs) >= self.MAX_MINIBATCH_GRAPHS:
            self._mb_graphs = None                                     # thrashing: stay eager from now on
            return None
        from types import SimpleNamespace
        dev = obs_f.device
        mbg = SimpleNamespace(rows=torch.zeros(mb_size, dtype=torch.int64, device=dev),
                              clip=torch.full((1,), float(self.ratio_clip), dtype=torch.float32, device=dev),
                              graph=torch.cuda.CUDAGraph(), keep=(obs_f, small))
        torch.cuda.synchronize(dev)
        n0 = _lib.launch_count
        issuing = torch.cuda.current_stream(dev)
        try:
            with torch.cuda.graph(mbg.graph):
                self.optimizer.zero_grad()
                mbg.sc = self._minibatch_forward_backward(obs_f, small, has_valid, mbg.rows, lazy_obs, mbg.clip)
        except RuntimeError as e:
            # an invalidated capture (a CUDA call from outside the body, a library allocating on first use) has executed
            # nothing on the device: drop what the host side of the body attached and issue minibatches eagerly from now on
            import warnings
            warnings.warn(f"rlpyt_b200: CUDA-graph capture of the PPO minibatch failed ({str(e).splitlines()[0][:120]}); "
                          "minibatches are issued eagerly from now on")
            _lib.launch_count = n0
            self._mb_graphs = None
            torch.cuda.set_stream(issuing)                   # torch.cuda.graph.__exit__ raised before it restored the stream
            torch.cuda.synchronize(dev)
            self.optimizer.zero_grad()
            return None
        mbg.n_launch, _lib.launch_count = _lib.launch_count - n0, n0   # capturing launched nothing
        graphs[key] = mbg
        return mbg

    def _optimize_recurrent(self, itr, samples):
        """The recurrent branch of ppo.py:59-115 (:84-86 ``init_rnn_state = prev_rnn_state[0]`` kept ``[B,N,H]`` for
        slicing, :89-97 minibatches of whole trajectories over B with ``T_idxs = slice(None)``, :127-131 transpose to
        ``[N,B,H]`` for cuDNN).  ``valid`` masks the steps after an episode end inside the batch (pg/base.py:60-63);
        the loss arithmetic is the same fused kernel, over [T*mb] rows."""
        dev = self._device()
        obs = self._on_device(samples.env.observation)
        prev_action = self._on_device(samples.agent.prev_action)
        prev_reward = self._on_device(samples.env.prev_reward)
        return_, advantage, valid = self.process_returns(samples)
        action = self._on_device(samples.agent.action)
        old_prob = self._on_device(samples.agent.agent_info.dist_info.prob)
        init_rnn_state = buffer_to(samples.agent.agent_info.prev_rnn_state[0], device=dev)     # T = 0: [B,N,H]
        T, B = samples.env.reward.shape[:2]
        mb_size = B // self.minibatches
        n_updates = self.epochs * (B // mb_size)
        stats = torch.zeros((n_updates, 4), dtype=torch.float32, device=dev)
        fused_opt = isinstance(self.optimizer, FlatAdam)
        u = 0
        for _ in range(self.epochs):
            for idxs in iterate_mb_idxs(B, mb_size, shuffle=True):
                cols = torch.from_numpy(np.ascontiguousarray(idxs)).to(dev, non_blocking=True)
                sel = lambda x: x.index_select(1, cols)
                self.optimizer.zero_grad()
                rnn_state = buffer_method(buffer_method(init_rnn_state[cols], "transpose", 0, 1), "contiguous")
                dist_info, value, _next = self.agent(sel(obs).contiguous(), sel(prev_action), sel(prev_reward), rnn_state)
                loss, sc = loss_ops.ppo_loss(dist_info.prob, value, sel(old_prob), sel(action), sel(return_), sel(advantage),
                                             None if valid is None else sel(valid), self.ratio_clip,
                                             self.value_loss_coeff, self.entropy_loss_coeff)
                loss.backward()
                if fused_opt:
                    grad_norm = self.optimizer.clip_and_step(self.clip_grad_norm)
                else:
                    grad_norm = torch.nn.utils.clip_grad_norm_(self.agent.parameters(), self.clip_grad_norm)
                    self.optimizer.step()
                stats[u, 0] = sc[0]
                stats[u, 1] = grad_norm.reshape(())
                stats[u, 2:4] = sc[1:3]
                u += 1
                self.update_counter += 1
        if self.linear_lr_schedule:
            self.lr_scheduler.step()
            self.ratio_clip = self._ratio_clip * (self.n_itr - itr) / self.n_itr
        host = stats[:u].cpu().numpy().astype(np.float64)
        return OptInfo(loss=host[:, 0].tolist(), gradNorm=host[:, 1].tolist(),
                       entropy=host[:, 2].tolist(), perplexity=host[:, 3].tolist())

    def loss(self, agent_inputs, action, return_, advantage, valid, old_dist_info, init_rnn_state=None):
        """ppo.py:117-154 signature; returns device 0-dim tensors ``(loss, entropy, perplexity)``."""
        if init_rnn_state is not None:                                    # [B,N,H] -> [N,B,H] (ppo.py:127-131)
            init_rnn_state = buffer_method(buffer_method(init_rnn_state, "transpose", 0, 1), "contiguous")
            dist_info, value, _rnn = self.agent(*agent_inputs, init_rnn_state)
            loss, sc = loss_ops.ppo_loss(dist_info.prob, value, old_dist_info.prob, action, return_, advantage,
                                         valid, self.ratio_clip, self.value_loss_coeff, self.entropy_loss_coeff)
            return loss, sc[1], sc[2]
        dist_info, value = self.agent(*agent_inputs)
        loss, sc = loss_ops.ppo_loss(dist_info.prob, value, old_dist_info.prob, action, return_, advantage,
                                     valid, self.ratio_clip, self.value_loss_coeff, self.entropy_loss_coeff)
        return loss, sc[1], sc[2]


def _recorded_event():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e
