"""Oracle: the DQN / Double-DQN loss after the network forwards (forward + gradient w.r.t. qs).

Test infrastructure only (see oracle/__init__.py).  torch-CPU fp32 ops in the reference's operation
order - the reference evaluates exactly this with torch on the CPU, after moving both networks'
outputs to the host (rlpyt/agents/dqn/dqn_agent.py:28,75).

Reference: rlpyt/algos/dqn/dqn.py:211-265 (DQN.loss), rlpyt/utils/tensor.py:5-15.
"""
import torch

from oracle.pg_loss import select_at_indexes


def dqn_loss(qs, target_qs, next_qs, action, return_, done_n, is_weights, discount, n_step_return,
             delta_clip=1.0):
    """-> (loss, td_abs_errors, grad_qs).  ``next_qs`` None: plain DQN (max over the target net);
    ``is_weights`` None: uniform replay; ``delta_clip`` None: MSE."""
    qs = qs.clone().requires_grad_(True)
    q = select_at_indexes(action, qs)                                   # dqn.py:231
    with torch.no_grad():
        if next_qs is not None:                                         # :236-239
            next_a = torch.argmax(next_qs, dim=-1)
            target_q = select_at_indexes(next_a, target_qs)
        else:
            target_q = torch.max(target_qs, dim=-1).values              # :241
    disc_target_q = (discount ** n_step_return) * target_q              # :242
    y = return_ + (1 - done_n.float()) * disc_target_q                  # :243
    delta = y - q
    losses = 0.5 * delta ** 2
    abs_delta = abs(delta)
    if delta_clip is not None:                                          # Huber, :247-249
        b = delta_clip * (abs_delta - delta_clip / 2)
        losses = torch.where(abs_delta <= delta_clip, losses, b)
    if is_weights is not None:                                          # :250-251
        losses = losses * is_weights
    td_abs_errors = abs_delta.detach()
    if delta_clip is not None:
        td_abs_errors = torch.clamp(td_abs_errors, 0, delta_clip)       # :254
    loss = torch.mean(losses)                                           # :263 (mid_batch_reset=True)
    loss.backward()
    return loss.detach(), td_abs_errors, qs.grad
