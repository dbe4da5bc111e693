"""Small tensor helpers used by agents/models (mirror of ``rlpyt/utils/tensor.py``)."""
import torch


def select_at_indexes(indexes, tensor):
    """``tensor[..., indexes]`` over the trailing dim (rlpyt/utils/tensor.py:5-15)."""
    lead = indexes.dim()
    assert tuple(indexes.shape) == tuple(tensor.shape[:lead])
    flat = tensor.reshape((indexes.numel(),) + tuple(tensor.shape[lead:]))
    rows = torch.arange(indexes.numel(), device=tensor.device)
    return flat[rows, indexes.reshape(-1).long()].reshape(tuple(tensor.shape[:lead]) + tuple(tensor.shape[lead + 1:]))


def to_onehot(indexes, num, dtype=None):
    """rlpyt/utils/tensor.py:18-27."""
    out = torch.zeros(tuple(indexes.shape) + (num,), dtype=dtype or indexes.dtype, device=indexes.device)
    out.scatter_(-1, indexes.unsqueeze(-1).long(), 1)
    return out


def from_onehot(onehot, dim=-1, dtype=None):
    """rlpyt/utils/tensor.py:30-36."""
    idx = torch.argmax(onehot, dim=dim)
    return idx if dtype is None else idx.type(dtype)


def valid_mean(tensor, valid=None, dim=None):
    """rlpyt/utils/tensor.py:39-46."""
    dim = () if dim is None else dim
    if valid is None:
        return tensor.mean(dim=dim)
    valid = valid.type(tensor.dtype)
    return (tensor * valid).sum(dim=dim) / valid.sum(dim=dim)


def infer_leading_dims(tensor, dim):
    """(lead_dim, T, B, data_shape) for [T,B,*], [B,*] or [*] inputs (tensor.py:49-68)."""
    lead_dim = tensor.dim() - dim
    assert lead_dim in (0, 1, 2)
    if lead_dim == 2:
        T, B = tensor.shape[:2]
    else:
        T, B = 1, (1 if lead_dim == 0 else tensor.shape[0])
    return lead_dim, T, B, tensor.shape[lead_dim:]


def restore_leading_dims(tensors, lead_dim, T=1, B=1):
    """Inverse of ``infer_leading_dims`` on model outputs shaped [T*B, ...] (tensor.py:71-86)."""
    many = isinstance(tensors, (tuple, list))
    ts = tuple(tensors) if many else (tensors,)
    if lead_dim == 2:
        ts = tuple(t.view((T, B) + tuple(t.shape[1:])) for t in ts)
    elif lead_dim == 0:
        assert B == 1
        ts = tuple(t.squeeze(0) for t in ts)
    return ts if many else ts[0]
