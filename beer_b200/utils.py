"""Helpers with the semantics of beer/utils.py:84-123 (`onehot`, `logsumexp`); index / reduction
plumbing on tensors, not part of the kernel path."""
import torch

__all__ = ['onehot', 'logsumexp']


def onehot(labels, max_label, dtype, device):
    """One-hot encoding [len(labels), max_label] (utils.py:84-102)."""
    labels = torch.as_tensor(labels, device=device).long()
    retval = torch.zeros(len(labels), max_label, dtype=dtype, device=device)
    idxs = torch.arange(len(labels), device=device) * max_label + labels
    retval.view(-1)[idxs] = 1
    return retval


def logsumexp(tensor, dim=0):
    """log-sum-exp that returns +-inf when the maximum is +-inf (utils.py:105-123)."""
    tmax, _ = torch.max(tensor, dim=dim, keepdim=True)
    inf = (tmax == float('-inf')) | (tmax == float('inf'))
    safe = torch.where(inf, torch.zeros_like(tmax), tmax)
    out = safe + (tensor - safe).exp().sum(dim=dim, keepdim=True).log()
    return torch.where(inf, tmax, out).squeeze(dim)
