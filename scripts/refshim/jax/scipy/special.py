import torch


def logsumexp(a, axis=None, b=None, return_sign=False):
    """``log|sum b exp(a)|`` (and its sign) with the max shift, as jax.scipy.special.logsumexp."""
    if b is None:
        b = torch.ones_like(a)
    m = a.real.max() if axis is None else a.real.amax(dim=axis, keepdim=True)
    s = (b * torch.exp(a - m)).sum() if axis is None else (b * torch.exp(a - m)).sum(dim=axis)
    mm = m if axis is None else m.squeeze(axis)
    out = torch.log(torch.abs(s)) + mm
    if return_sign:
        sign = torch.sign(s) if not s.is_complex() else s / torch.abs(s)
        return out, sign
    return out
