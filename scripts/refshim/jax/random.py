"""``jax.random`` stand-in.  Keys are opaque; ``normal`` / ``uniform`` return the arrays queued by the fixture script
(``queue_normal`` / ``queue_uniform``) so that the reference's sampler consumes exactly the noise the kernels get.
With nothing queued they fall back to a seeded torch generator (parameter initialisers)."""
import itertools

import torch

_counter = itertools.count()
_normals: list = []
_uniforms: list = []
_gen = torch.Generator().manual_seed(1234)


class _Key:
    def __init__(self, tag):
        self.tag = tag

    def __repr__(self):
        return f"Key({self.tag})"


def PRNGKey(seed):
    _gen.manual_seed(int(seed))
    return _Key(("root", int(seed)))


key = PRNGKey


def split(k, num=2):
    return [_Key((getattr(k, "tag", k), next(_counter))) for _ in range(num)]


def fold_in(k, data):
    return _Key((getattr(k, "tag", k), "fold", int(data)))


def queue_normal(t):
    _normals.append(t)


def queue_uniform(t):
    _uniforms.append(t)


def normal(k, shape=(), dtype=None):
    if _normals:
        t = _normals.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t
    return torch.randn(tuple(shape), generator=_gen, dtype=torch.float64)


def uniform(k, shape=(), dtype=None, minval=0.0, maxval=1.0):
    if _uniforms:
        t = _uniforms.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t
    return minval + (maxval - minval) * torch.rand(tuple(shape), generator=_gen, dtype=torch.float64)
