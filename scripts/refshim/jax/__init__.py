"""Eager torch-float64 stand-in for the parts of ``jax`` the reference's hot path touches (see ../README.md)."""
import torch

from . import lax, nn, numpy, random, scipy, tree, typing  # noqa: F401
from . import core  # noqa: F401

Array = torch.Tensor
__version__ = "0.0-refshim"


def vmap(fn, in_axes=0, out_axes=0):
    """Loop-and-stack ``vmap`` over the mapped positional arguments (dict / tuple outputs supported)."""

    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                leaf = tree.leaves(a)[0]
                n = leaf.shape[ax]
                break
        outs = []
        for i in range(n):
            sl = [a if ax is None else tree.map(lambda t, ax=ax: t.select(ax, i), a) for a, ax in zip(args, axes)]
            outs.append(fn(*sl))
        return tree.map(lambda *xs: torch.stack([torch.as_tensor(x) for x in xs], dim=out_axes), *outs)

    return mapped


def jit(fn, **kw):
    return fn


def pmean(x, axis_name=None):
    return x


def pvary(x, axis_name=None):
    return x


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()
