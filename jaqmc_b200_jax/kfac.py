"""``train.optim.module=jaqmc_b200_jax.kfac:kfac``: KFAC over the Flax twin.

``kfac_jax`` registers curvature blocks by pattern-matching ``dot_general (+ bias)`` in the jaxpr of the log-psi function
(reference optimizer/kfac/kfac.py:169-189, tag_registration.py:22-93); an FFI custom call is opaque to it.  The
optimizer therefore differentiates ``wf.twin_logpsi`` (the reference's own graph); sampling and the local energy -- the
hot path -- still run on the kernels.  Adam / SR need no shim (they consume ``LossAndGrad``'s gradients, which come from
the ``custom_vjp`` of ``wavefunction._B200Mixin._op``)."""
from __future__ import annotations

import dataclasses

from jaqmc.optimizer.kfac import kfac as _ref_kfac

__all__ = ["kfac"]


def kfac(*args, **kwargs):
    """The reference's KFAC factory with ``f_log_psi`` redirected to the Flax twin when it is bound to a B200 class."""
    opt = _ref_kfac(*args, **kwargs)
    f = getattr(opt, "f_log_psi", None)
    owner = getattr(f, "__self__", None)
    if owner is not None and hasattr(owner, "twin_logpsi"):
        opt = dataclasses.replace(opt, f_log_psi=owner.twin_logpsi)
    return opt
