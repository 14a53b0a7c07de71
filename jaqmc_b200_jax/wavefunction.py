"""``wf.module=jaqmc_b200_jax.wavefunction:<Class>``: subclasses of the reference's wavefunction classes (same fields,
same Flax parameter tree, so checkpoints / pretraining / KFAC keep working) whose ``logpsi`` / ``phase_logpsi`` /
``evaluate`` / ``orbitals`` run on the B200 kernels.  See the package docstring for the layering and STATUS."""
from __future__ import annotations

import functools

import jax
import jax.numpy as jnp

from jaqmc.app.molecule.wavefunction.ferminet import FermiNetWavefunction as _RefFermiNet
from jaqmc.app.molecule.wavefunction.lapnet import LapNetWavefunction as _RefLapNet
from jaqmc.app.molecule.wavefunction.psiformer import PsiformerWavefunction as _RefPsiformer
from jaqmc.laplacian import AutoLaplacianFallback, LapTuple, custom_laplacian, is_local1_laptuple

from . import _ffi
from ._config import pack_config

__all__ = ["FermiNetWavefunction", "LapNetWavefunction", "PsiformerWavefunction"]


def _identity_local1_seed(x) -> bool:
    """True when ``x`` is the LapTuple ``EuclideanKinetic`` plants: electron i owns its own 3 coordinates, unit Jacobian,
    zero Laplacian (laplacian/seed.py:75-121).  The FFI rule is only valid for that seed."""
    if not (isinstance(x, LapTuple) and is_local1_laptuple(x)):
        return False
    jac = x.jacobian
    n = x.x.shape[0]
    owner = jac.owners[0]
    return bool(jac.input_n_particles == n and owner.axis == 0 and (owner.values == jnp.arange(n)).all())


class _B200Mixin:
    """Shared implementation; ``_kind`` and ``_config_fields`` are set by the concrete classes."""

    _kind: str = ""

    # -- configuration -> FFI attributes -------------------------------------------------------------------------
    def _packed(self, n_atoms: int):
        fields = {f: getattr(self, f) for f in self.__dataclass_fields__ if f not in ("parent", "name")}
        fields["envelope"] = str(getattr(self, "envelope").value if hasattr(getattr(self, "envelope"), "value")
                                 else getattr(self, "envelope"))
        for k in ("layer_norm_mode", "jastrow"):
            if k in fields and hasattr(fields[k], "value"):
                fields[k] = fields[k].value
        return pack_config(self._kind, n_atoms=n_atoms, **fields)

    # -- the Flax twin (reference graph): derivative rules and KFAC use it --------------------------------------------
    def twin_logpsi(self, params, data):
        return super().logpsi(params, data)

    # -- kernels ---------------------------------------------------------------------------------------------------
    def _op(self, n_atoms: int):
        """``f(params, electrons, atoms, charges) -> logpsi`` with the three transform rules attached; cached per
        atom count (the reference builds one wavefunction per system)."""
        cache = self.__dict__.setdefault("_b200_ops", {})
        if n_atoms in cache:
            return cache[n_atoms]
        packed = self._packed(n_atoms)
        twin = lambda params, el, atoms, charges: type(self).twin_logpsi(  # noqa: E731
            self, params, self._data_like(el, atoms, charges))

        @jax.custom_vjp
        def value(params, el, atoms, charges):
            return _ffi.logpsi(packed, params, el, atoms)[0]

        def fwd(params, el, atoms, charges):
            return value(params, el, atoms, charges), (params, el, atoms, charges)

        def bwd(res, ct):
            # derivative of the reference graph; evaluated only under LossAndGrad / SR (one walker, vmapped)
            _, vjp = jax.vjp(twin, *res)
            return vjp(ct)

        value.defvjp(fwd, bwd)

        @custom_laplacian
        def op(params, el, atoms, charges):
            return value(params, el, atoms, charges)

        @op.def_laplacian_rule
        def rule(params, el, atoms, charges):
            if any(isinstance(a, LapTuple) for a in jax.tree.leaves((params, atoms, charges), is_leaf=lambda a: isinstance(a, LapTuple))):
                raise AutoLaplacianFallback("only the electron positions may be tracked")
            if not _identity_local1_seed(el):
                raise AutoLaplacianFallback("not the identity Local1 seed")
            out = _ffi.local_energy(packed, params, el.x, atoms, charges)
            return LapTuple(out["logpsi"], out["grad"], out["lap"])

        cache[n_atoms] = op
        return op

    def _data_like(self, electrons, atoms, charges):
        from jaqmc.app.molecule.data import MoleculeData

        return MoleculeData(electrons=electrons, atoms=atoms, charges=charges)

    # -- MoleculeWavefunction protocol (app/molecule/wavefunction/base.py:17-72) ---------------------------------------
    def logpsi(self, params, data):
        return self._op(data.atoms.shape[0])(params, data.electrons, data.atoms, data.charges)

    def phase_logpsi(self, params, data):
        lp, sign = _ffi.logpsi(self._packed(data.atoms.shape[0]), params, data.electrons, data.atoms)
        return sign, lp

    def evaluate(self, params, data):
        lp, sign = _ffi.logpsi(self._packed(data.atoms.shape[0]), params, data.electrons, data.atoms)
        return {"logpsi": lp, "sign_logpsi": sign}

    def orbitals(self, params, data):
        return _ffi.orbitals(self._packed(data.atoms.shape[0]), params, data.electrons, data.atoms, self.ndets)


class FermiNetWavefunction(_B200Mixin, _RefFermiNet):
    _kind = "ferminet"


class LapNetWavefunction(_B200Mixin, _RefLapNet):
    _kind = "lapnet"


class PsiformerWavefunction(_B200Mixin, _RefPsiformer):
    _kind = "psiformer"
