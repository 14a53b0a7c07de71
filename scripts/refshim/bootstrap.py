"""Makes ``import jaqmc.<hot-path module>`` work from /root/reference/src without jax / flax / pyserde / pyscf.

``install()`` puts the stand-ins of this directory first on ``sys.path``, registers namespace stubs for the reference
packages whose ``__init__`` imports the workflow / optimizer / SCF stack, and by-name stubs for the few support modules
the hot path imports but does not compute with.  Test infrastructure (fixture generation only)."""
from __future__ import annotations

import dataclasses
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("JAQMC_REFERENCE_SRC", "/root/reference/src")


def _pkg_stub(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


def install():
    if "jaqmc" in sys.modules and getattr(sys.modules["jaqmc"], "_refshim", False):
        return
    if not os.path.isdir(os.path.join(REF_SRC, "jaqmc")):
        raise RuntimeError(f"reference sources not found under {REF_SRC}")
    sys.path.insert(0, HERE)          # jax, flax, serde stand-ins
    import jax  # noqa: F401  (the stand-in)
    import torch

    torch.set_default_dtype(torch.float64)
    J = os.path.join(REF_SRC, "jaqmc")
    root = _pkg_stub("jaqmc", J)
    root._refshim = True
    # packages entered without running their __init__ (they import workflow / optimizers / pyscf)
    for sub in ("app", "app/molecule", "app/solid", "estimator", "estimator/kinetic", "estimator/ecp", "sampler", "utils",
                "utils/atomic", "geometry", "optimizer", "workflow"):
        _pkg_stub("jaqmc." + sub.replace("/", "."), os.path.join(J, sub))

    # ---- by-name stubs ------------------------------------------------------------------------------------------
    class Data:   # jaqmc/data.py: a dataclass pytree; the path only reads fields and calls merge
        def merge(self, updates):
            return dataclasses.replace(self, **updates)

        def __getitem__(self, key):
            return getattr(self, key)

    class BatchedData:
        pass

    _mod("jaqmc.data", Data=Data, BatchedData=BatchedData)

    @dataclasses.dataclass
    class MoleculeData(Data):   # app/molecule/data.py:46-53
        electrons: object = None
        atoms: object = None
        charges: object = None

    _mod("jaqmc.app.molecule.data", MoleculeData=MoleculeData)

    @dataclasses.dataclass
    class SolidData(Data):      # app/solid/data.py:52-79
        electrons: object = None
        atoms: object = None
        charges: object = None
        primitive_atoms: object = None

    _mod("jaqmc.app.solid.data", SolidData=SolidData)

    def configurable_dataclass(cls=None, **kw):
        def deco(c):
            return dataclasses.dataclass(c, kw_only=True) if not dataclasses.is_dataclass(c) else c

        return deco(cls) if cls is not None else deco

    _mod("jaqmc.utils.config", configurable_dataclass=configurable_dataclass, ConfigManager=object)
    _mod("jaqmc.utils.parallel_jax", BATCH_AXIS_NAME="batch", pmean=lambda x, *a, **k: x, pvary=lambda x, *a, **k: x,
         process_allgather=lambda x, *a, **k: x)

    # jaqmc.laplacian: only the decorator surface wavefunction/backbone/lapnet/_attention.py touches at import time
    class _CustomLaplacian:
        def __init__(self, fn):
            self.fn = fn
            self.__name__ = getattr(fn, "__name__", "custom")

        def __call__(self, *a, **k):
            return self.fn(*a, **k)

        def def_laplacian_rule(self, rule):
            return rule

    class _Dummy:
        def __class_getitem__(cls, item):
            return cls

    _mod("jaqmc.laplacian", custom_laplacian=_CustomLaplacian, LapTuple=_Dummy, Local1Jacobian=_Dummy,
         ArrayOrLapTuple=object, AutoLaplacianFallback=type("AutoLaplacianFallback", (Exception,), {}),
         is_local1_laptuple=lambda x: False)

    # estimator base classes used as bases by hamiltonian / total-energy estimators
    class PerWalkerEstimator:
        def __class_getitem__(cls, item):
            return cls

    _mod("jaqmc.estimator.base", PerWalkerEstimator=PerWalkerEstimator, mean_reduce=None)
    sys.modules["jaqmc.estimator"].PerWalkerEstimator = PerWalkerEstimator
    sys.modules["jaqmc.estimator"].EstimatorLike = object
    sys.path.insert(1, REF_SRC)


def ref(name):
    """Import a reference module by dotted name (after ``install()``)."""
    install()
    return importlib.import_module(name)
