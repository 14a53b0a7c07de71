"""Walker data containers (host-side mirror of the reference's ``data.py`` / ``app/molecule/data.py``).

The reference's ``Data`` is a JAX pytree dataclass and ``BatchedData`` records which fields carry the walker
axis (``data.py:159-205``).  Here the leaves are ``torch`` CUDA tensors; only ``electrons`` is batched
(``app/molecule/data.py:46-53``), ``atoms`` / ``charges`` are replicated.
"""

from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field

import torch


@dataclass
class MoleculeData:
    """``electrons`` (W, n, 3) [or (n, 3) for one walker], ``atoms`` (A, 3), ``charges`` (A,)."""

    electrons: torch.Tensor
    atoms: torch.Tensor
    charges: torch.Tensor

    def merge(self, updates: dict) -> "MoleculeData":
        """``Data.merge`` (data.py): a copy with some fields replaced."""
        return dataclasses.replace(self, **updates)


@dataclass
class BatchedData:
    """``data`` plus the names of the fields whose axis 0 is the walker axis (reference data.py:159-179)."""

    data: MoleculeData
    fields_with_batch: list = field(default_factory=lambda: ["electrons"])

    @property
    def batch_size(self) -> int:
        return int(getattr(self.data, self.fields_with_batch[0]).shape[0])

    def shard(self, rank: int, world_size: int) -> "BatchedData":
        """Contiguous block of walkers for ``rank`` (the reference's 1-D mesh sharding, data.py:181-205)."""
        upd = {}
        for name in self.fields_with_batch:
            t = getattr(self.data, name)
            if t.shape[0] % world_size != 0:
                raise ValueError(f"batch size {t.shape[0]} is not divisible by the number of devices {world_size}")
            per = t.shape[0] // world_size
            upd[name] = t[rank * per:(rank + 1) * per].contiguous()
        return BatchedData(self.data.merge(upd), list(self.fields_with_batch))


@dataclass
class SolidData:
    """``electrons`` (W, n, 3), ``atoms`` (A_cell, 3) / ``charges`` (A_cell,) of the simulation cell,
    ``primitive_atoms`` (A_prim, 3) -- reference app/solid/data.py:52-79."""

    electrons: torch.Tensor
    atoms: torch.Tensor
    charges: torch.Tensor
    primitive_atoms: torch.Tensor

    def merge(self, updates: dict) -> "SolidData":
        return dataclasses.replace(self, **updates)
