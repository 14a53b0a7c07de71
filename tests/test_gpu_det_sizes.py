"""GPU counterpart of test_emu_det_sizes.py: one network-level parity case per determinant-kernel selection."""

import pytest
import torch

import test_emu_det_sizes as D
import test_emu_edge_spins as E

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nspins,charge,ndets", D.CASES, ids=D.IDS)
def test_determinant_sizes(nspins, charge, ndets):
    from jaqmc_b200._runtime import runtime

    E.check(runtime(torch.device("cuda", 0)), nspins, charge, "ferminet", device=torch.device("cuda", 0), ndets=ndets)
