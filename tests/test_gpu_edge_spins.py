"""GPU counterpart of test_emu_edge_spins.py: single-channel spin configurations (1, 0), (0, 2), (3, 0), (0, 1)
(reference tests/wavefunction/molecule_wavefunction_test.py:148-190) on the CUDA kernels."""

import pytest
import torch

import test_emu_edge_spins as E

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["ferminet", "psiformer", "lapnet"])
@pytest.mark.parametrize("nspins,charge", E.CASES, ids=["%d_%d" % c[0] for c in E.CASES])
def test_single_channel_spin_configurations(nspins, charge, kind):
    from jaqmc_b200._runtime import runtime

    E.check(runtime(torch.device("cuda", 0)), nspins, charge, kind, device=torch.device("cuda", 0))
