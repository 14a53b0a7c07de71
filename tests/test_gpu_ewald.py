"""The Ewald kernel on the GPU against the known answers the reference tests hold (tests/estimator/ewald_test.py:72-152:
Madelung constants of NaCl, primitive and conventional cell, and CaF2, absolute error < 1e-4) and against the oracle on
random systems in cubic / rotated orthorhombic / fcc cells."""

import numpy as np
import pytest
import torch

import test_emu_ewald as E
from jaqmc_b200.ewald import EwaldSum
from oracle import estimators as OE

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _rt():
    from jaqmc_b200._runtime import runtime

    return runtime(DEV)


@pytest.mark.parametrize("name", ["nacl_primitive", "nacl_conventional", "caf2"])
def test_madelung_constants(name):
    E.check_madelung(_rt(), name, device=DEV)


@pytest.mark.parametrize("name", list(E.LATTICES))
def test_ewald_matches_oracle_on_random_systems(name):
    lat = E.LATTICES[name]
    ew = EwaldSum(lat, device=DEV)
    ref_ew = OE.EwaldSum(lat)
    el, at, ch = E._rand_system(lat, 6, 3, seed=3)
    t = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731
    got = ew.energy(t(el), t(at), t(ch), _rt=_rt()).cpu().numpy()
    ref = np.array([OE.solid_potential_energy(ref_ew, el[w].astype(np.float64), at.astype(np.float64), ch) for w in range(5)])
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-5)
