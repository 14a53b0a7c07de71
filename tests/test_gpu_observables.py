"""GPU tests of the psi-ratio consumers through the host classes: ``SpinSquared`` (reference
estimator/spin.py) and the non-local ECP integral (estimator/ecp/nonlocal_integral.py) against the golden vectors the
reference's own code produced (tests/golden/ref_observables_*.npz), plus size-independent properties."""

import numpy as np
import pytest
import torch

import helpers as H
import test_reference_fixtures as R
from jaqmc_b200.data import MoleculeData

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


@pytest.mark.parametrize("name", ["observables_lih", "observables_li"])
def test_spin_squared_and_nonlocal_integral_match_reference(name):
    from jaqmc_b200.ecp import get_quadrature, make_nonlocal_integral
    from jaqmc_b200.spin import SpinSquared
    from jaqmc_b200.wavefunction import FermiNetWavefunction

    meta, p64, z = R.load(name)
    nspins = tuple(meta["nspins"])
    wf = FermiNetWavefunction(nspins=nspins, **meta["kwargs"])
    f32 = lambda k: torch.from_numpy(z[k].astype(np.float32)).to(DEV)  # noqa: E731
    data = MoleculeData(f32("electrons"), f32("atoms"), f32("charges"))
    params = H.to_f32(p64, DEV)
    spin = SpinSquared(n_up=nspins[0], n_down=nspins[1], phase_logpsi=wf)
    spin.init(data)
    s2 = spin.evaluate_batch_walkers(params, data)[0]["spin:s2"].cpu().numpy()
    np.testing.assert_allclose(s2, z["s2"], rtol=2e-4, atol=2e-4)
    W, n = z["electrons"].shape[:2]
    evaluate = make_nonlocal_integral(meta["num_channels"], get_quadrature(meta["quadrature"]))
    atom_pos = data.atoms[None, None].expand(W, n, -1, -1).contiguous()
    got = evaluate(wf, params, data, atom_pos, (f32("u1"), f32("u2"))).cpu().numpy()
    scale = np.abs(z["nonlocal_integrals"]).max()
    assert np.abs(got - z["nonlocal_integrals"]).max() < 3e-4 * scale


def test_spin_squared_properties_full_batch():
    """4096 walkers of Li (S_z = 1/2): finite values; a closed-shell single-determinant-like check is not available for a
    random network, but S^2 must be invariant under relabelling the majority electrons."""
    from jaqmc_b200.spin import SpinSquared
    from jaqmc_b200.wavefunction import FermiNetWavefunction

    atoms, charges, nspins = H.molecule("Li")
    wf = FermiNetWavefunction(nspins=nspins, ndets=4, hidden_dims_single=[64, 64], hidden_dims_double=[16, 16])
    W = 4096
    el = H.synthetic_walkers(atoms, charges, nspins, W, seed=3).float().to(DEV)
    data = MoleculeData(el, atoms.float().to(DEV), charges.float().to(DEV))
    params = wf.init_params(data, 1)
    spin = SpinSquared(n_up=nspins[0], n_down=nspins[1], phase_logpsi=wf)
    a = spin.evaluate_batch_walkers(params, data)[0]["spin:s2"]
    assert torch.isfinite(a).all()
    sw = el.clone()
    sw[:, [0, 1]] = sw[:, [1, 0]]          # exchange the two majority (spin-up) electrons
    b = spin.evaluate_batch_walkers(params, data.merge({"electrons": sw.contiguous()}))[0]["spin:s2"]
    d = (a - b).abs() / (1.0 + a.abs())
    assert d.median() < 1e-5 and d.quantile(0.99) < 1e-2
