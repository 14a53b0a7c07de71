"""Ragged / degenerate walker batches on the GPU through the C ABI: empty batch, one walker, counts that do not fill a
tensor-core tile (176 rows = 4 groups) or a 16-determinant block.  A walker's result must be bit-identical whatever
batch it is evaluated in: every kernel works row- / matrix-wise and the reductions are per walker."""

import pytest
import torch

import helpers as H
import test_emu_edge_batches as E
from jaqmc_b200 import _marshal as M
from oracle import networks as ON

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _rt():
    from jaqmc_b200._runtime import runtime

    return runtime(DEV)


def test_ferminet_default_widths_sub_batches():
    """N2, 256 x 4 / 32 x 4: tcgen05 dense layers with a ragged last tile, half-warp determinant blocks."""
    wf, sysh, atoms, charges, nspins = E._ferminet(DEV, (256,) * 4, (32,) * 4, 16, "N2")
    el = H.synthetic_walkers(atoms, charges, nspins, 37, seed=2).float().to(DEV).contiguous()
    E.check_sub_batches(_rt(), wf, sysh, el, (0, 1, 3, 17, 37))


@pytest.mark.parametrize("net", ["psiformer", "lapnet"])
def test_attention_nets_sub_batches(net):
    atoms, charges, nspins = H.molecule("LiH")
    A = atoms.shape[0]
    if net == "psiformer":
        p = H.to_f32(H.round_f32(ON.init_psiformer_params(nspins, A, 4, 2, 4, 64, (256,), seed=5)), DEV)
        wf = M.psiformer_handle(p, nspins, A, 4, 2, 4, 64, (256,), "pre")
    else:
        p = H.to_f32(H.round_f32(ON.init_lapnet_params(nspins, A, 4, 2, 4, 64, 2, seed=5)), DEV)
        wf = M.lapnet_handle(p, nspins, A, 4, 2, 4, 64, 2)
    sysh = M.system_handle(atoms.float().to(DEV), charges.float().to(DEV))
    el = H.synthetic_walkers(atoms, charges, nspins, 21, seed=3).float().to(DEV).contiguous()
    E.check_sub_batches(_rt(), wf, sysh, el, (0, 1, 5, 21))


def test_empty_batch_mh_step_is_a_no_op():
    wf, sysh, atoms, charges, nspins = E._ferminet(DEV)
    n = sum(nspins)
    z = lambda *s: torch.zeros(*s, device=DEV)  # noqa: E731
    n_acc, _ = _rt().mh_step(wf, sysh, z(0, n, 3), z(0), z(3, 0, n, 3), z(3, 0), torch.full((1,), 0.1, device=DEV),
                             logpsi_valid=False)
    assert float(n_acc) == 0.0
