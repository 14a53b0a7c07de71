"""Kernel arithmetic (CPU emulation build, through the product's marshalling layer) against the golden vectors produced
by the REFERENCE's own source (tests/golden/ref_*.npz; see tests/test_reference_fixtures.py)."""

import numpy as np
import pytest
import torch

import helpers as H
import test_reference_fixtures as R
from jaqmc_b200 import _marshal as M


def handles(meta, p64, z, device="cpu"):
    kw = dict(meta["kwargs"])
    nspins = tuple(meta["nspins"])
    A = z["atoms"].shape[0]
    params = H.to_f32(p64, device)
    env = kw.get("envelope", "abs_isotropic")
    if meta["kind"] == "ferminet":
        wf = M.ferminet_handle(params, nspins, A, kw["ndets"], kw["hidden_dims_single"], kw["hidden_dims_double"], env,
                               kw.get("orbitals_spin_split", True), kw.get("use_last_layer", False))
    elif meta["kind"] == "lapnet":
        wf = M.lapnet_handle(params, nspins, A, kw["ndets"], kw["num_layers"], kw["num_heads"], kw["heads_dim"],
                             kw.get("num_local_updates", 2), env, True, kw.get("jastrow", "simple_ee") == "simple_ee",
                             kw.get("use_layernorm", False))
    else:
        wf = M.psiformer_handle(params, nspins, A, kw["ndets"], kw["num_layers"], kw["num_heads"], kw["heads_dim"],
                                tuple(kw["mlp_hidden_dims"]), kw.get("layer_norm_mode", "pre"), env, True, True,
                                kw.get("jastrow", "simple_ee") == "simple_ee")
    f32 = lambda k: torch.from_numpy(z[k].astype(np.float32)).to(device)  # noqa: E731
    return wf, M.system_handle(f32("atoms"), f32("charges")), f32("electrons").contiguous()


@pytest.mark.parametrize("name", R.MOLECULE)
def test_emulated_kernels_match_reference(name):
    meta, p64, z = R.load(name)
    rt = H.emu_runtime()
    try:
        wf, sysh, el = handles(meta, p64, z)
    except NotImplementedError as e:
        pytest.skip(str(e))
    out = {k: v.numpy() for k, v in rt.local_energy(wf, sysh, el).items()}
    ref = {k: z[k] for k in ("logpsi", "sign", "grad", "lap", "e_kin", "e_pot")}
    assert np.array_equal(out["sign"], ref["sign"])
    H.assert_fp32_parity(out, ref, z["electrons"])
    np.testing.assert_allclose(out["e_pot"], ref["e_pot"], rtol=3e-6)
    orb = rt.orbitals(wf, sysh, el, meta["kwargs"]["ndets"]).numpy()
    scale = np.abs(z["orbitals"]).max(axis=(-1, -2), keepdims=True)
    assert (np.abs(orb - z["orbitals"]) / scale).max() < 2e-5
    lp, sg = rt.logpsi(wf, sysh, el)
    assert np.array_equal(sg.numpy(), ref["sign"])
