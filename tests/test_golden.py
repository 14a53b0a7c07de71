"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py): the CPU test freezes the oracle, the
GPU test checks the CUDA path against the committed numbers."""

import os
import sys

import numpy as np
import pytest
import torch

import helpers as H
from jaqmc_b200 import _marshal as M

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_golden as G  # noqa: E402


def _load(name):
    z = np.load(os.path.join(GOLD, f"{name}.npz"))
    ref = {k[4:]: z[k] for k in z.files if k.startswith("out:")}
    return z, ref


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_reproduces_golden(name):
    z, ref = _load(name)
    p64, el, out = G.evaluate(name)
    flat = G.flatten(p64)
    chk = sum(float(np.abs(v).sum()) for v in flat.values())
    assert np.isclose(chk, z["param_checksum"][0], rtol=1e-12) and sum(v.size for v in flat.values()) == z["param_checksum"][1]
    for k in flat:
        if f"param:{k}" in z.files:
            assert np.array_equal(flat[k].astype(np.float32), z[f"param:{k}"]), k
    assert np.array_equal(el.numpy().astype(np.float32), z["electrons"])
    assert np.array_equal(out["sign"], ref["sign"])
    for k in ("logpsi", "grad", "lap", "e_kin", "e_pot"):
        np.testing.assert_allclose(out[k], ref[k], rtol=1e-9, atol=1e-9, err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.CASES))
def test_cuda_matches_golden(name):
    from jaqmc_b200._runtime import runtime

    dev = torch.device("cuda", 0)
    rt = runtime(dev)
    z, ref = _load(name)
    kind, mol, W, kw = G.CASES[name]
    atoms, charges, nspins, p64, _ = G.build(kind, mol, kw)
    p32 = H.to_f32(p64, dev)
    A = atoms.shape[0]
    if kind == "ferminet":
        wf = M.ferminet_handle(p32, nspins, A, kw["ndets"], kw["hidden_single"], kw["hidden_double"])
    elif kind == "lapnet":
        wf = M.lapnet_handle(p32, nspins, A, kw["ndets"], kw["num_layers"], kw["heads"], kw["heads_dim"])
    else:
        wf = M.psiformer_handle(p32, nspins, A, kw["ndets"], kw["num_layers"], kw["heads"], kw["heads_dim"], kw["mlp_hidden"])
    sysh = M.system_handle(atoms.float().to(dev), charges.float().to(dev))
    el = torch.from_numpy(z["electrons"]).to(dev)
    out = {k: v.cpu().numpy() for k, v in rt.local_energy(wf, sysh, el).items()}
    assert np.array_equal(out["sign"], ref["sign"])
    wide = kind != "ferminet" and kw["heads"] * kw["heads_dim"] >= 64  # tensor-core path, see test_gpu_attention_nets
    H.assert_fp32_parity(out, ref, z["electrons"], l_tol=4e-6 if wide else 1e-6)
