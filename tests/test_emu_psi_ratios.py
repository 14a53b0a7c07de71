"""``jaqmc_b200_psi_ratios`` (psi-ratio consumers: ECP non-local integral, SpinSquared) on the CPU emulation build:
ratios of one- and two-electron moves against the float64 oracle, with and without configuration tiling."""

import numpy as np
import torch

import helpers as H
from jaqmc_b200 import _marshal as M
from oracle import networks as ON


def test_psi_ratios_match_oracle_and_tile():
    rt = H.emu_runtime()
    atoms, charges, nspins = H.molecule("LiH")
    hs, hd, ndets = (16, 16), (8, 8), 3
    p64 = H.round_f32(ON.init_ferminet_params(nspins, 2, ndets, hs, hd, seed=4))
    W, n = 4, sum(nspins)
    el = H.synthetic_walkers(atoms, charges, nspins, W, seed=2)
    wf = M.ferminet_handle(H.to_f32(p64), nspins, 2, ndets, hs, hd)
    sysh = M.system_handle(atoms.float(), None)
    g = torch.Generator().manual_seed(0)
    # moves: a same-spin swap (0 <-> 1), a single displacement of electron 1, a single displacement of electron 3, the identity
    idx = torch.tensor([[0, 1], [1, -1], [3, -1], [2, -1]], dtype=torch.int32)
    Q = idx.shape[0]
    pos = torch.zeros(W, Q, 2, 3)
    pos[:, 0, 0] = el[:, 1].float()
    pos[:, 0, 1] = el[:, 0].float()
    pos[:, 1, 0] = el[:, 1].float() + 0.3 * torch.randn(W, 3, generator=g)
    pos[:, 2, 0] = el[:, 3].float() + 0.3 * torch.randn(W, 3, generator=g)
    pos[:, 3, 0] = el[:, 2].float()
    e32 = el.float().contiguous()
    lr, sr = rt.psi_ratios(wf, sysh, e32, idx, pos.contiguous())
    for w in range(W):
        s0, lp0 = ON.ferminet_logpsi(p64, el[w], atoms, nspins)
        for q in range(Q):
            x = el[w].clone()
            for s in range(2):
                if idx[q, s] >= 0:
                    x[idx[q, s]] = pos[w, q, s].double()
            s1, lp1 = ON.ferminet_logpsi(p64, x, atoms, nspins)
            assert float(sr[w, q]) == float(s1 * s0)
            assert abs(float(lr[w, q]) - float(lp1 - lp0)) < 5e-5
    assert (sr[:, 0] == -1).all() and (lr[:, 0].abs() < 1e-4).all()    # same-spin swap: antisymmetry
    assert (sr[:, 3] == 1).all() and (lr[:, 3].abs() < 1e-6).all()      # identity move
    # a small workspace tiles the configurations: identical results
    need1 = rt.workspace_bytes(wf, 1, False)
    rt2 = H.Runtime(rt.lib, "cpu", workspace_limit_bytes=int(need1 * 3 + 4096), _emulation=True)
    lr2, sr2 = rt2.psi_ratios(wf, sysh, e32, idx, pos.contiguous())
    assert torch.equal(lr, lr2) and torch.equal(sr, sr2)
