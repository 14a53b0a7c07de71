"""Generates the committed golden vectors ``tests/golden/*.npz``.

The reference (bytedance/jaqmc) needs jax / flax, which are not installable offline, so it cannot be run here to
produce vectors.  These fixtures are therefore outputs of the float64 oracle (``oracle/``), which
is itself pinned against the reference's known-answer tests (``tests/test_oracle_known_answers.py``).  They freeze
the oracle (a change in its arithmetic shows up as a CPU test failure) and give the GPU tests a fixed target that
does not depend on torch's CPU kernels on the GPU box.

    python tests/golden/make_golden.py          # rewrites the fixtures (seeded, deterministic)
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from oracle import networks as ON  # noqa: E402

# name: (kind, molecule, walkers, kwargs of the oracle's parameter initialiser)
CASES = {
    "ferminet_li": ("ferminet", "Li", 6, dict(ndets=4, hidden_single=(32, 32, 32), hidden_double=(8, 8, 8), seed=101)),
    "ferminet_lih": ("ferminet", "LiH", 4, dict(ndets=16, hidden_single=(256,) * 4, hidden_double=(32,) * 4, seed=102)),
    "lapnet_li": ("lapnet", "Li", 6, dict(ndets=4, num_layers=2, heads=2, heads_dim=16, seed=103)),
    "lapnet_lih": ("lapnet", "LiH", 4, dict(ndets=16, num_layers=4, heads=4, heads_dim=64, seed=104)),
    "psiformer_li": ("psiformer", "Li", 6, dict(ndets=4, num_layers=2, heads=2, heads_dim=16, mlp_hidden=(32,), seed=105)),
    "psiformer_lih": ("psiformer", "LiH", 4, dict(ndets=16, num_layers=4, heads=4, heads_dim=64, mlp_hidden=(256,), seed=106)),
}


def build(kind, mol, kwargs):
    atoms, charges, nspins = H.molecule(mol)
    A = atoms.shape[0]
    if kind == "ferminet":
        p = ON.init_ferminet_params(nspins, A, **kwargs)
        fn = lambda p64, e: ON.ferminet_logpsi(p64, e, atoms, nspins)  # noqa: E731
    elif kind == "lapnet":
        p = ON.init_lapnet_params(nspins, A, **kwargs)
        fn = lambda p64, e: ON.lapnet_logpsi(p64, e, atoms, nspins, kwargs["heads"])  # noqa: E731
    else:
        p = ON.init_psiformer_params(nspins, A, **kwargs)
        fn = lambda p64, e: ON.psiformer_logpsi(p64, e, atoms, nspins)  # noqa: E731
    return atoms, charges, nspins, H.round_f32(p), fn


def flatten(tree, prefix=""):
    out = {}
    for k in sorted(tree):
        v = tree[k]
        if isinstance(v, dict):
            out.update(flatten(v, f"{prefix}{k}/"))
        else:
            out[f"{prefix}{k}"] = v.numpy()
    return out


def unflatten(flat):
    tree = {}
    for path, v in flat.items():
        node = tree
        keys = path.split("/")
        for k in keys[:-1]:
            node = node.setdefault(k, {})
        node[keys[-1]] = torch.from_numpy(np.asarray(v))
    return tree


def evaluate(name):
    kind, mol, W, kwargs = CASES[name]
    atoms, charges, nspins, p64, fn = build(kind, mol, kwargs)
    el = H.synthetic_walkers(atoms, charges, nspins, W, seed=kwargs["seed"])
    ref = H.oracle_batch(lambda e: fn(p64, e), el, atoms, charges)
    return p64, el, ref


def main():
    for name in CASES:
        p64, el, ref = evaluate(name)
        flat = flatten(p64)
        # small nets carry their parameters; the full-size ones are regenerated from the seed (torch's CPU generator is
        # reproducible) and guarded by a checksum
        nparam = sum(v.size for v in flat.values())
        arrays = {f"param:{k}": v.astype(np.float32) for k, v in flat.items()} if nparam < 50_000 else {}
        arrays["param_checksum"] = np.array([sum(float(np.abs(v).sum()) for v in flat.values()), float(nparam)])
        arrays["electrons"] = el.numpy().astype(np.float32)
        for k, v in ref.items():
            arrays[f"out:{k}"] = np.asarray(v, dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **arrays)
        print(name, "walkers", el.shape[0], "E_L", (ref["e_kin"] + ref["e_pot"])[:3])


if __name__ == "__main__":
    main()
