"""Shared test utilities: float32 parameter trees from the oracle's initialisers, synthetic walkers, oracle
evaluation of a walker batch, and the two runtimes (host emulation for CPU tests, CUDA for -m gpu tests)."""

from __future__ import annotations

import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from jaqmc_b200 import _abi  # noqa: E402
from jaqmc_b200._runtime import Runtime  # noqa: E402
from jaqmc_b200.systems import molecule, solid_system, solid_walkers, synthetic_walkers  # noqa: E402,F401
from oracle import estimators as OE  # noqa: E402
from oracle import networks as ON  # noqa: E402

F64 = torch.float64
_emu_rt = None


def emu_runtime() -> Runtime:
    """Host-emulation build of the kernels (tests/emu) behind the product's own marshalling layer."""
    global _emu_rt
    if _emu_rt is None:
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu

        lib = _abi.bind(ctypes.CDLL(build_emu.build()))
        _emu_rt = Runtime(lib, "cpu", _emulation=True)
    return _emu_rt


def to_f32(tree, device="cpu"):
    return ON.tree_map(lambda t: t.to(torch.float32).contiguous().to(device), tree)


def round_f32(tree):
    """float64 tree holding exactly the float32-representable values (so oracle and kernels see identical numbers)."""
    return ON.tree_map(lambda t: t.to(torch.float32).to(F64), tree)


def oracle_batch(logpsi_fn, electrons, atoms=None, charges=None, track=True):
    """Evaluate the float64 oracle walker by walker.  ``logpsi_fn(e)`` -> (sign, logpsi)."""
    W = electrons.shape[0]
    out = dict(logpsi=[], sign=[], grad=[], lap=[], e_kin=[], e_pot=[])
    for w in range(W):
        e = electrons[w]
        if track:
            sign_holder = {}

            def f(x):
                s, lp = logpsi_fn(x)
                sign_holder["s"] = s
                return lp

            v, g, lap = OE.forward_laplacian(f, e)
            out["logpsi"].append(float(v))
            out["sign"].append(float(sign_holder["s"]))
            out["grad"].append(g.numpy())
            out["lap"].append(float(lap))
            out["e_kin"].append(float(-0.5 * lap - 0.5 * (g * g).sum()))
        else:
            s, lp = logpsi_fn(e)
            out["logpsi"].append(float(lp))
            out["sign"].append(float(s))
        if charges is not None:
            out["e_pot"].append(float(OE.potential_energy(e, atoms, charges)))
    return {k: np.asarray(v) for k, v in out.items() if len(v)}


def fp32_scales(ref, electrons):
    """Magnitudes the float32 tolerances are relative to (see tests/test_gpu_ferminet.py docstring)."""
    g2 = (ref["grad"] ** 2).sum(1)
    e_scale = 0.5 * np.abs(ref["lap"]) + 0.5 * g2 + np.abs(ref["e_pot"])
    r = np.linalg.norm(np.asarray(electrons, dtype=np.float64).reshape(len(g2), -1), axis=1)
    l_scale = np.abs(ref["logpsi"]) + np.sqrt(g2) * r
    return e_scale, l_scale


def fp32_errors(out, ref, electrons):
    e_scale, l_scale = fp32_scales(ref, electrons)
    e_out = out["e_loc"] if "e_loc" in out else out["e_kin"] + out["e_pot"]
    e_err = np.abs(e_out - (ref["e_kin"] + ref["e_pot"])) / e_scale
    l_err = np.abs(out["logpsi"] - ref["logpsi"]) / l_scale
    return e_err, l_err


def assert_fp32_parity(out, ref, electrons, e_tol=1e-5, l_tol=1e-6, outlier=10.0):
    """float32 parity against the float64 oracle, errors taken relative to the magnitude of the terms summed
    (``fp32_scales``): the typical walker (median) meets the north-star tolerances (E_L 1e-5, log|psi| 1e-6) and no
    walker exceeds ``outlier`` times them (near-node walkers: ill-conditioned orbital matrices, see DESIGN.md
    "Numerics").  Unscaled medians are also held to 10x the tolerances."""
    e_err, l_err = fp32_errors(out, ref, electrons)
    assert np.median(e_err) < e_tol, ("E_L", e_err)
    assert np.median(l_err) < l_tol, ("logpsi", l_err)
    assert e_err.max() < outlier * e_tol, ("E_L", e_err)
    assert l_err.max() < outlier * l_tol, ("logpsi", l_err)
    e_ref = ref["e_kin"] + ref["e_pot"]
    e_out = out["e_loc"] if "e_loc" in out else out["e_kin"] + out["e_pot"]
    plain_e = np.abs(e_out - e_ref) / (np.abs(ref["e_kin"]) + np.abs(ref["e_pot"]))
    plain_l = np.abs(out["logpsi"] - ref["logpsi"]) / np.maximum(1.0, np.abs(ref["logpsi"]))
    assert np.median(plain_e) < 10 * e_tol, ("E_L median", plain_e)
    assert np.median(plain_l) < 20 * l_tol, ("logpsi median", plain_l)
    if "grad" in out:
        gs = np.abs(ref["grad"]).max(axis=1, keepdims=True) + 1.0
        gerr = np.abs(out["grad"] - ref["grad"]) / gs
        assert np.median(gerr) < 1e-5, np.median(gerr)
        assert np.median(gerr.max(axis=1)) < 1e-4 and gerr.max() < 1e-2, gerr.max(axis=1)
    return e_err, l_err


# ---------------------------------------------------------------------------------------------------------------
# Unscaled north-star errors (BASELINE.json: E_L 1e-5 relative, log|psi| 1e-6, float32).  Every GPU parity test that
# evaluates a named configuration records them here; the table in DESIGN.md is generated from the JSON the GPU run
# writes (gpurun_out/parity_table.json), next to the same numbers of the float32 twin of the oracle where computed.
# ---------------------------------------------------------------------------------------------------------------
def _quantiles(x):
    x = np.asarray(x, dtype=np.float64)
    return {"median": float(np.median(x)), "p99": float(np.quantile(x, 0.99)), "max": float(x.max())}


def unscaled_errors(out, ref):
    """``|dE_L| / |E_L|``, ``|dlog psi|`` (absolute) and ``|dlog psi|`` in float32 ulps of ``log psi``.  An absolute
    error of 1e-6 is below one float32 ulp once ``|log psi| >= 16`` (N2: log psi ~ -27), so no float32 evaluation --
    the reference's included -- can meet "1e-6" read as an absolute bound; it is read as a relative one, like E_L's."""
    e_ref = ref["e_kin"] + ref["e_pot"]
    e_out = out["e_loc"] if "e_loc" in out else out["e_kin"] + out["e_pot"]
    e_rel = np.abs(e_out - e_ref) / np.abs(e_ref)
    l_abs = np.abs(out["logpsi"] - ref["logpsi"])
    l_ulp = l_abs / np.spacing(np.abs(ref["logpsi"]).astype(np.float32)).astype(np.float64)
    return e_rel, l_abs, l_ulp


def parity_report(name, out, ref, electrons, twin=None, assert_literal=True):
    """Records the unscaled errors of configuration ``name`` and asserts the LITERAL north-star tolerances on the
    well-conditioned walkers: those whose cancellation ratios ``(1/2|lap| + 1/2|grad|^2 + |V|) / |E_L|`` and
    ``(|log psi| + |grad||r|) / |log psi|`` are below 10 -- for them the median error must meet 1e-5 (E_L, relative) and
    1e-6 (log|psi|, relative)."""
    e_rel, l_abs, l_ulp = unscaled_errors(out, ref)
    e_scale, l_scale = fp32_scales(ref, electrons)
    e_ref = np.abs(ref["e_kin"] + ref["e_pot"])
    well_e = e_scale / e_ref < 10.0
    well_l = l_scale / np.maximum(np.abs(ref["logpsi"]), 1e-30) < 10.0
    rec = {"walkers": int(len(e_rel)), "E_L_rel": _quantiles(e_rel), "logpsi_abs": _quantiles(l_abs),
           "logpsi_rel": _quantiles(l_abs / np.abs(ref["logpsi"])),
           "logpsi_ulp": _quantiles(l_ulp), "well_conditioned_E": int(well_e.sum()), "well_conditioned_l": int(well_l.sum())}
    if well_e.any():
        rec["E_L_rel_well"] = _quantiles(e_rel[well_e])
    if well_l.any():
        rec["logpsi_abs_well"] = _quantiles(l_abs[well_l])
        rec["logpsi_rel_well"] = _quantiles(l_abs[well_l] / np.abs(ref["logpsi"][well_l]))
    if twin is not None:
        te, tl, tu = unscaled_errors(twin, ref)
        rec["twin_E_L_rel"] = _quantiles(te)
        rec["twin_logpsi_abs"] = _quantiles(tl)
        rec["twin_logpsi_rel"] = _quantiles(tl / np.abs(ref["logpsi"]))
        rec["twin_logpsi_ulp"] = _quantiles(tu)
    path = os.path.join(ROOT, "gpurun_out", "parity_table.json")
    try:
        import json

        os.makedirs(os.path.dirname(path), exist_ok=True)
        table = {}
        if os.path.exists(path):
            with open(path) as f:
                table = json.load(f)
        table[name] = rec
        with open(path, "w") as f:
            json.dump(table, f, indent=1, sort_keys=True)
    except OSError:
        pass
    print(f"[parity] {name}: E_L rel median {rec['E_L_rel']['median']:.2e} p99 {rec['E_L_rel']['p99']:.2e} max "
          f"{rec['E_L_rel']['max']:.2e} | logpsi abs median {rec['logpsi_abs']['median']:.2e} max "
          f"{rec['logpsi_abs']['max']:.2e} ({rec['logpsi_ulp']['median']:.2f} / {rec['logpsi_ulp']['max']:.2f} ulp)")
    if assert_literal:
        if well_e.any():
            assert np.median(e_rel[well_e]) < 1e-5, ("literal E_L tolerance", e_rel[well_e])
        if well_l.any():
            l_rel = l_abs[well_l] / np.abs(ref["logpsi"][well_l])
            assert np.median(l_rel) < 1e-6, ("literal logpsi tolerance", l_rel)
    return rec
