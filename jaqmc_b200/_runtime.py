"""Thin call layer over the C ABI: pointer marshalling, workspace ownership, stream selection.

PyTorch is used for device memory and streams only.  ``Runtime`` is parameterised by the bound library so
that the CPU test-suite can drive the host-emulation build of the same kernels (tests/emu) through the very
same marshalling code; the product singleton ``runtime()`` binds the CUDA library and only accepts CUDA tensors.
"""

from __future__ import annotations

import contextlib
import ctypes as C
import functools

import torch

from . import _abi


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _on_device(fn):
    """Run ``fn`` with the runtime's device current: the library queries the current device for its per-device
    one-off setup (shared-memory opt-ins, SM count) and launches on the stream it is handed."""

    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        with self._device_ctx():
            return fn(self, *args, **kwargs)

    return wrapped


class Runtime:
    def __init__(self, cdll, device, workspace_limit_bytes: int | None = None, _emulation: bool = False):
        self.lib = cdll
        self.device = torch.device(device)
        ver = cdll.jaqmc_b200_version()
        if (b"EMULATION" in ver) != bool(_emulation):
            raise RuntimeError(f"jaqmc_b200: refusing library {ver!r} (emulation builds are test infrastructure only)")
        if not _emulation and self.device.type != "cuda":
            raise RuntimeError("jaqmc_b200: the product path runs on CUDA devices only")
        self.workspace_limit_bytes = workspace_limit_bytes
        self._ws = None

    # ---- plumbing -------------------------------------------------------------------------------
    def _device_ctx(self):
        if self.device.type == "cuda":
            return torch.cuda.device(self.device)
        return contextlib.nullcontext()

    def _stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return C.c_void_p(0)

    def _check_tensor(self, t, name, dtype=torch.float32):
        same = t.device.type == self.device.type and (
            self.device.type != "cuda" or self.device.index is None or t.device.index == self.device.index)
        if not same:
            raise ValueError(f"{name}: tensor on {t.device}, runtime on {self.device}")
        if t.dtype != dtype or not t.is_contiguous():
            raise ValueError(f"{name}: expected contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")

    def workspace(self, nbytes: int):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def workspace_bytes(self, wf, n_walkers: int, track: bool) -> int:
        need = int(self.lib.jaqmc_b200_workspace_bytes(C.byref(wf.struct), int(n_walkers), int(track)))
        if need == 0 and n_walkers > 0:
            raise _abi.JaqmcB200Error(_abi.ERR_INVALID_ARGUMENT, self.lib.jaqmc_b200_last_error().decode())
        return need

    def _ws_for(self, wf, n_walkers, track):
        need = self.workspace_bytes(wf, n_walkers, track)
        if self.workspace_limit_bytes is not None:
            need = min(need, self.workspace_limit_bytes)
        elif self.device.type == "cuda":
            free, _ = torch.cuda.mem_get_info(self.device)
            held = self._ws.numel() if self._ws is not None else 0
            need = min(need, int(0.8 * (free + held)))
        return self.workspace(max(need, 1 << 16))

    # ---- entry points ---------------------------------------------------------------------------
    @_on_device
    def logpsi(self, wf, system, electrons):
        """``electrons`` (W, n, 3) -> (logpsi (W,), sign (W,))."""
        self._check_tensor(electrons, "electrons")
        W = electrons.shape[0]
        logpsi = torch.empty(W, dtype=torch.float32, device=self.device)
        sign = torch.empty(W, dtype=torch.float32, device=self.device)
        ws = self._ws_for(wf, W, False)
        rc = self.lib.jaqmc_b200_logpsi(C.byref(wf.struct), C.byref(system.struct), _ptr(electrons), W, _ptr(logpsi),
                                        _ptr(sign), _ptr(ws), ws.numel(), self._stream())
        _abi.check(self.lib, rc)
        return logpsi, sign

    @_on_device
    def orbitals(self, wf, system, electrons, n_dets, complex_valued=False):
        """``electrons`` (W, n, 3) -> orbital matrices (W, ndets, n, n) (complex64 for the periodic network)."""
        self._check_tensor(electrons, "electrons")
        W, n = electrons.shape[0], electrons.shape[1]
        shape = (W, n_dets, n, n, 2) if complex_valued else (W, n_dets, n, n)
        orb = torch.empty(shape, dtype=torch.float32, device=self.device)
        ws = self._ws_for(wf, W, False)
        rc = self.lib.jaqmc_b200_orbitals(C.byref(wf.struct), C.byref(system.struct), _ptr(electrons), W, _ptr(orb),
                                          _ptr(ws), ws.numel(), self._stream())
        _abi.check(self.lib, rc)
        return torch.view_as_complex(orb) if complex_valued else orb

    @_on_device
    def local_energy(self, wf, system, electrons, sums=None):
        """Returns a dict with logpsi, sign, grad (W,3n), lap, e_kin, e_pot, e_loc."""
        self._check_tensor(electrons, "electrons")
        W, n = electrons.shape[0], electrons.shape[1]
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)  # noqa: E731
        out = dict(logpsi=f(W), sign=f(W), grad=f(W, 3 * n), lap=f(W), e_kin=f(W), e_pot=f(W), e_loc=f(W))
        ws = self._ws_for(wf, W, True)
        rc = self.lib.jaqmc_b200_local_energy(
            C.byref(wf.struct), C.byref(system.struct), _ptr(electrons), W, _ptr(out["logpsi"]), _ptr(out["sign"]),
            _ptr(out["grad"]), _ptr(out["lap"]), _ptr(out["e_kin"]), _ptr(out["e_pot"]), _ptr(out["e_loc"]),
            _ptr(sums), _ptr(ws), ws.numel(), self._stream())
        _abi.check(self.lib, rc)
        return out

    @_on_device
    def local_energy_complex(self, wf, system, electrons, ewald=None, cell_atoms=None, cell_charges=None, sums=None):
        """Periodic (complex log psi) local energy: dict with complex64 logpsi / grad (W,3n) / lap / e_kin / e_loc and
        float32 e_pot (present when ``ewald`` -- a ``jaqmc_b200.ewald.EwaldSum`` -- is given)."""
        self._check_tensor(electrons, "electrons")
        W, n = electrons.shape[0], electrons.shape[1]
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)  # noqa: E731
        raw = dict(logpsi=f(W, 2), grad=f(W, 3 * n, 2), lap=f(W, 2), e_kin=f(W, 2), e_loc=f(W, 2))
        e_pot = f(W) if ewald is not None else None
        if ewald is not None:
            self._check_tensor(cell_atoms, "cell_atoms")
            self._check_tensor(cell_charges, "cell_charges")
        ws = self._ws_for(wf, W, True)
        rc = self.lib.jaqmc_b200_local_energy_complex(
            C.byref(wf.struct), C.byref(system.struct), C.byref(ewald._struct) if ewald is not None else None,
            _ptr(cell_atoms), _ptr(cell_charges), 0 if cell_atoms is None else cell_atoms.shape[0], _ptr(electrons), W,
            _ptr(raw["logpsi"]), _ptr(raw["grad"]), _ptr(raw["lap"]), _ptr(raw["e_kin"]), _ptr(e_pot), _ptr(raw["e_loc"]),
            _ptr(sums), _ptr(ws), ws.numel(), self._stream())
        _abi.check(self.lib, rc)
        out = {k: torch.view_as_complex(v) for k, v in raw.items()}
        if e_pot is not None:
            out["e_pot"] = e_pot
        return out

    @_on_device
    def capture_local_energy(self, wf, system, electrons, sums=None):
        """CUDA-graph version of :meth:`local_energy` for a fixed walker-batch shape: the ~60 kernel launches of one
        evaluation are captured once and replayed with a single launch.  Returns ``(replay, out)``: ``replay()``
        re-evaluates on the current contents of ``electrons`` (same tensor) and refreshes the tensors in ``out``.
        The library only enqueues on the stream and never allocates, so the whole call is capturable."""
        self.local_energy(wf, system, electrons, sums=sums)  # warm-up: one-off attribute / descriptor setup
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.local_energy(wf, system, electrons, sums=sums)
        keep = (wf, system, electrons, sums, self._ws)

        def replay():
            graph.replay()
            return out

        replay._keep = keep
        return replay, out

    @_on_device
    def coulomb(self, system, electrons):
        self._check_tensor(electrons, "electrons")
        W, n = electrons.shape[0], electrons.shape[1]
        e_pot = torch.empty(W, dtype=torch.float32, device=self.device)
        rc = self.lib.jaqmc_b200_coulomb(C.byref(system.struct), _ptr(electrons), W, n, _ptr(e_pot), self._stream())
        _abi.check(self.lib, rc)
        return e_pot

    @_on_device
    def mh_step(self, wf, system, electrons, logpsi, normals, uniforms, stddev, logpsi_valid=True,
                record_accepts=False, wrap_lattice=None):
        """In-place MH sub-steps on ``electrons`` / ``logpsi``; returns (n_accept (1,), accepted u8 (S,W) or None)."""
        for t, nm in ((electrons, "electrons"), (logpsi, "logpsi"), (normals, "normals"), (uniforms, "uniforms"),
                      (stddev, "stddev")):
            self._check_tensor(t, nm)
        S, W = normals.shape[0], electrons.shape[0]
        if tuple(normals.shape) != (S, *electrons.shape) or tuple(uniforms.shape) != (S, W):
            raise ValueError(f"mh_step: normals {tuple(normals.shape)} / uniforms {tuple(uniforms.shape)} do not match "
                             f"electrons {tuple(electrons.shape)}")
        n_accept = torch.zeros(1, dtype=torch.float32, device=self.device)
        accepted = torch.empty(S, W, dtype=torch.uint8, device=self.device) if record_accepts else None
        ws = self._ws_for(wf, W, False)
        if wrap_lattice is not None:  # periodic proposal (solid): wrap into the simulation cell
            lat = (C.c_float * 9)(*[float(v) for v in torch.as_tensor(wrap_lattice).reshape(-1).tolist()])
            rc = self.lib.jaqmc_b200_mh_step_pbc(
                C.byref(wf.struct), C.byref(system.struct), _ptr(electrons), _ptr(logpsi), int(bool(logpsi_valid)),
                _ptr(normals), _ptr(uniforms), _ptr(stddev), S, W, _ptr(n_accept), _ptr(accepted), lat, _ptr(ws),
                ws.numel(), self._stream())
            _abi.check(self.lib, rc)
            return n_accept, accepted
        rc = self.lib.jaqmc_b200_mh_step(
            C.byref(wf.struct), C.byref(system.struct), _ptr(electrons), _ptr(logpsi), int(bool(logpsi_valid)),
            _ptr(normals), _ptr(uniforms), _ptr(stddev), S, W, _ptr(n_accept), _ptr(accepted), _ptr(ws), ws.numel(),
            self._stream())
        _abi.check(self.lib, rc)
        return n_accept, accepted

    @_on_device
    def capture_mh_step(self, wf, system, electrons, normals, uniforms, stddev, wrap_lattice=None):
        """CUDA-graph version of :meth:`mh_step` for fixed shapes (the ~350 launches of ten sub-steps replayed with one
        launch: at a few hundred walkers per GPU the sampling pass is launch-bound).  ``replay()`` runs the sub-steps on
        the CURRENT contents of ``electrons`` / ``normals`` / ``uniforms`` / ``stddev`` (same tensors) in place and
        returns ``n_accept`` (1,)."""
        W = electrons.shape[0]
        logpsi = torch.empty(W, dtype=torch.float32, device=self.device)
        keep0 = electrons.clone()
        self.mh_step(wf, system, electrons, logpsi, normals, uniforms, stddev, logpsi_valid=False,
                     wrap_lattice=wrap_lattice)  # warm-up
        electrons.copy_(keep0)
        torch.cuda.synchronize(self.device)
        n_accept = torch.zeros(1, dtype=torch.float32, device=self.device)
        ws = self._ws_for(wf, W, False)
        S = normals.shape[0]
        graph = torch.cuda.CUDAGraph()
        lat = None
        if wrap_lattice is not None:
            lat = (C.c_float * 9)(*[float(v) for v in torch.as_tensor(wrap_lattice).reshape(-1).tolist()])
        with torch.cuda.graph(graph):
            n_accept.zero_()
            if lat is None:
                rc = self.lib.jaqmc_b200_mh_step(
                    C.byref(wf.struct), C.byref(system.struct), _ptr(electrons), _ptr(logpsi), 0, _ptr(normals),
                    _ptr(uniforms), _ptr(stddev), S, W, _ptr(n_accept), None, _ptr(ws), ws.numel(), self._stream())
            else:
                rc = self.lib.jaqmc_b200_mh_step_pbc(
                    C.byref(wf.struct), C.byref(system.struct), _ptr(electrons), _ptr(logpsi), 0, _ptr(normals),
                    _ptr(uniforms), _ptr(stddev), S, W, _ptr(n_accept), None, lat, _ptr(ws), ws.numel(),
                    self._stream())
            _abi.check(self.lib, rc)
        keep = (wf, system, electrons, normals, uniforms, stddev, logpsi, ws)

        def replay():
            graph.replay()
            return n_accept

        replay._keep = keep
        return replay

    @_on_device
    def psi_ratios(self, wf, system, electrons, move_index, move_pos):
        """``(log|psi'/psi|, sign ratio)`` (W, Q) for Q moved configurations per walker: ``move_index`` (Q, 2) int32
        (second entry -1: one electron), ``move_pos`` (W, Q, 2, 3)."""
        self._check_tensor(electrons, "electrons")
        self._check_tensor(move_pos, "move_pos")
        self._check_tensor(move_index, "move_index", dtype=torch.int32)
        W, n = electrons.shape[0], electrons.shape[1]
        Q = move_index.shape[0]
        if tuple(move_index.shape) != (Q, 2) or tuple(move_pos.shape) != (W, Q, 2, 3):
            raise ValueError(f"psi_ratios: move_index {tuple(move_index.shape)} / move_pos {tuple(move_pos.shape)} do not "
                             f"match ({Q}, 2) / ({W}, {Q}, 2, 3)")
        lr = torch.empty(W, Q, dtype=torch.float32, device=self.device)
        sr = torch.empty(W, Q, dtype=torch.float32, device=self.device)
        if W * Q == 0:
            return lr, sr
        need = self.workspace_bytes(wf, W * Q, False) + 4 * (3 * n * W * Q + 2 * W) + 4096
        if self.workspace_limit_bytes is not None:
            need = min(need, self.workspace_limit_bytes)
        elif self.device.type == "cuda":
            free, _ = torch.cuda.mem_get_info(self.device)
            held = self._ws.numel() if self._ws is not None else 0
            need = min(need, int(0.8 * (free + held)))
        ws = self.workspace(max(need, 1 << 16))
        rc = self.lib.jaqmc_b200_psi_ratios(C.byref(wf.struct), C.byref(system.struct), _ptr(electrons), W, Q,
                                            _ptr(move_index), _ptr(move_pos), _ptr(lr), _ptr(sr), _ptr(ws), ws.numel(),
                                            self._stream())
        _abi.check(self.lib, rc)
        return lr, sr

    @_on_device
    def ferminet_logpsi_vjp(self, wf, grads_handle, system, electrons, cotangent):
        """``sum_w cotangent[w] d log|psi|(x_w) / d theta`` into the leaves behind ``grads_handle`` (a ``ferminet_handle``
        built on a tree of gradient buffers); returns ``(logpsi (W,), sign (W,))``."""
        self._check_tensor(electrons, "electrons")
        self._check_tensor(cotangent, "cotangent")
        W = electrons.shape[0]
        if tuple(cotangent.shape) != (W,):
            raise ValueError(f"cotangent: expected shape ({W},), got {tuple(cotangent.shape)}")
        logpsi = torch.empty(W, dtype=torch.float32, device=self.device)
        sign = torch.empty(W, dtype=torch.float32, device=self.device)
        need = int(self.lib.jaqmc_b200_ferminet_vjp_workspace_bytes(C.byref(wf.config_struct), W))
        if need == 0 and W > 0:
            raise _abi.JaqmcB200Error(_abi.ERR_INVALID_ARGUMENT, self.lib.jaqmc_b200_last_error().decode())
        ws = self.workspace(max(need, 1 << 16))
        rc = self.lib.jaqmc_b200_ferminet_logpsi_vjp(
            C.byref(wf.config_struct), C.byref(wf.params_struct), C.byref(system.struct), _ptr(electrons), W,
            _ptr(cotangent), C.byref(grads_handle.params_struct), _ptr(logpsi), _ptr(sign), _ptr(ws), ws.numel(),
            self._stream())
        _abi.check(self.lib, rc)
        return logpsi, sign

    @_on_device
    def mh_propose(self, x1, normals, stddev):
        """``x1 + normals * stddev`` (sampler/mcmc.py:53-54) for samplers that drive their own loop."""
        for t, nm in ((x1, "x1"), (normals, "normals"), (stddev, "stddev")):
            self._check_tensor(t, nm)
        x2 = torch.empty_like(x1)
        rc = self.lib.jaqmc_b200_mh_propose(_ptr(x1), _ptr(normals), _ptr(stddev), _ptr(x2), x1.numel(), self._stream())
        _abi.check(self.lib, rc)
        return x2

    @_on_device
    def mh_accept(self, x1, x2, logprob1, logprob2, uniforms, n_accept, accepted=None):
        """In-place accept / select on ``x1`` / ``logprob1`` (sampler/mcmc.py:128-137): accept iff
        ``logprob2 - logprob1 > log(u)``."""
        for t, nm in ((x1, "x1"), (x2, "x2"), (logprob1, "logprob1"), (logprob2, "logprob2"), (uniforms, "uniforms"),
                      (n_accept, "n_accept")):
            self._check_tensor(t, nm)
        W = x1.shape[0]
        rc = self.lib.jaqmc_b200_mh_accept(_ptr(x1), _ptr(x2), _ptr(logprob1), _ptr(logprob2), _ptr(uniforms), W,
                                           x1.numel() // max(W, 1), _ptr(n_accept), _ptr(accepted), self._stream())
        _abi.check(self.lib, rc)

    def launch_count(self) -> int:
        return int(self.lib.jaqmc_b200_launch_count())

    def reset_launch_count(self) -> None:
        self.lib.jaqmc_b200_reset_launch_count()


_runtimes: dict = {}


def runtime(device=None) -> Runtime:
    """Product runtime for a CUDA device (default: current device).  Raises when the CUDA library is missing."""
    from ._lib import cuda_library

    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("jaqmc_b200: no CUDA device available; there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"jaqmc_b200: device {device} is not a CUDA device; there is no CPU fallback")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = device.index
    if key not in _runtimes:
        _runtimes[key] = Runtime(cuda_library(), device)
    return _runtimes[key]
