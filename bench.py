#!/usr/bin/env python
"""Benchmark of the hot path: local-energy evaluations per second (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload n2|li|...]

A *step* is ``evals_per_step`` (default 32 for the FermiNet workloads, 4 for the heavier ones; recorded in ``config``)
consecutive local-energy evaluations (forward-Laplacian kinetic + Coulomb potential + total) of the whole walker batch
of the named system -- the estimator half of the reference's ``EvaluationWorkStage.compute_step``
(workflow/stage/evaluation.py:190-192) -- so that the timed region stays above a second at 8 GPUs too (one evaluation
of the 512-walker shard takes 2 ms).  The walker batch is global (4096, the reference's ``workflow.batch_size``) and
sharded over the ranks: strong scaling, as ``docs/guide/multi-device.md:25-29`` defines it.

``value``   walkers * evals_per_step * K / device time, inputs resident in HBM (max over ranks, CUDA events on the
            launch stream); ``ms_per_eval`` is the time of one evaluation of the batch.
``e2e``     the same metric through the public API (``FermiNetWavefunction.local_energy``) with the step's electrons
            copied from pinned host memory and the per-walker local energies read back, inside the timed region.
``roofline``the dominant kernel's achieved algorithmic throughput against the measured peak (MEASURED_PEAKS.json).
``cpu_baseline`` / ``--impl reference``: the float32 ``torch.func.vmap`` port of the reference's CPU path
            (oracle/, vmapped forward-Laplacian + potential) on the host cores; a step of that arm is a bounded SAMPLE of
            the workload (the walkers it actually evaluated are in ``cpu_baseline.sample`` and ``ms_per_step`` is the time
            of that sample, not an extrapolation).  The port carries dense 3n-wide Jacobians through the pair stream where
            the reference's interpreter keeps them 6 wide (Local2): it does several times the reference's FLOPs there and
            is a lower bound on the reference's JAX-CPU throughput.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (molecule, ndets, hidden_single, hidden_double, description[, network kind])
    "n2": ("N2", 16, (256,) * 4, (32,) * 4, "FermiNet-N2 (14 electrons, 16 dets, 256x4/32x4)"),
    "li": ("Li", 16, (256,) * 4, (32,) * 4, "FermiNet-Li (3 electrons, 16 dets, 256x4/32x4)"),
    "lih": ("LiH", 16, (256,) * 4, (32,) * 4, "FermiNet-LiH (4 electrons, 16 dets, 256x4/32x4)"),
    "n2-lapnet": ("N2", 16, None, None, "LapNet-N2 (14 electrons, 16 dets, 4 layers x 4 heads x 64)", "lapnet"),
    "n2-psiformer": ("N2", 16, None, None, "Psiformer-N2 (14 electrons, 16 dets, 4 x 4 x 64, MLP 256)", "psiformer"),
    # configs[3] / configs[4]: 4096 walkers are sharded over 8 GPUs there, i.e. 512 per GPU -- pass --walkers 512
    "benzene-psiformer": ("C6H6", 16, None, None, "Psiformer-C6H6 (42 electrons, 16 dets, 4 x 4 x 64, MLP 256)", "psiformer"),
    "lih-solid": ("fcc_lih_222", 16, (256,) * 4, (32,) * 4,
                  "periodic FermiNet, LiH rock salt 2x2x2 (32 electrons, 16 atoms, complex orbitals, Ewald)", "solid"),
}


def workload_kind(name):
    w = WORKLOADS[name]
    return w[5] if len(w) > 5 else "ferminet"


def make_wavefunction(name, nspins):
    from jaqmc_b200.wavefunction import FermiNetWavefunction, LapNetWavefunction, PsiformerWavefunction

    mol, ndets, hs, hd, desc = WORKLOADS[name][:5]
    kind = workload_kind(name)
    if kind == "lapnet":
        return LapNetWavefunction(nspins=nspins, ndets=ndets)
    if kind == "psiformer":
        return PsiformerWavefunction(nspins=nspins, ndets=ndets)
    if kind == "solid":
        raise ValueError("the periodic workload is built by run_ours_solid")
    return FermiNetWavefunction(nspins=nspins, ndets=ndets, hidden_dims_single=list(hs), hidden_dims_double=list(hd))
METRIC = "local_energy_evals_per_sec"
UNIT = "evals/s"


def workload_string(name, walkers):
    """``config.workload``: identical in both arms."""
    kind = workload_kind(name)
    what = "forward-Laplacian kinetic + Ewald potential" if kind == "solid" else "forward-Laplacian local energy"
    return f"{WORKLOADS[name][4]}, {walkers} walkers, {what}"


def default_evals_per_step(name):
    return 32 if workload_kind(name) == "ferminet" else 4


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ---------------------------------------------------------------------------------------------
# CPU arm: vmapped float32 port of the reference path (oracle/), bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_rate(workload: str, budget_s: float, chunk: int = 64):
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would time one core)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    import helpers as H
    from oracle import estimators as OE
    from oracle import lap as L
    from oracle import networks as ON

    mol, ndets, hs, hd = WORKLOADS[workload][:4]
    kind = workload_kind(workload)
    atoms, charges, nspins = H.molecule(mol)
    at, ch = atoms.float(), charges.float()
    if kind == "lapnet":
        p = ON.tree_map(lambda t: t.float(), ON.init_lapnet_params(nspins, atoms.shape[0], ndets, seed=1))
        net = lambda x: ON.lapnet_logpsi(p, x, at, nspins)  # noqa: E731
    elif kind == "psiformer":
        p = ON.tree_map(lambda t: t.float(), ON.init_psiformer_params(nspins, atoms.shape[0], ndets, seed=1))
        net = lambda x: ON.psiformer_logpsi(p, x, at, nspins)  # noqa: E731
    else:
        p = ON.tree_map(lambda t: t.float(), ON.init_ferminet_params(nspins, atoms.shape[0], ndets, hs, hd, seed=1))
        net = lambda x: ON.ferminet_logpsi(p, x, at, nspins)  # noqa: E731

    def one(e):
        out = net(L.seed(e))[1]
        g = out.jac.reshape(-1)
        return -0.5 * out.lap - 0.5 * (g * g).sum() + OE.potential_energy(e, at, ch)

    f = torch.func.vmap(one)
    el = H.synthetic_walkers(atoms, charges, nspins, chunk, seed=0).float()
    f(el)  # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        f(el)
        done += chunk
        dt = time.perf_counter() - t0
        if dt >= budget_s:
            break
    return done / dt, done, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if workload_kind(args.workload) == "solid":
        print(json.dumps({"impl": "reference", "unavailable": "no vmapped CPU port of the periodic network (float64 oracle only)"}))
        return
    # K steps, each a bounded sample of the workload sized so that the whole run ends within a few minutes; the line
    # reports what actually ran: ms_per_step is the time of one sample, value = walkers evaluated / time
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    rates, total, secs = [], 0, 0.0
    for i in range(args.warmup + args.steps):
        r, done, threads = cpu_rate(args.workload, per_step)
        if i >= args.warmup:
            rates.append(r)
            total += done
            secs += done / r
    value = total / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.workload, args.walkers),
                   "step": f"bounded sample: {total // max(1, args.steps)} walkers of the workload per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{total} walkers in chunks of 64 over {args.steps} steps (torch.func.vmap float32 "
                                   f"port of the reference graph with dense Jacobians -- jax/jaxlib are not installable "
                                   f"offline; a lower bound on the reference's sparse-Jacobian JAX-CPU path)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled every 5 ms during the timed region through NVML (in-process:
    an `nvidia-smi -lms` child needs ~100 ms per sample, longer than a short timed region)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.index, self.sm, self.mask, self.h, self.nv, self.stop_flag = index, [], 0, None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def _sample(self):
        self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
        self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))

    def _loop(self):
        while not self.stop_flag:
            try:
                self._sample()
            except Exception:
                break
            time.sleep(0.005)

    def start(self):
        if self.nv is None:
            return
        self.stop_flag = False
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=1)
        try:
            mx = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        reasons = sorted(v for k, v in self.REASONS.items() if self.mask & k)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.sm)}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def parse_profile(rt):
    buf = ctypes.create_string_buffer(1 << 16)
    n = rt.lib.jaqmc_b200_profile_fetch(buf, len(buf))
    out = {}
    for ln in buf.raw[:n].decode().splitlines():
        name, cnt, ms, fl, by = ln.rsplit(None, 4)   # template kernels carry spaces in their name
        name = name.replace(" ", "").replace("(", "").replace(")", "")
        out[name] = {"launches": int(cnt), "ms": float(ms), "flops": float(fl), "bytes": float(by)}
    return out


def run_ours_solid(args, dev, world, rank, dist):
    """configs[4]: periodic FermiNet + Ewald local energy (complex log psi); resident and end-to-end timings."""
    import numpy as np
    from jaqmc_b200 import systems as H
    from jaqmc_b200.data import SolidData
    from jaqmc_b200.ewald import EwaldSum
    from jaqmc_b200.wavefunction import SolidWavefunction
    from jaqmc_b200._runtime import runtime

    mol, ndets, hs, hd, desc = WORKLOADS[args.workload][:5]
    prim, sim, patoms, cell_atoms, cell_charges, nspins, klist = H.solid_system(mol)
    n = sum(nspins)
    W = args.walkers
    if W % world:
        raise SystemExit(f"--walkers {W} not divisible by {world} ranks")
    Wl = W // world
    f32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).to(dev)  # noqa: E731
    wf = SolidWavefunction(nspins=nspins, simulation_lattice=sim, primitive_lattice=prim, klist=klist, ndets=ndets,
                           hidden_dims_single=list(hs), hidden_dims_double=list(hd))
    el_host = torch.from_numpy(H.solid_walkers(cell_atoms, n, W, seed=0)[rank * Wl:(rank + 1) * Wl]).contiguous().pin_memory()
    data = SolidData(electrons=el_host.to(dev), atoms=f32(cell_atoms), charges=f32(cell_charges), primitive_atoms=f32(patoms))
    params = wf.init_params(data, 42)
    ew = EwaldSum(sim, device=dev)
    rt = runtime(dev)
    sums = torch.zeros(3, device=dev)
    e_host = torch.empty(Wl, dtype=torch.float32).pin_memory()

    def step_resident():
        sums.zero_()
        out = wf.local_energy(params, data, ewald=ew, sums=sums)
        if dist:
            torch.distributed.all_reduce(sums)
        return out

    def step_e2e():
        data.electrons.copy_(el_host, non_blocking=True)
        out = step_resident()
        e_host.copy_(out["e_loc"].real, non_blocking=True)

    def timed(fn, k):
        if dist:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        if dist:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if dist:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t)
        return ms

    R = args.evals_per_step or default_evals_per_step(args.workload)

    def repeat(fn):
        def many():
            for _ in range(R):
                fn()
        return many

    for _ in range(max(3, args.warmup)):
        step_resident()
    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        clocks.start()
    rt.reset_launch_count()
    ms = timed(repeat(step_resident), args.steps)
    launches = rt.launch_count()
    clk = clocks.stop() if rank == 0 else None
    step_e2e()
    ms_e2e = timed(repeat(step_e2e), args.steps)
    finite = bool(torch.isfinite(e_host).all())
    kernels = None
    if rank == 0:
        rt.lib.jaqmc_b200_profile_enable(1)
        step_resident()
        torch.cuda.synchronize(dev)
        rt.lib.jaqmc_b200_profile_enable(0)
        prof = parse_profile(rt)
        tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
        kernels = {}
        for k, v in prof.items():
            e = kernels.setdefault(k.split("@")[0], {"launches": 0, "ms": 0.0})
            e["launches"] += v["launches"]
            e["ms"] += v["ms"]
        kernels = {k: {"launches": v["launches"], "share": round(v["ms"] / tot_ms, 4),
                       "ms_per_launch": round(v["ms"] / v["launches"], 4)} for k, v in kernels.items()}
        line = {
            "metric": METRIC, "value": round(W * R * args.steps / (ms * 1e-3), 1), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms / args.steps, 4),
            "ms_per_eval": round(ms / args.steps / R, 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 / complex64", "data": "synthetic",
            "config": {"workload": workload_string(args.workload, W), "walkers_per_gpu": Wl, "evals_per_step": R,
                       "step": f"{R} consecutive evaluations of the {W}-walker batch",
                       "parallelism": f"walkers sharded over {world} GPU(s); all-reduce of 3 floats per evaluation",
                       "launch": "kernel by kernel"},
            "e2e": {"value": round(W * R * args.steps / (ms_e2e * 1e-3), 1), "unit": UNIT,
                    "h2d_bytes_per_step": int(Wl * n * 3 * 4) * world * R, "d2h_bytes_per_step": int(Wl * 4) * world * R,
                    "finite": finite},
            "gpu_launches": int(launches), "clocks": clk, "roofline": None, "kernels": kernels,
        }
        print(json.dumps(line), flush=True)
    if dist:
        torch.distributed.destroy_process_group()


def run_ours(args):
    from jaqmc_b200 import systems as H
    from jaqmc_b200.data import MoleculeData
    from jaqmc_b200.sampler import MCMCSampler, SamplePlan
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = world > 1
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if dist:
        torch.distributed.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    mol, ndets, hs, hd, desc = WORKLOADS[args.workload][:5]
    if workload_kind(args.workload) == "solid":
        return run_ours_solid(args, dev, world, rank, dist)
    atoms64, charges64, nspins = H.molecule(mol)
    n = sum(nspins)
    W = args.walkers
    if W % world:
        raise SystemExit(f"--walkers {W} not divisible by {world} ranks")
    Wl = W // world
    wf = make_wavefunction(args.workload, nspins)
    atoms, charges = atoms64.float().to(dev), charges64.float().to(dev)
    el_all = H.synthetic_walkers(atoms64, charges64, nspins, W, seed=0).float()
    el_host = el_all[rank * Wl:(rank + 1) * Wl].contiguous().pin_memory()
    data = MoleculeData(electrons=el_host.to(dev), atoms=atoms, charges=charges)
    params = wf.init_params(data, 42)
    # equilibrate a little so that no walker sits at a pathological random position
    plan = SamplePlan(wf, MCMCSampler(steps=10), graph=not args.no_graph)
    st = plan.init(data)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    for _ in range(0 if args.no_equilibrate else 2):
        data, _, st = plan.step(params, data, st, gen)
    el_host.copy_(data.electrons.cpu())
    from jaqmc_b200._runtime import runtime
    rt = runtime(dev)
    sums = torch.zeros(3, device=dev)

    from jaqmc_b200.wavefunction import capture_local_energy

    use_graph = not args.no_graph
    if use_graph:
        replay, graph_out = capture_local_energy(wf, params, data, sums=sums)

    def step_resident():
        sums.zero_()
        out = replay() if use_graph else wf.local_energy(params, data, sums=sums)
        if dist:
            torch.distributed.all_reduce(sums)  # the only cross-GPU traffic: energy statistics (3 floats)
        return out

    e_host = torch.empty(Wl, dtype=torch.float32).pin_memory()
    el_dev = torch.empty_like(data.electrons)

    def step_e2e():
        # public API with host buffers: pinned H2D of the walkers into the captured input, graph replay, D2H of E_L
        sums.zero_()
        if use_graph:
            data.electrons.copy_(el_host, non_blocking=True)
            out = replay()
        else:
            el_dev.copy_(el_host, non_blocking=True)
            out = wf.local_energy(params, MoleculeData(el_dev, atoms, charges), sums=sums)
        if dist:
            torch.distributed.all_reduce(sums)
        e_host.copy_(out["e_loc"], non_blocking=True)
        return out

    def barrier():
        if dist:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t)
        return ms

    R = args.evals_per_step or default_evals_per_step(args.workload)   # evaluations of the batch per step

    def repeat(fn):
        def many():
            for _ in range(R):
                fn()
        return many

    for _ in range(max(3, args.warmup)):
        repeat(step_resident)()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    rt.reset_launch_count()
    ms = timed(repeat(step_resident), args.steps)
    launches = rt.launch_count()
    if use_graph:  # replays do not pass through the launcher: count the kernels of one captured evaluation
        rt.reset_launch_count()
        wf.local_energy(params, data, sums=sums)
        launches = rt.launch_count() * args.steps * R
    clk = clocks.stop() if rank == 0 else None
    value = W * R * args.steps / (ms * 1e-3)

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(repeat(step_e2e), args.steps)
    e2e_value = W * R * args.steps / (ms_e2e * 1e-3)
    finite = bool(torch.isfinite(e_host).all())

    # second half of the BASELINE metric: sampling + energy part of one VMC iteration (workflow/stage/vmc.py:227-265):
    # 10 MH sub-steps (11 value-only forward passes) followed by one local-energy evaluation; gradients and the
    # optimizer stay in the reference's JAX code and are not part of this number
    vmc_data = data   # the walkers of the sampling chain; `data.electrons` stays the captured graph input

    def vmc_iteration():
        nonlocal vmc_data, st
        vmc_data, _, st = plan.step(params, vmc_data, st, gen)
        sums.zero_()
        if use_graph:
            data.electrons.copy_(vmc_data.electrons)
            replay()
        else:
            wf.local_energy(params, vmc_data, sums=sums)
        if dist:
            torch.distributed.all_reduce(sums)

    vmc = None
    if not args.no_vmc:
        for _ in range(2):
            vmc_iteration()
        k_vmc = min(args.steps, 5)
        ms_vmc = timed(vmc_iteration, k_vmc)
        vmc = {"iters_per_sec": round(k_vmc / (ms_vmc * 1e-3), 2), "ms_per_iter": round(ms_vmc / k_vmc, 3),
               "what": "10 MH sub-steps + 1 forward-Laplacian local-energy evaluation of all walkers "
                       "(no parameter gradients / optimizer)"}
        if workload_kind(args.workload) == "ferminet" and n <= 16:
            # + LossAndGrad (estimator/loss_grad.py:70-128): MAD clipping, one reverse pass, all-reduce of the gradient --
            # everything of VMCWorkStage.compute_step (workflow/stage/vmc.py:227-265) except the optimizer update
            from jaqmc_b200.estimator import LossAndGrad

            lag = LossAndGrad(f_log_psi=wf)

            def vmc_iteration_grad():
                vmc_iteration()
                e = graph_out["e_loc"] if use_graph else wf.local_energy(params, vmc_data)["e_loc"]
                return lag.evaluate(params, vmc_data, {"total_energy": e})

            for _ in range(2):
                vmc_iteration_grad()
            ms_g = timed(vmc_iteration_grad, k_vmc)
            vmc["with_loss_and_grad"] = {
                "iters_per_sec": round(k_vmc / (ms_g * 1e-3), 2), "ms_per_iter": round(ms_g / k_vmc, 3),
                "what": "the above + LossAndGrad (clipped-energy-weighted parameter gradient by one reverse pass, "
                        "all-reduced over the GPUs): VMCWorkStage.compute_step without the optimizer update"}

    # roofline leg: per-kernel CUDA events on the launch stream over the same steps (rank 0 only)
    roof, kernels = None, None
    if rank == 0:
        # two untimed evaluations first (eager launches after a graph-replay loop: the first ones run at a different
        # clock / cache state; r2 saw the dominant kernel's mean move by 9 % between three-evaluation passes), then
        # n_prof evaluations with an event pair around every launch
        n_prof = 8 if args.steps >= 3 else max(1, args.steps)
        for _ in range(2 if args.steps >= 3 else 0):
            wf.local_energy(params, data, sums=sums)
        torch.cuda.synchronize(dev)
        rt.lib.jaqmc_b200_profile_enable(1)
        for _ in range(n_prof):
            wf.local_energy(params, data, sums=sums)
        torch.cuda.synchronize(dev)
        rt.lib.jaqmc_b200_profile_enable(0)
        prof = parse_profile(rt)   # keyed "kernel@declared-work": one entry per launch shape
        tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
        kernels = {}
        for k, v in prof.items():
            nm = k.split("@")[0]
            e = kernels.setdefault(nm, {"launches": 0, "ms": 0.0})
            e["launches"] += v["launches"]
            e["ms"] += v["ms"]
        kernels = {k: {"launches": v["launches"], "share": round(v["ms"] / tot_ms, 4),
                       "ms_per_launch": round(v["ms"] / v["launches"], 4)} for k, v in kernels.items()}
        name, v = max(prof.items(), key=lambda kv: kv[1]["ms"])   # dominant launch shape
        peaks, src = load_peaks()
        kname = name.split("@")[0]
        per_launch_ms = v["ms"] / v["launches"]
        # DRAM bytes of the same launch from an `ncu --set full` capture (profiles/ncu_traffic.json, per launch)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                # the capture is keyed by the kernel's base name (template arguments of the instantiation dropped)
                traffic = json.load(f).get(f"{kname.split('<')[0]}@{args.workload}@{W // world}", {}).get("dram_bytes")
        tensor_kernel = kname.startswith("k_dense_tc")
        if v["flops"] > 0 and tensor_kernel:
            # 3xTF32: three tensor-core products per multiply-add; TF32 runs at half the bf16 rate.  The kernel is timed
            # launch by launch with CUDA events (a burst, not a long saturated run): the burst figure is the denominator;
            # the fraction of the sustained figure is given next to it
            peak = peaks["bf16_tflops"] / 2.0 / 3.0
            peak_sus = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / 2.0 / 3.0
            ach = v["flops"] / (v["ms"] * 1e-3) / 1e12
            note = ("executed FLOPs of the restructured FermiNet layer (320-wide contraction + per-walker addend), not "
                    "the reference's 832-wide formulation" if workload_kind(args.workload) == "ferminet" else
                    "executed FLOPs of one Dense layer over all (value, Jacobian, Laplacian) rows")
            roof = {"kernel": kname, "bound": "tensor", "achieved": round(ach, 3), "peak": round(peak, 1),
                    "unit": "TFLOP/s", "frac": round(ach / peak, 4), "frac_burst": round(ach / peak, 4),
                    "frac_sustained": round(ach / peak_sus, 4), "traffic": traffic,
                    "peak_source": f"{src}: bf16_tflops (burst) / 2 (tf32) / 3 (split products); frac_sustained uses "
                                   f"bf16_tflops_sustained",
                    "launches_per_step": v["launches"] // n_prof, "profiled_evaluations": n_prof,
                    "ms_per_launch": round(per_launch_ms, 4),
                    "flops_per_launch": v["flops"] / v["launches"], "bytes_per_launch": v["bytes"] / v["launches"],
                    "hbm_gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1),
                    "share_of_step": round(v["ms"] / tot_ms, 4), "note": note}
        elif v["flops"] > 0:
            # CUDA-core FP32 contraction (attention / LogDet kernels): against the split-precision tensor peak, which
            # is where the north star wants these contractions
            peak = peaks["bf16_tflops"] / 2.0 / 3.0
            ach = v["flops"] / (v["ms"] * 1e-3) / 1e12
            roof = {"kernel": kname, "bound": "tensor", "achieved": round(ach, 3), "peak": round(peak, 1),
                    "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": traffic,
                    "peak_source": f"{src}: bf16_tflops / 2 / 3", "ms_per_launch": round(per_launch_ms, 4),
                    "flops_per_launch": v["flops"] / v["launches"], "share_of_step": round(v["ms"] / tot_ms, 4),
                    "note": "FP32 CUDA-core kernel measured against the split-precision tensor-core peak"}
        else:
            ach = v["bytes"] / (v["ms"] * 1e-3) / 1e9
            roof = {"kernel": kname, "bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": round(ach / peaks["hbm_gbs"], 4), "traffic": traffic, "peak_source": src,
                    "ms_per_launch": round(per_launch_ms, 4), "bytes_per_launch": v["bytes"] / v["launches"],
                    "share_of_step": round(v["ms"] / tot_ms, 4)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r, done, threads = cpu_rate(args.workload, args.cpu_seconds)
        cpu = {"value": round(r, 2), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{done} walkers of the same workload in chunks of 64 (float32 torch.func.vmap port of the "
                         f"reference graph), scaled per walker"}

    if rank == 0:
        ws_gb = rt.workspace_bytes(wf._handle(params, atoms.shape[0]), Wl, True) / 1e9
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms / args.steps, 4),
            "ms_per_eval": round(ms / args.steps / R, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.workload, W), "walkers_per_gpu": Wl, "evals_per_step": R,
                       "step": f"{R} consecutive evaluations of the {W}-walker batch",
                       "l2": f"no flush: each evaluation streams a {ws_gb:.1f} GB working set (>> 126 MB L2) per GPU",
                       "parallelism": f"walkers sharded over {world} GPU(s); all-reduce of 3 floats per evaluation",
                       "launch": "one CUDA-graph replay per evaluation" if use_graph else "kernel by kernel"},
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": int(Wl * n * 3 * 4) * world * R,
                    "d2h_bytes_per_step": int(Wl * 4) * world * R, "finite": finite},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "kernels": kernels,
        }
        if vmc is not None:
            line["vmc_iteration"] = vmc
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="n2", choices=sorted(WORKLOADS))
    ap.add_argument("--walkers", type=int, default=4096, help="global walker batch (reference workflow.batch_size)")
    ap.add_argument("--evals-per-step", type=int, default=0,
                    help="evaluations of the walker batch per timed step (0: 32 for FermiNet workloads, 4 otherwise)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--no-vmc", action="store_true", help="skip the MH + energy iteration timing")
    ap.add_argument("--no-equilibrate", action="store_true", help="skip the MH sweeps before timing (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
