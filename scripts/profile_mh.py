"""Times the MH sampling pass of a workload (value-only forward passes + fused accept) outside bench.py:
ms per call of ``jaqmc_b200_mh_step`` (S sub-steps, eager launches and CUDA-graph replay).  Run under
``ncu --metrics gpu__time_duration.sum`` for the per-kernel list of one pass (use --eager --calls 1)."""
import argparse
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from jaqmc_b200._runtime import runtime  # noqa: E402
from jaqmc_b200.systems import molecule, synthetic_walkers  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="n2")
    ap.add_argument("--walkers", type=int, default=4096)
    ap.add_argument("--substeps", type=int, default=10)
    ap.add_argument("--calls", type=int, default=10)
    ap.add_argument("--eager", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    rt = runtime(dev)
    mol = B.WORKLOADS[a.workload][0]
    atoms64, charges64, nspins = molecule(mol)
    wf = B.make_wavefunction(a.workload, nspins)
    from jaqmc_b200.data import MoleculeData
    atoms, charges = atoms64.float().to(dev), charges64.float().to(dev)
    el = synthetic_walkers(atoms64, charges64, nspins, a.walkers, seed=1).float().to(dev).contiguous()
    data = MoleculeData(electrons=el, atoms=atoms, charges=charges)
    params = wf.init_params(data, 42)
    h, sysh = wf._sampling_handles(params, data)
    S, W, n = a.substeps, a.walkers, el.shape[1]
    g = torch.Generator(device=dev).manual_seed(3)
    normals = torch.randn(S, W, n, 3, device=dev, generator=g)
    uniforms = torch.rand(S, W, device=dev, generator=g)
    stddev = torch.full((1,), 0.05, device=dev)
    logpsi = torch.empty(W, device=dev)
    rt.mh_step(h, sysh, el, logpsi, normals, uniforms, stddev, logpsi_valid=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if a.eager:
        rt.reset_launch_count()
        e0.record()
        for _ in range(a.calls):
            rt.mh_step(h, sysh, el, logpsi, normals, uniforms, stddev, logpsi_valid=False)
        e1.record()
        torch.cuda.synchronize()
        print("eager: %.3f ms per call of %d sub-steps, %d launches per call" % (e0.elapsed_time(e1) / a.calls, S, rt.launch_count() // a.calls))
        return
    replay = rt.capture_mh_step(h, sysh, el, normals, uniforms, stddev)
    for _ in range(3):
        replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.calls):
        replay()
    e1.record()
    torch.cuda.synchronize()
    print("graph: %.3f ms per call of %d sub-steps (%.3f ms per sub-step)" % (e0.elapsed_time(e1) / a.calls, S, e0.elapsed_time(e1) / a.calls / S))


if __name__ == "__main__":
    main()
