"""One reverse pass (jaqmc_b200_ferminet_logpsi_vjp) of FermiNet on a workload, timed with CUDA events; run under
``ncu --metrics gpu__time_duration.sum`` for the per-kernel list."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from jaqmc_b200.data import MoleculeData  # noqa: E402
from jaqmc_b200.systems import molecule, synthetic_walkers  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="n2")
    ap.add_argument("--walkers", type=int, default=4096)
    ap.add_argument("--calls", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    atoms64, charges64, nspins = molecule(B.WORKLOADS[a.workload][0])
    wf = B.make_wavefunction(a.workload, nspins)
    el = synthetic_walkers(atoms64, charges64, nspins, a.walkers, seed=1).float().to(dev).contiguous()
    data = MoleculeData(electrons=el, atoms=atoms64.float().to(dev), charges=charges64.float().to(dev))
    params = wf.init_params(data, 42)
    wgt = torch.randn(a.walkers, device=dev) / a.walkers
    wf.logpsi_vjp(params, data, wgt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.calls):
        wf.logpsi_vjp(params, data, wgt)
    e1.record()
    torch.cuda.synchronize()
    print("logpsi_vjp: %.3f ms per call (%d walkers)" % (e0.elapsed_time(e1) / a.calls, a.walkers))


if __name__ == "__main__":
    main()
