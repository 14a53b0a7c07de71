"""Per-kernel device times of one MH sampling pass (10 sub-steps, value-only forwards) on a bench workload."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench as B
import helpers as H
from jaqmc_b200.data import MoleculeData
from jaqmc_b200.sampler import MCMCSampler, SamplePlan
from jaqmc_b200._runtime import runtime

wl = sys.argv[1] if len(sys.argv) > 1 else "n2"
W = int(os.environ.get("WALKERS", 4096))
dev = torch.device("cuda", 0)
mol = B.WORKLOADS[wl][0]
atoms64, charges64, nspins = H.molecule(mol)
wf = B.make_wavefunction(wl, nspins)
el = H.synthetic_walkers(atoms64, charges64, nspins, W, seed=0).float().to(dev)
data = MoleculeData(electrons=el, atoms=atoms64.float().to(dev), charges=charges64.float().to(dev))
params = wf.init_params(data, 42)
plan = SamplePlan(wf, MCMCSampler(steps=10))
st = plan.init(data)
gen = torch.Generator(device=dev).manual_seed(1)
rt = runtime(dev)
for _ in range(2):
    data, _, st = plan.step(params, data, st, gen)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    data, _, st = plan.step(params, data, st, gen)
e1.record(); torch.cuda.synchronize()
print(f"plan.step (10 sub-steps): {e0.elapsed_time(e1) / 3:.3f} ms")
rt.lib.jaqmc_b200_profile_enable(1)
data, _, st = plan.step(params, data, st, gen)
torch.cuda.synchronize()
rt.lib.jaqmc_b200_profile_enable(0)
prof = B.parse_profile(rt)
tot = sum(v["ms"] for v in prof.values())
print(f"sum of kernel times {tot:.3f} ms over {sum(v['launches'] for v in prof.values())} launches")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:25]:
    print(f"{v['ms']:8.3f} ms  {v['launches']:4d} x {v['ms'] / v['launches'] * 1e3:8.1f} us  {k[:90]}")
