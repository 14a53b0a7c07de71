"""N > 1 host logic on CPU: two gloo ranks shard the walkers like the reference's 1-D mesh (data.py:181-205), each
evaluates its block (host emulation build), and the cross-rank reduction of the energy statistics
(estimator/base.py:27-53 pmean) reproduces the single-process result."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helpers as H
    from jaqmc_b200 import _marshal as M
    from jaqmc_b200.data import BatchedData, MoleculeData
    from jaqmc_b200.estimator import mean_reduce
    from oracle import networks as ON

    rt = H.emu_runtime()
    atoms, charges, nspins = H.molecule("LiH")
    hs, hd, ndets = (16, 16), (8, 8), 2
    p64 = H.round_f32(ON.init_ferminet_params(nspins, atoms.shape[0], ndets, hs, hd, seed=3))
    el = H.synthetic_walkers(atoms, charges, nspins, 8, seed=5).float()
    full = BatchedData(MoleculeData(el, atoms.float(), charges.float()))
    mine = full.shard(rank, world)
    assert mine.batch_size == 8 // world
    wf = M.ferminet_handle(H.to_f32(p64), nspins, atoms.shape[0], ndets, hs, hd)
    sysh = M.system_handle(atoms.float(), charges.float())
    sums = torch.zeros(3)
    out = rt.local_energy(wf, sysh, mine.data.electrons.contiguous(), sums=sums)
    stats = mean_reduce({"total_energy": out["e_loc"]})   # all-reduce of the per-rank means (gloo)
    dist.all_reduce(sums)                                    # the 3-float partial sums the CUDA path emits
    if rank == 0:
        ref = rt.local_energy(wf, sysh, el.contiguous())["e_loc"]
        ret["mean"], ret["ref_mean"] = float(stats["total_energy"]), float(ref.mean())
        ret["var"], ret["ref_var"] = float(stats["total_energy_var"]), float((ref * ref).mean() - ref.mean() ** 2)
        ret["sums"], ret["ref_sums"] = sums.tolist(), [float(ref.sum()), float((ref * ref).sum()), 8.0]
    dist.destroy_process_group()


def test_two_rank_walker_sharding_and_reduction():
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert np.isclose(ret["mean"], ret["ref_mean"], rtol=1e-6)
    assert np.isclose(ret["var"], ret["ref_var"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ret["sums"], ret["ref_sums"], rtol=1e-5)
