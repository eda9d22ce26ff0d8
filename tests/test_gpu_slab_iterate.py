"""SlabRLPlan.iterate() itself - the code path bench.py times at N > 1 - against the whole-volume run.

Two ranks as two processes.  With one visible GPU both ranks share it and talk over gloo (ring isend/irecv staged
through host memory): that drives iterate()'s own-planes-first / halo-planes-later split of the x/y passes and its
request handling with real kernels.  With two or more GPUs the ranks get one GPU each and use NCCL plus the
symmetric-memory halo pull and the CUDA graph replay.  Reference: mvregfus/multiview.py:5678-5727 (one iteration)."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case():
    import torch
    g = torch.Generator()
    g.manual_seed(5)
    shape, V = (96, 40, 64), 3
    views = [(torch.rand(shape, generator=g) * 3000).to(torch.int32).to(torch.uint16) for _ in range(V)]
    w = [torch.rand(shape, generator=g) for _ in range(V)]
    for x in w:
        x[x < 0.05] = 0.0     # exercise the (w > 1e-5) mask
    k = np.exp(-0.5 * (np.arange(-6, 7) / 3.0) ** 2)
    kz = (k / k.sum()).astype(np.float32)
    k = np.exp(-0.5 * (np.arange(-6, 7) / 1.5) ** 2)
    kxy = (k / k.sum()).astype(np.float32)
    psfs = [kz[:, None, None] * kxy[None, :, None] * kxy[None, None, :]] * V
    return views, w, psfs


def _rank(rank, world, port, backend, halo, out_dir, iters):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), MVF_HALO=halo)
    import torch
    import torch.distributed as dist
    from mvregfus_b200 import rl, slab
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    views, w, psfs = _case()
    nz = views[0].shape[0]
    z0, z1 = slab.slab_bounds(nz, world)[rank]
    plan = rl.SlabRLPlan([v[z0:z1].to(dev) for v in views], [x[z0:z1].to(dev) for x in w], psfs, nz, rank, world)
    plan.init_estimate()
    for _ in range(iters):
        plan.iterate()
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, "est_%d.npy" % rank), plan.est.cpu().numpy())
    with open(os.path.join(out_dir, "mode_%d.txt" % rank), "w") as fh:
        fh.write("symm" if plan.symm is not None else "ring")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("iters", [4])
def test_slab_iterate_two_ranks_equals_whole(cuda, iters):
    import torch
    import torch.multiprocessing as mp
    from mvregfus_b200 import rl
    two_gpus = torch.cuda.device_count() >= 2
    backend, halo = ("nccl", "symm") if two_gpus else ("gloo", "nccl")
    views, w, psfs = _case()
    plan = rl.RLPlan([v.cuda() for v in views], [x.cuda() for x in w], psfs)
    plan.init_estimate()
    for _ in range(iters):
        plan.iterate()
    whole = plan.est.cpu().numpy()
    with tempfile.TemporaryDirectory() as d:
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_rank, args=(2, port, backend, halo, d, iters), nprocs=2, join=True)
        got = np.concatenate([np.load(os.path.join(d, "est_%d.npy" % r)) for r in range(2)], 0)
        modes = [open(os.path.join(d, "mode_%d.txt" % r)).read() for r in range(2)]
    assert modes == (["symm", "symm"] if two_gpus else ["ring", "ring"])
    assert np.array_equal(got, whole)   # same kernels, same accumulation order: bit for bit
