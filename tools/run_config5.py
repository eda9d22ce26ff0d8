"""BASELINE.json configs[4] end to end on N GPUs (torchrun): 8 views (512,1024,1024) uint16 fused to
(1024,2048,2048) at spacing 0.5 with blending weights, then 10 iterations of multi-view RL (sigma z=6, xy=2 ->
37^3 PSF at spacing 0.5) on z-slabs with ring halo exchange over NCCL.  Even views are axis aligned (separable
kernel), odd views are 7 degrees off axis (cuFFT path).  Prints one JSON object on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvregfus_b200 import fusion, rl, slab  # noqa: E402
from mvregfus_b200 import multiview as mv  # noqa: E402
from tools.run_configs import V, make_views, timed  # noqa: E402

ITERS = int(os.environ.get("C5_ITERS", "10"))


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    views, params, props = make_views(dev)
    fused = (1024, 2048, 2048)
    sp = {"size": np.array(fused), "spacing": np.array([0.5] * 3), "origin": np.zeros(3)}
    z0, z1 = slab.slab_bounds(fused[0], world)[rank]
    dims = [z1 - z0, fused[1], fused[2]]
    # fusion: one fused pass per rank on its slab
    vs = fusion.build_viewset(views, props, params, sp, weights="blending")
    out = torch.empty(dims, dtype=torch.uint16, device=dev)
    fuse_ms = timed(lambda: fusion.fuse_blend(vs, dims, origin=(z0, 0, 0), out=out), steps=3, warm=1)
    del out
    # RL inputs: transformed views and blending weights of the slab
    tv = []
    for v in range(V):
        M, off = mv.index_space_affine(params[v], props[v]["spacing"], props[v]["origin"], sp["spacing"], sp["origin"])
        tv.append(fusion.affine_resample(views[v], M, off, dims, (z0, 0, 0), rule="scipy"))
    w = fusion.blend_weights(fusion.build_viewset(None, props, params, sp, weights="blending"), dims, origin=(z0, 0, 0))
    del views
    torch.cuda.empty_cache()
    psfs = [mv.get_psf(params[v], sp, 6.0, 2.0) for v in range(V)]
    plan = rl.SlabRLPlan(tv, [w[v] for v in range(V)], psfs, fused[0], rank, world)
    sep = [s is None for s in plan.spectra]
    plan.init_estimate()
    plan.iterate()
    torch.cuda.synchronize()
    plan.init_estimate()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(ITERS):
        plan.iterate()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / ITERS], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    rl_ms = float(t.item())
    if rank == 0:
        nvox = int(np.prod(fused))
        peak = 6556.5 * world
        print(json.dumps({
            "config": "configs[4]: 8 views 512x1024x1024 u16 -> 1024x2048x2048 (spacing 0.5) + %d RL iterations, %d GPUs" % (ITERS, world),
            "fuse_ms": fuse_ms, "fused_gvoxel_per_s": nvox / fuse_ms / 1e6,
            "fuse_roofline_frac": (8 * 512 * 1024 * 1024 * 2 + nvox * 2) / (fuse_ms * 1e-3) / 1e9 / peak,
            "rl_ms_per_iteration": rl_ms, "rl_iterations_per_s": 1e3 / rl_ms,
            "rl_roofline_frac": (12 + 8 * 18) * nvox / (rl_ms * 1e-3) / 1e9 / peak,
            "psf": list(psfs[0].shape), "separable_views": sep, "slab_planes": dims[0]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
