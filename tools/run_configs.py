"""BASELINE.json configs[2] and configs[4] (fusion part) on N GPUs, z-slab sharded.  Launch with torchrun.

  configs[2]: 8 views (512,1024,1024) uint16 -> fused (512,1024,1024), DCT-entropy quality weights
  configs[4]: 8 views (512,1024,1024) uint16 -> fused (1024,2048,2048) at spacing 0.5, blending weights
Every rank holds all views (8.6 GB) and produces its z-slab of the fused frame; no data-path collective.
Prints one JSON object per config on rank 0 (max-over-ranks CUDA-event time)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvregfus_b200 import dct, fusion, slab, synthetic  # noqa: E402
from mvregfus_b200 import multiview as mv  # noqa: E402

V, VSHAPE = 8, (512, 1024, 1024)


def make_views(dev):
    g = torch.Generator(device=dev)
    g.manual_seed(3003)
    import torch.nn.functional as F
    fine = torch.rand((1, 1) + tuple(s // 16 for s in VSHAPE), generator=g, device=dev)
    gt = F.interpolate(fine, size=VSHAPE, mode="trilinear", align_corners=True)[0, 0]
    ax = [torch.linspace(-1, 1, s, device=dev) for s in VSHAPE]
    r2 = (ax[0][:, None, None] / 0.9) ** 2 + (ax[1][None, :, None] / 0.9) ** 2 + (ax[2][None, None, :] / 0.9) ** 2
    gt = (gt * 3000.0 * (r2 <= 1.0)).contiguous()
    params = synthetic.view_params(VSHAPE, V, seed=3003, odd_extra_deg=7.0)
    props = [{"size": np.array(VSHAPE), "spacing": np.ones(3), "origin": np.zeros(3)} for _ in range(V)]
    views = []
    for v in range(V):
        A, c = params[v][:9].reshape(3, 3), params[v][9:]
        Ainv = np.linalg.inv(A)
        img = fusion.affine_resample(gt, Ainv, -Ainv @ c, VSHAPE, rule="scipy")
        views.append(img.clamp_(0, 4095).round_().to(torch.int32).to(torch.uint16))
    return views, params, props


def timed(fn, steps=5, warm=2):
    for _ in range(warm):
        fn()
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device="cuda", dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    views, params, props = make_views(dev)
    peak = 6556.5
    res = []
    # ---- configs[4] fusion part: fused (1024,2048,2048), spacing 0.5, blending
    sp5 = {"size": np.array([1024, 2048, 2048]), "spacing": np.array([0.5] * 3), "origin": np.zeros(3)}
    z0, z1 = slab.slab_bounds(1024, world)[rank]
    dims = [z1 - z0, 2048, 2048]
    vs = fusion.build_viewset(views, props, params, sp5, weights="blending")
    out = torch.empty(dims, dtype=torch.uint16, device=dev)
    ms = timed(lambda: fusion.fuse_blend(vs, dims, origin=(z0, 0, 0), out=out))
    nvox = 1024 * 2048 * 2048
    alg = sum(v.numel() * 2 for v in views) + nvox * 2
    res.append({"config": "configs[4] fusion: 8 views 512x1024x1024 u16 -> 1024x2048x2048 (spacing 0.5), blending", "n_gpus": world,
                "ms": ms, "gvoxel_per_s": nvox / ms / 1e6, "roofline_frac": alg / (ms * 1e-3) / 1e9 / (peak * world)})
    del out
    # ---- configs[2]: fused (512,1024,1024), spacing 1, DCT weights (quality grid from materialised transformed views)
    sp3 = {"size": np.array([512, 1024, 1024]), "spacing": np.ones(3), "origin": np.zeros(3)}
    z0, z1 = slab.slab_bounds(512, world)[rank]
    dims = [z1 - z0, 1024, 1024]
    tv = []
    for v in range(V):
        M, off = mv.index_space_affine(params[v], props[v]["spacing"], props[v]["origin"], sp3["spacing"], sp3["origin"])
        tv.append(fusion.affine_resample(views[v], M, off, (512, 1024, 1024), rule="scipy"))
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    ws, b, size = dct.quality_grid(tv, params, props, sp3)      # every rank computes the (tiny) grid: no exchange
    torch.cuda.synchronize()
    t_q = time.perf_counter() - t0
    t0 = time.perf_counter()
    coarse = dct.coarse_weights(ws, sp3, b, size)
    t_c = time.perf_counter() - t0
    del tv
    vs = fusion.build_viewset(views, props, params, sp3, weights="dct", coarse=coarse)
    out = torch.empty(dims, dtype=torch.uint16, device=dev)
    ms = timed(lambda: fusion.fuse_blend(vs, dims, origin=(z0, 0, 0), out=out))
    vsb = fusion.build_viewset(views, props, params, sp3, weights="blending")
    ms_blend = timed(lambda: fusion.fuse_blend(vsb, dims, origin=(z0, 0, 0), out=out))
    nvox = 512 * 1024 * 1024
    res.append({"config": "configs[2]: 8 views 512x1024x1024 u16, DCT-entropy weights, z-slabs", "n_gpus": world, "fuse_ms": ms,
                "gvoxel_per_s_fuse_kernel": nvox / ms / 1e6, "same_with_blending_only_ms": ms_blend, "quality_grid_s (K3 + host probes, whole volume)": t_q,
                "coarse_heuristics_s (host scipy)": t_c, "dct_block": [int(b), int(size)], "grid": list(ws.shape)})
    if rank == 0:
        for r in res:
            print(json.dumps(r))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
