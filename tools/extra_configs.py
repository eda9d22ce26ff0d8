"""BASELINE.json configs[2] and configs[4] as functions for bench.py (extra keys at N >= 2 / N = 8) and for the
stand-alone drivers tools/run_configs.py / tools/run_config5.py.  torch.distributed must already be initialised when
world > 1.  Every rank holds all views and produces its z-slab of the fused frame; rank 0 gets the result dict."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvregfus_b200 import dct, fusion, rl, slab, synthetic  # noqa: E402
from mvregfus_b200 import multiview as mv  # noqa: E402

V, VSHAPE = 8, (512, 1024, 1024)
PEAK = 6556.5


def make_views(dev):
    """8 uint16 views (512,1024,1024): 45-degree steps about y, every other view 7 degrees further off axis."""
    import torch.nn.functional as F
    g = torch.Generator(device=dev)
    g.manual_seed(3003)
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
    """ms per call, CUDA events, max over ranks."""
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


def config2(dev, rank, world, views=None, params=None, props=None):
    """configs[2]: 8 views fused into (512,1024,1024) at spacing 1 with DCT-entropy quality weights, z-slabs."""
    if views is None:
        views, params, props = make_views(dev)
    sp = {"size": np.array([512, 1024, 1024]), "spacing": np.ones(3), "origin": np.zeros(3)}
    z0, z1 = slab.slab_bounds(512, world)[rank]
    dims = [z1 - z0, 1024, 1024]
    tv = []
    for v in range(V):
        M, off = mv.index_space_affine(params[v], props[v]["spacing"], props[v]["origin"], sp["spacing"], sp["origin"])
        tv.append(fusion.affine_resample(views[v], M, off, (512, 1024, 1024), rule="scipy"))
    torch.cuda.synchronize()
    if rank == 0:   # rank 0 fits the coarse grid for everybody: it may use all host cores (worker processes started up front)
        from mvregfus_b200 import _hostpool
        os.environ["MVF_HOST_WORKERS"] = str(max(0, min(16, len(os.sched_getaffinity(0)) - 2)))
        if _hostpool.workers() >= 2:
            _hostpool._start(_hostpool.workers())
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    ws, b, size = dct.quality_grid(tv, params, props, sp)      # every rank computes the (tiny) grid: no exchange
    torch.cuda.synchronize()
    t_q = time.perf_counter() - t0
    t0 = time.perf_counter()
    box = [dct.coarse_weights(ws, sp, b, size) if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)   # a (V, 8, 16, 16) float64 grid + two 3-vectors
    coarse = box[0]
    t_c = time.perf_counter() - t0
    del tv
    vs = fusion.build_viewset(views, props, params, sp, weights="dct", coarse=coarse)
    out = torch.empty(dims, dtype=torch.uint16, device=dev)
    ms = timed(lambda: fusion.fuse_blend(vs, dims, origin=(z0, 0, 0), out=out))
    nvox = 512 * 1024 * 1024
    alg = sum(v.numel() * 2 for v in views) * 2 + nvox * 2     # B_F + B_D (views read once for the quality pass, once for fusion)
    total_s = t_q + t_c + ms * 1e-3
    return {"workload": "BASELINE.json configs[2]: 8 views 512x1024x1024 uint16, DCT-entropy quality weights, %d z-slab(s)" % world,
            "n_gpus": world, "fuse_kernel_ms": ms, "fuse_kernel_gvoxel_per_s": nvox / ms / 1e6,
            "quality_grid_s": t_q, "coarse_grid_heuristics_s": t_c, "end_to_end_s": total_s,
            "gvoxel_per_s": nvox / total_s / 1e9,
            "roofline_frac_fuse_kernel": (sum(v.numel() * 2 for v in views) + nvox * 2) / (ms * 1e-3) / 1e9 / (PEAK * world),
            "roofline_frac_end_to_end": alg / total_s / 1e9 / (PEAK * world),
            "dct_block": [int(b), int(size)], "quality_grid_shape": list(ws.shape)}


def config4(dev, rank, world, iters=10, views=None, params=None, props=None):
    """configs[4] (the north-star configuration): 8 views fused to (1024,2048,2048) at spacing 0.5 with blending weights,
    then `iters` multi-view RL iterations (sigma z=6, xy=2 -> 37^3 PSF) on z-slabs with PSF-radius halos."""
    if views is None:
        views, params, props = make_views(dev)
    fused = (1024, 2048, 2048)
    sp = {"size": np.array(fused), "spacing": np.array([0.5] * 3), "origin": np.zeros(3)}
    z0, z1 = slab.slab_bounds(fused[0], world)[rank]
    dims = [z1 - z0, fused[1], fused[2]]
    vs = fusion.build_viewset(views, props, params, sp, weights="blending")
    out = torch.empty(dims, dtype=torch.uint16, device=dev)
    fuse_ms = timed(lambda: fusion.fuse_blend(vs, dims, origin=(z0, 0, 0), out=out), steps=3, warm=1)
    del out
    tv = []
    for v in range(V):
        M, off = mv.index_space_affine(params[v], props[v]["spacing"], props[v]["origin"], sp["spacing"], sp["origin"])
        tv.append(fusion.affine_resample(views[v], M, off, dims, (z0, 0, 0), rule="scipy"))
    w = fusion.blend_weights(fusion.build_viewset(None, props, params, sp, weights="blending"), dims, origin=(z0, 0, 0))
    del views, vs
    torch.cuda.empty_cache()
    psfs = [mv.get_psf(params[v], sp, 6.0, 2.0) for v in range(V)]
    plan = rl.SlabRLPlan(tv, [w[v] for v in range(V)], psfs, fused[0], rank, world)
    sep = [s is None for s in plan.spectra]
    plan.init_estimate()
    for _ in range(2):
        plan.iterate()
    torch.cuda.synchronize()
    plan.init_estimate()
    rl_ms = timed(plan.iterate, steps=iters, warm=0)
    nvox = int(np.prod(fused))
    peak = PEAK * world
    return {"workload": "BASELINE.json configs[4]: 8 views 512x1024x1024 uint16 -> 1024x2048x2048 at spacing 0.5 (blending), "
                        "then %d RL iterations with a 37^3 PSF, %d z-slab(s)" % (iters, world),
            "n_gpus": world, "fuse_ms": fuse_ms, "fused_gvoxel_per_s": nvox / fuse_ms / 1e6,
            "fuse_roofline_frac": (8 * 512 * 1024 * 1024 * 2 + nvox * 2) / (fuse_ms * 1e-3) / 1e9 / peak,
            "rl_ms_per_iteration": rl_ms, "rl_iterations_per_s": 1e3 / rl_ms,
            "rl_roofline_frac": (12 + 8 * 18) * nvox / (rl_ms * 1e-3) / 1e9 / peak,
            "psf": list(psfs[0].shape), "separable_views": sep, "slab_planes": dims[0],
            "halo": "symmetric-memory pull + CUDA graph" if plan.symm is not None else "NCCL ring (oblique views use the cuFFT blur)"}
