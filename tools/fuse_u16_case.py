"""BASELINE configs[1] geometry with uint16 views (the production dtype): fused kernel time on device-resident views."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mvregfus_b200 import fusion
dev = torch.device("cuda:0")
for dt in ("u16", "f32"):
    views, params, props, sp = bench.make_workload(dev, 0, dtype=dt)
    dims = [int(s) for s in sp["size"]]
    vs = fusion.build_viewset(views, props, params, sp, weights="blending")
    out = fusion.fuse_blend(vs, dims)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        out = fusion.fuse_blend(vs, dims)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    nvox = int(np.prod(dims))
    alg = sum(int(v.numel()) * v.element_size() for v in views) + nvox * 2
    print("%s views: %.3f ms, %.1f Gvoxel/s, algorithmic %.0f GB/s" % (dt, ms, nvox / ms / 1e6, alg / ms / 1e6))
    del views, vs, out
    torch.cuda.empty_cache()
