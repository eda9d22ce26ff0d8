"""Timing + parity of the fused kernel on BASELINE.json configs[1] (4 x 512^3): axis-aligned (TMA) kernel vs the general
per-voxel kernel.  python tools/fuse_axis_case.py [f32|u16]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mvregfus_b200 import fusion  # noqa: E402

dt = sys.argv[1] if len(sys.argv) > 1 else "f32"
dev = torch.device("cuda:0")
views, params, props, sp = bench.make_workload(dev, 0, dtype=dt)
dims = [int(s) for s in sp["size"]]
vs = fusion.build_viewset(views, props, params, sp, weights="blending")
res = {}
for name in ("axis", "general"):
    if name == "general":
        os.environ["MVF_FUSE_KERNEL"] = "general"
    else:
        os.environ.pop("MVF_FUSE_KERNEL", None)
    out = torch.empty(dims, dtype=torch.uint16, device=dev)
    outf = torch.empty(dims, dtype=torch.float32, device=dev)
    fusion.fuse_blend(vs, dims, out=out, out_f32=outf)
    res[name] = (out.cpu().numpy().astype(np.int64), outf.cpu().numpy())
    for _ in range(3):
        fusion.fuse_blend(vs, dims, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fusion.fuse_blend(vs, dims, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    nb = sum(v.numel() * v.element_size() for v in views) + int(np.prod(dims)) * 2
    print("%s %s: %.3f ms  %.1f Gvox/s  %.0f GB/s algorithmic (%.1f %% of 6556.5)" %
          (dt, name, ms, np.prod(dims) / ms / 1e6, nb / ms / 1e6, nb / ms / 1e6 / 65.565))
d = np.abs(res["axis"][0] - res["general"][0])
f = np.abs(res["axis"][1] - res["general"][1]) / np.maximum(np.abs(res["general"][1]), 1.0)
print("axis vs general: uint16 max|d| %d, mismatch fraction %.3e; float32 rel err %.3e" % (d.max(), (d > 0).mean(), f.max()))
