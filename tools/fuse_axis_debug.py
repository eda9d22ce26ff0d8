"""Locate differences between the axis-aligned fused kernel and the general kernel. python tools/fuse_axis_debug.py N [u16|f32]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvregfus_b200 import fusion, synthetic
from mvregfus_b200 import multiview as mv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dt = np.uint16 if (len(sys.argv) > 2 and sys.argv[2] == "u16") else np.float32
shape = (n, n, n)
views, params, props = synthetic.make_views(shape, 4, dtype=dt, seed=1000)
sp = mv.calc_stack_properties_from_views_and_params(props, params, spacing=[1.0] * 3)
dims = [int(s) for s in sp["size"]]
dv = [fusion.to_device(v) for v in views]
vs = fusion.build_viewset(dv, props, params, sp, weights="blending")
res = {}
for name in ("axis", "general"):
    if name == "general":
        os.environ["MVF_FUSE_KERNEL"] = "general"
    else:
        os.environ.pop("MVF_FUSE_KERNEL", None)
    res[name] = fusion.fuse_blend(vs, dims).cpu().numpy().astype(np.int64)
d = np.abs(res["axis"] - res["general"])
bad = np.argwhere(d > 2)
print("dims", dims, "max", d.max(), "n_bad(>2)", len(bad), "of", d.size)
if len(bad):
    tz, ty, tx = bad[:, 0] // 16, bad[:, 1] // 8, bad[:, 2] // 32
    tiles = sorted(set(zip(tz.tolist(), ty.tolist(), tx.tolist())))
    print("bad tiles:", len(tiles), tiles[:20])
    print("k within tile (z%16) histogram:", np.bincount(bad[:, 0] % 16, minlength=16))
    print("first bad voxels:", bad[:8].tolist(), [(int(res['axis'][tuple(b)]), int(res['general'][tuple(b)])) for b in bad[:8]])
