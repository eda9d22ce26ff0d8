"""Where the time of the drop-in e2e path goes (configs[1], host numpy in page-locked memory)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mvregfus_b200 import fusion, pipeline
from mvregfus_b200 import multiview as mv
from mvregfus_b200.image_array import ImageArray
dev = torch.device("cuda:0")
views, params, props, sp = bench.make_workload(dev, 0)
host = []
for v in views:
    h = fusion.pinned_empty(v.shape, np.float32); torch.from_numpy(h).copy_(v); host.append(ImageArray(h))
torch.cuda.synchronize()
def T(f, n=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
ms, dv = T(lambda: [fusion.to_device(h) for h in host]); print("H2D 4 views: %.1f ms" % ms)
ms, fused = T(lambda: pipeline.fuse_volume_blockwise(dv, props, params, sp)); print("fuse_volume_blockwise (device): %.1f ms" % ms)
ms, _ = T(lambda: fusion.to_host(fused)); print("D2H uint16: %.1f ms" % ms)
def e2e():
    tv = [mv.transform_stack_dask_numpy(h, params[i], stack_properties=sp) for i, h in enumerate(host)]
    return mv.fuse_blockwise_arrays(tv, params, sp, props)
ms, _ = T(e2e); print("e2e: %.1f ms" % ms)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); e2e(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
