"""Measured parity numbers at BASELINE.json configs[1] size (4 views 512^3), GPU vs oracle on sub-boxes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mvregfus_b200 import fusion
from oracle import mvregfus_oracle as O

for dtype in ("f32", "u16"):
    views, params, props, sp = bench.make_workload(torch.device("cuda:0"), 0, dtype=dtype)
    dims = [int(s) for s in sp["size"]]
    vs = fusion.build_viewset(views, props, params, sp, weights="blending")
    outf = torch.empty(dims, dtype=torch.float32, device="cuda")
    out = fusion.fuse_blend(vs, dims, out_f32=outf).cpu().numpy(); outf = outf.cpu().numpy()
    host = [v.cpu().numpy() for v in views]
    maps = [O.index_space_affine(params[v], props[v]["spacing"], props[v]["origin"], sp["spacing"], sp["origin"]) for v in range(4)]
    box = 64; n = 0; mism = 0; mx = 0; relf = 0.0; nflip = 0
    for corner in ([60, 70, 80], [230, 240, 250], [440, 200, 430], [100, 400, 300], [300, 100, 150], [5, 5, 5]):
        off = np.array(corner)
        tv = [O.transform_chunk(host[v], maps[v][0], maps[v][1], off, [box] * 3) for v in range(4)]
        bp = {"size": np.array([box] * 3), "spacing": sp["spacing"], "origin": sp["origin"] + off * sp["spacing"]}
        w = O.get_weights_simple(props, params, bp)
        ref = np.asarray(O.fuse_views_weights(tv, params, bp, weights=w)); ref_f = O.fuse_views_weights_float(tv, w)
        sl = tuple(slice(int(o), int(o) + box) for o in off)
        d = np.abs(out[sl].astype(np.int64) - ref.astype(np.int64))
        n += d.size; mism += int((d > 0).sum()); mx = max(mx, int(d.max()))
        rel = np.abs(outf[sl] - ref_f) / np.maximum(np.abs(ref_f), 1.0)
        if dtype == "f32":
            relf = max(relf, float(rel.max()))
        else:   # uint16 views are rounded per view before weighting: a flipped rounding moves the sum by w_v (<= 1)
            nflip += int((rel > 1e-5).sum())
            relf = max(relf, float(np.max(np.abs(outf[sl] - ref_f))))
    if dtype == "f32":
        print("f32: %d voxels compared, uint16 result: %d mismatches (%.2e), max |d| %d; float32 pre-truncation sum: max rel err %.2e"
              % (n, mism, mism / n, mx, relf))
    else:
        print("u16: %d voxels compared, uint16 result: %d mismatches (%.2e), max |d| %d; float32 pre-truncation sum: %d voxels (%.2e) "
              "beyond 1e-5 (a view's value rounded to the other integer), max |d| %.3f" % (n, mism, mism / n, mx, nflip, nflip / n, relf))
    del views, vs, outf
    torch.cuda.empty_cache()
