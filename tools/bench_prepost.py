"""Streaming kernels either side of the path on full-size volumes: achieved HBM bandwidth (CUDA events)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mvregfus_b200 import _cabi, prepost
from mvregfus_b200.fusion import _stream_ptr

PEAK = 6556.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
lib = _cabi.load()
shape = (1024, 2048, 2048) if "--c5" in sys.argv else (512, 1024, 1024)
g = torch.Generator(device="cuda"); g.manual_seed(3)
vol = (torch.rand(shape, generator=g, device="cuda") ** 3 * 4000).to(torch.int32).to(torch.uint16)
n = vol.numel()


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rows = []
out = torch.empty_like(vol)
ms = timed(lambda: _cabi.check(lib.mvf_background_u16(C.c_void_p(vol.data_ptr()), C.c_void_p(out.data_ptr()), n, 200, _stream_ptr())))
rows.append(("background_u16", 4 * n, ms))
mm = torch.empty(2, dtype=torch.int32, device="cuda")
ms = timed(lambda: _cabi.check(lib.mvf_range_u16(C.c_void_p(vol.data_ptr()), n, C.c_void_p(mm.data_ptr()), _stream_ptr())))
rows.append(("range_u16", 2 * n, ms))
lo, hi = (int(x) for x in mm.cpu().numpy())
lut = torch.from_numpy(prepost.value_lut(lo, hi)[0]).cuda()
cnt = torch.empty(256, dtype=torch.int64, device="cuda")
ms = timed(lambda: _cabi.check(lib.mvf_hist_u16(C.c_void_p(vol.data_ptr()), n, C.c_void_p(lut.data_ptr()), C.c_void_p(cnt.data_ptr()), _stream_ptr())))
rows.append(("hist_u16", 2 * n, ms))
assert int(cnt.sum()) == n
sub = torch.empty([s // 2 for s in shape], dtype=torch.uint16, device="cuda")
ms = timed(lambda: _cabi.check(lib.mvf_subsample(C.c_void_p(vol.data_ptr()), 0, _cabi.i32x3(shape), _cabi.i32x3((2, 2, 2)), C.c_void_p(sub.data_ptr()), _stream_ptr())))
# strided read: every touched 32-byte sector of the even rows/planes is compulsory traffic (n/4 voxels' sectors), + the write
rows.append(("subsample 2x2x2", 2 * (n // 4) + 2 * (n // 8), ms))
ms = timed(lambda: _cabi.check(lib.mvf_bin_shrink(C.c_void_p(vol.data_ptr()), 0, _cabi.i32x3(shape), _cabi.i32x3((1, 2, 2)), C.c_void_p(sub.data_ptr()), _stream_ptr())))
rows.append(("bin_shrink 1x2x2", 2 * n + 2 * (n // 4), ms))
print("volume %s uint16, HBM peak %.1f GB/s" % (shape, PEAK))
for name, nbytes, ms in rows:
    gbs = nbytes / ms / 1e6
    print("%-18s %8.3f ms  %8.1f GB/s  %5.1f %% of peak" % (name, ms, gbs, 100 * gbs / PEAK))
