"""Small driver for profiling one RL iteration (4 views, 512x1024x1024, 19^3 separable PSF)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mvregfus_b200 import rl
shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (512, 1024, 1024)
V = 4
g = torch.Generator(device="cuda"); g.manual_seed(1)
views = [(torch.rand(shape, generator=g, device="cuda") * 3000).to(torch.int32).to(torch.uint16) for _ in range(V)]
w = [torch.full(shape, 1.0 / V, device="cuda") for _ in range(V)]
Rk = (int(os.environ.get("RL_K", "19")) - 1) // 2   # RL_K=37: the PSF of BASELINE configs[4] (spacing 0.5)
sc = Rk / 9.0
k = np.exp(-0.5 * (np.arange(-Rk, Rk + 1) / (6.0 * sc)) ** 2); kz = (k / k.sum()).astype(np.float32)
k = np.exp(-0.5 * (np.arange(-Rk, Rk + 1) / (2.0 * sc)) ** 2); kxy = (k / k.sum()).astype(np.float32)
psf = kz[:, None, None] * kxy[None, :, None] * kxy[None, None, :]
def view_psf(v):   # RL_DISTINCT=1: every view its own factors (no shared x/y passes), as with real registration parameters
    f = 1.0 + 0.004 * v
    a = np.exp(-0.5 * (np.arange(-Rk, Rk + 1) / (6.0 * sc * f)) ** 2); b = np.exp(-0.5 * (np.arange(-Rk, Rk + 1) / (2.0 * sc * f)) ** 2)
    a, b = (a / a.sum()).astype(np.float32), (b / b.sum()).astype(np.float32)
    return (a[:, None, None] * b[None, :, None] * b[None, None, :]) if v % 2 == 0 else (b[:, None, None] * b[None, :, None] * a[None, None, :])
if os.environ.get("RL_OBLIQUE"):
    from scipy import ndimage
    psf = ndimage.rotate(psf, 7.0, axes=(0, 2), reshape=False, order=1).astype(np.float32); psf /= psf.sum()
plan = rl.RLPlan(views, w, [view_psf(v) for v in range(V)] if os.environ.get('RL_DISTINCT') else [psf] * V)
plan.init_estimate()
for _ in range(3):   # eager run, graph capture, first replay
    plan.iterate()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n_it = int(os.environ.get("RL_ITERS", "2"))
a.record()
for _ in range(n_it):
    plan.iterate()
b.record(); torch.cuda.synchronize()
print("ms/iteration", a.elapsed_time(b) / n_it)
