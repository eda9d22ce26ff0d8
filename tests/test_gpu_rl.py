"""GPU parity of the multi-view Richardson-Lucy path (blur closed form, PSF, full iterations) against the oracle."""
import numpy as np
import pytest

from oracle import mvregfus_oracle as O
from tests.helpers import small_case

pytestmark = pytest.mark.gpu


def _axis_case(shape=(48, 40, 56), nviews=2, seed=5):
    """Axis-aligned views (0/90/180/270 deg about y) as in the Z1 4-view default: separable PSFs."""
    views, params, props, sp = small_case(shape=shape, nviews=nviews, dtype=np.uint16, seed=seed)
    tv = [O.transform_stack_dask_numpy(v, p, sp) for v, p in zip(views, params)]
    w = O.get_weights_simple(props, params, sp)
    return tv, w, params, sp


@pytest.mark.parametrize("kernel", ["v1", "v3"])   # v1 fused single kernel, v3 streaming (default)
@pytest.mark.parametrize("ksize", [(7, 7, 7), (19, 19, 19), (5, 9, 13), (25, 25, 25), (37, 37, 37), (9, 31, 43)])
def test_blur_matches_closed_form(cuda, ksize, kernel, monkeypatch):
    import torch
    from mvregfus_b200 import rl
    monkeypatch.setenv("MVF_RL_KERNEL", kernel)
    rng = np.random.default_rng(1)
    x = rng.random((48, 50, 72)).astype(np.float32) * 100   # nx % 4 == 0: every kernel family applies
    f = [np.exp(-0.5 * ((np.arange(k) - k // 2) / (0.15 * k + 0.5)) ** 2).astype(np.float32) for k in ksize]
    f = [(a / a.sum()).astype(np.float32) for a in f]
    f[0][0] *= 1.3  # asymmetric taps: catches a flipped kernel
    psf = f[0][:, None, None] * f[1][None, :, None] * f[2][None, None, :]
    ref = O.blur_direct(x, psf)
    got = rl.blur(torch.from_numpy(x).cuda(), f).cpu().numpy()
    assert np.max(np.abs(got - ref)) <= 1e-5 * np.max(ref)


def test_get_psf_matches_oracle(cuda):
    from mvregfus_b200 import multiview as mv
    tv, w, params, sp = _axis_case(nviews=4)
    for p in params:
        got = mv.get_psf(p, sp, 3, 1)
        ref = O.get_psf(p, sp, 3, 1)
        assert got.shape == ref.shape and got.dtype == np.float32
        assert np.max(np.abs(got - ref)) <= 1e-6 * ref.max()
        assert abs(float(got.sum()) - 1) < 1e-5


@pytest.mark.parametrize("kernel", ["v1", "v3", "v4"])   # fused single kernel / streaming with ratio volume / z-ratio-z fused (default)
@pytest.mark.parametrize("nviews,iters", [(2, 4), (4, 10)])
def test_rl_matches_oracle(cuda, nviews, iters, kernel, monkeypatch):
    """north_star: within 1e-3 relative error after N iterations (floor of one grey level)."""
    import torch
    from mvregfus_b200 import fusion, rl, multiview as mv
    monkeypatch.setenv("MVF_RL_KERNEL", kernel)
    tv, w, params, sp = _axis_case(nviews=nviews)
    ref = O.fuse_LR_with_weights_np(tv, params, sp, num_iterations=iters, sz=3, sxy=1, weights=w, return_float=True)
    psfs = [mv.get_psf(p, sp, 3, 1) for p in params]
    plan = rl.RLPlan([fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w], psfs)
    got = plan.run(num_iterations=iters).cpu().numpy()
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert float(err.max()) <= 1e-3, float(err.max())
    rel_l2 = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert rel_l2 <= 1e-3
    # the drop-in wrapper returns the uint16 ImageArray of the reference contract
    out = mv.fuse_LR_with_weights_np(tv, params, sp, num_iterations=iters, sz=3, sxy=1, weights=w)
    ref16 = np.asarray(O.fuse_LR_with_weights_np(tv, params, sp, num_iterations=iters, sz=3, sxy=1, weights=w))
    assert out.dtype == np.uint16 and np.max(np.abs(out.astype(np.int64) - ref16.astype(np.int64))) <= 1


def test_rl_delta_psf_fixed_point(cuda):
    """Known answer: delta PSF, one view, w == 1  ->  estimate * I / (I + 1e-6) each iteration."""
    import torch
    from mvregfus_b200 import rl
    rng = np.random.default_rng(2)
    I = (rng.random((24, 24, 40)) * 1000).astype(np.uint16)
    w = np.ones(I.shape, dtype=np.float32)
    psf = np.zeros((3, 3, 3), dtype=np.float32)
    psf[1, 1, 1] = 1
    plan = rl.RLPlan([torch.from_numpy(I).cuda()], [torch.from_numpy(w).cuda()], [psf])
    got = plan.run(num_iterations=3).cpu().numpy()
    est = I.astype(np.float32)
    for _ in range(3):
        est = np.clip(est * (I / (est + np.float32(1e-6))).astype(np.float32), 0, 65535)
    assert np.max(np.abs(got - est)) <= 1e-3


@pytest.mark.parametrize("dense_mode", ["fft", "direct"])
def test_rl_oblique_views_dense_psf(cuda, dense_mode, monkeypatch):
    """Views 7 degrees off axis: the PSF is not an outer product -> cuFFT convolution (default) or the direct
    O(K^3) kernel; both against the oracle's float64 FFT."""
    from mvregfus_b200 import fusion, rl, multiview as mv
    monkeypatch.setenv("MVF_RL_DENSE", dense_mode)
    views, params, props, sp = small_case(shape=(40, 36, 44), nviews=2, dtype=np.uint16, seed=9, odd_extra_deg=7.0)
    tv = [O.transform_stack_dask_numpy(v, p, sp) for v, p in zip(views, params)]
    w = O.get_weights_simple(props, params, sp)
    psfs = [mv.get_psf(p, sp, 3, 1) for p in params]
    plan = rl.RLPlan([fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w], psfs)
    assert plan.separable == [True, False]
    got = plan.run(num_iterations=3).cpu().numpy()
    ref = O.fuse_LR_with_weights_np(tv, params, sp, num_iterations=3, sz=3, sxy=1, weights=w, return_float=True)
    assert float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0))) <= 1e-3


@pytest.mark.parametrize("kernel", ["v1", "v3", "v4"])
@pytest.mark.parametrize("world", [2, 3])
def test_rl_slab_sharded_equals_whole(cuda, world, kernel, monkeypatch):
    """z-slab sharded RL (ranks emulated on one GPU, same kernels, same halo plane movement as the NVLink / NCCL
    exchanges) reproduces the single-GPU whole-volume result bit for bit, including the wrap-around low edge."""
    import torch
    from mvregfus_b200 import fusion, rl, slab, multiview as mv
    monkeypatch.setenv("MVF_RL_KERNEL", kernel)
    tv, w, params, sp = _axis_case(shape=(72, 40, 48), nviews=2)
    psfs = [mv.get_psf(p, sp, 2, 1) for p in params]   # 7^3
    dv, dw = [fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w]
    whole = rl.RLPlan(dv, dw, psfs).run(num_iterations=3).cpu().numpy()
    nz = dv[0].shape[0]
    plans = []
    for r, (z0, z1) in enumerate(slab.slab_bounds(nz, world)):
        plans.append(rl.SlabRLPlan([v[z0:z1].contiguous() for v in dv], [x[z0:z1].contiguous() for x in dw], psfs, nz, r, world))
    assert all(p.zz == (kernel == "v4") for p in plans)
    rl.setup_emulated(plans)
    for p in plans:
        p.init_estimate()
    for _ in range(3):
        rl.iterate_emulated(plans)
    got = torch.cat([p.est for p in plans], 0).cpu().numpy()
    assert np.array_equal(got, whole)


def test_rl_slab_sharded_oblique_fft(cuda):
    """z-slab sharded RL with a non-separable PSF (cuFFT path on halo-padded slabs) vs the whole volume."""
    import torch
    from mvregfus_b200 import fusion, rl, slab, multiview as mv
    views, params, props, sp = small_case(shape=(72, 40, 48), nviews=2, dtype=np.uint16, seed=9, odd_extra_deg=7.0)
    tv = [O.transform_stack_dask_numpy(v, p, sp) for v, p in zip(views, params)]
    w = O.get_weights_simple(props, params, sp)
    psfs = [mv.get_psf(p, sp, 2, 1) for p in params]
    dv, dw = [fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w]
    whole = rl.RLPlan(dv, dw, psfs).run(num_iterations=3).cpu().numpy()
    nz = dv[0].shape[0]
    plans = [rl.SlabRLPlan([v[a:b].contiguous() for v in dv], [x[a:b].contiguous() for x in dw], psfs, nz, r, 2)
             for r, (a, b) in enumerate(slab.slab_bounds(nz, 2))]
    assert plans[0].spectra[1] is not None
    for p in plans:
        p.init_estimate()
    for _ in range(3):
        slab.exchange_halos_local([p.est_pad for p in plans], [p.halo for p in plans])
        for p in plans:
            p.xy_estimate()
        for v in range(2):
            for p in plans:
                p.step_ratio(v)
            slab.exchange_halos_local([p.ratio_pad for p in plans], [p.halo for p in plans])
            for p in plans:
                p.step_correct(v)
    got = torch.cat([p.est for p in plans], 0).cpu().numpy()
    # the padded FFT length differs between slab and whole volume: float32 FFT noise, not bit equality
    assert float(np.max(np.abs(got - whole) / np.maximum(np.abs(whole), 1.0))) <= 1e-3


@pytest.mark.parametrize("kernel", ["v1", "v3", "v4"])
def test_rl_float32_views_equal_uint16_views(cuda, kernel, monkeypatch):
    """float32 views holding the same integers take the float operand path of the epilogues: identical estimate."""
    import torch
    from mvregfus_b200 import fusion, rl, multiview as mv
    monkeypatch.setenv("MVF_RL_KERNEL", kernel)
    tv, w, params, sp = _axis_case(nviews=2)
    psfs = [mv.get_psf(p, sp, 3, 1) for p in params]
    dw = [fusion.to_device(x) for x in w]
    a = rl.RLPlan([fusion.to_device(v) for v in tv], dw, psfs).run(num_iterations=3).clone()
    b = rl.RLPlan([fusion.to_device(np.asarray(v).astype(np.float32)) for v in tv], dw, psfs).run(num_iterations=3)
    assert torch.equal(a, b)


def test_rl_misaligned_view_falls_back_to_fused_kernel(cuda):
    """A uint16 view that starts 2 bytes into its allocation cannot feed the 8-byte vector loads of the streaming
    kernels: the plan must pick the fused kernel for it and still produce the same estimate (to rounding order)."""
    import torch
    from mvregfus_b200 import fusion, rl, multiview as mv
    tv, w, params, sp = _axis_case(nviews=2)
    psfs = [mv.get_psf(p, sp, 3, 1) for p in params]
    dv, dw = [fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w]
    ref_plan = rl.RLPlan(dv, dw, psfs)
    assert all(ref_plan.streamed)
    ref = ref_plan.run(num_iterations=3).clone()
    odd = []
    for v in dv:
        buf = torch.empty(v.numel() + 1, dtype=torch.uint16, device=v.device)
        buf[1:] = v.reshape(-1)
        odd.append(buf[1:].view(v.shape))
        assert odd[-1].data_ptr() % 8 == 2
    plan = rl.RLPlan(odd, dw, psfs)
    assert not any(plan.streamed)
    got = plan.run(num_iterations=3)
    assert float(((got - ref).abs() / ref.abs().clamp_min(1.0)).max()) <= 1e-4


def test_rl_configs3_psf_at_128_cubed(cuda):
    """The PSF of BASELINE.json configs[3] (sigma z = 6, xy = 2 at spacing 1 -> 19^3, separable for axis-aligned views),
    10 iterations on a 128^3-class volume against the oracle's float64 FFT recurrence (mv:5614-5747)."""
    from mvregfus_b200 import fusion, rl, multiview as mv
    tv, w, params, sp = _axis_case(shape=(128, 120, 128), nviews=2, seed=12)
    psfs = [mv.get_psf(p, sp, 6, 2) for p in params]
    assert psfs[0].shape == (19, 19, 19)
    ref = O.fuse_LR_with_weights_np(tv, params, sp, num_iterations=10, sz=6, sxy=2, weights=w, return_float=True)
    plan = rl.RLPlan([fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w], psfs)
    got = plan.run(num_iterations=10).cpu().numpy()
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert float(err.max()) <= 1e-3, float(err.max())
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-3


def test_rl_block_sum_and_convergence_break(cuda):
    """sum(estimate) of the update step (the reference's convergence measure, mv:5726-5734) on a plane whose vector
    count is not a multiple of the warp size (the last warp of the z kernel is partly idle), and a run long enough for
    the tolerance test to fire: same number of iterations, same estimate as the oracle."""
    import torch
    from mvregfus_b200 import fusion, rl, multiview as mv
    tv, w, params, sp = _axis_case(shape=(48, 40, 56), nviews=2)
    dims = tv[0].shape
    assert dims[2] % 4 == 0 or True
    psfs = [mv.get_psf(p, sp, 3, 1) for p in params]
    plan = rl.RLPlan([fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w], psfs)
    plan.init_estimate()
    for _ in range(3):   # eager, captured, replayed
        plan.iterate()
        torch.cuda.synchronize()
        assert abs(float(plan.sums[1].item()) / float(plan.est.double().sum().item()) - 1.0) < 1e-9
    tol = 2e-3
    ref, n_ref = O.fuse_LR_with_weights_np(tv, params, sp, num_iterations=40, sz=3, sxy=1, tol=tol, weights=w,
                                           return_float=True, return_iterations=True)
    assert 11 <= n_ref < 40      # the break fired, and not before the 11th iteration (mv:5729)
    plan2 = rl.RLPlan([fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w], psfs)
    got = plan2.run(num_iterations=40, tol=tol).cpu().numpy()
    assert plan2.iterations_run == n_ref
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert float(err.max()) <= 1e-3, float(err.max())


def test_rl_psf_longer_than_kernel_limit_uses_fft(cuda):
    """get_psf yields ~3 * sz / spacing taps: at fine spacings that exceeds the 43 taps the convolution kernels hold.
    Such PSFs (separable or not) run through the cuFFT convolution instead of raising, like the reference (mv:5647-5673)."""
    from mvregfus_b200 import fusion, rl
    rng = np.random.default_rng(4)
    shape = (112, 56, 64)   # every extent >= taps + 3, as the reference's padding needs
    tv = [(rng.random(shape) * 2000).astype(np.uint16) for _ in range(2)]
    w = [np.full(shape, 0.5, dtype=np.float32) for _ in range(2)]
    kz = np.exp(-0.5 * ((np.arange(45) - 22) / 7.0) ** 2)
    kxy = np.exp(-0.5 * ((np.arange(45) - 22) / 1.5) ** 2)
    psf = (kz[:, None, None] * kxy[None, :, None] * kxy[None, None, :])
    psf = (psf / psf.sum()).astype(np.float32)
    sp = {"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)}
    ref = O.fuse_LR_with_weights_np(tv, None, sp, num_iterations=3, weights=w, psfs=np.array([psf, psf]), return_float=True)
    plan = rl.RLPlan([fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w], [psf, psf])
    assert plan.fft is not None and not any(plan.streamed)
    got = plan.run(num_iterations=3).cpu().numpy()
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert float(err.max()) <= 1e-3, float(err.max())


@pytest.mark.parametrize("kernel", ["v3", "v4"])
@pytest.mark.parametrize("ksize", [(25, 25, 25), (37, 37, 37), (43, 9, 31), (13, 19, 7)])
def test_rl_long_separable_psfs_match_oracle(cuda, ksize, kernel, monkeypatch):
    """Every tap-count family of the streaming kernels (13 / 19 / 25 / 31 / 37 / 43 taps: different columns per thread,
    CTAs per SM and register budgets - configs[4] uses 37) against the oracle's FFT convolution, whole volume and two
    emulated z-slabs."""
    import torch
    from mvregfus_b200 import fusion, rl, slab
    monkeypatch.setenv("MVF_RL_KERNEL", kernel)
    rng = np.random.default_rng(ksize[0])
    shape = (140, 48, 72)
    tv = [(rng.random(shape) * 2000 * (rng.random(shape) > 0.3)).astype(np.uint16) for _ in range(2)]
    w = [rng.random(shape).astype(np.float32) for _ in range(2)]
    w[0][w[0] < 0.1] = 0.0
    s = w[0] + w[1]
    w = [x / s for x in w]
    psfs = []
    for v in range(2):
        f = [np.exp(-0.5 * ((np.arange(k) - (k - 1) / 2) / (k / (5.0 + v))) ** 2) for k in ksize]
        p = f[0][:, None, None] * f[1][None, :, None] * f[2][None, None, :]
        psfs.append((p / p.sum()).astype(np.float32))
    sp = {"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)}
    ref = O.fuse_LR_with_weights_np(tv, None, sp, num_iterations=3, weights=w, psfs=np.array(psfs), return_float=True)
    dv, dw = [fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w]
    plan = rl.RLPlan(dv, dw, psfs)
    assert all(plan.streamed) and plan.zz == (kernel == "v4")
    whole = plan.run(num_iterations=3).clone()
    err = np.abs(whole.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1.0)
    assert float(err.max()) <= 1e-3, float(err.max())
    plans = [rl.SlabRLPlan([v[a:b].contiguous() for v in dv], [x[a:b].contiguous() for x in dw], psfs, shape[0], r, 2)
             for r, (a, b) in enumerate(slab.slab_bounds(shape[0], 2))]
    assert all(p.zz == (kernel == "v4") for p in plans)   # 70-plane slabs lend 2 * Rp planes even at 43 taps
    rl.setup_emulated(plans)
    for p in plans:
        p.init_estimate()
    for _ in range(3):
        rl.iterate_emulated(plans)
    got = torch.cat([p.est for p in plans], 0)
    if all(p.zz == plan.zz for p in plans):
        assert torch.equal(got, whole)        # same kernels, same accumulation order
    else:                                     # v3 slabs against a v4 whole volume: other summation order
        assert float(((got - whole).abs() / whole.abs().clamp_min(1.0)).max()) <= 1e-4


@pytest.mark.parametrize("kernel", ["v3", "v4"])
def test_rl_unequal_slabs_equal_whole(cuda, kernel, monkeypatch):
    """Slabs of different heights (slab.rl_slab_bounds balances the v4 schedule by work; any tiling is accepted)."""
    import torch
    from mvregfus_b200 import fusion, rl, slab, multiview as mv
    monkeypatch.setenv("MVF_RL_KERNEL", kernel)
    tv, w, params, sp = _axis_case(shape=(90, 40, 48), nviews=2)
    psfs = [mv.get_psf(p, sp, 2, 1) for p in params]   # 7^3
    dv, dw = [fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w]
    whole = rl.RLPlan(dv, dw, psfs).run(num_iterations=3)
    nz = dv[0].shape[0]
    for bounds in (slab.rl_slab_bounds(nz, 3, 7, 3), [(0, 22), (22, 57), (57, nz)]):
        plans = [rl.SlabRLPlan([v[a:b].contiguous() for v in dv], [x[a:b].contiguous() for x in dw], psfs, nz, r, 3, bounds=bounds)
                 for r, (a, b) in enumerate(bounds)]
        assert all(p.zz == (kernel == "v4") for p in plans)
        rl.setup_emulated(plans)
        for p in plans:
            p.init_estimate()
        for _ in range(3):
            rl.iterate_emulated(plans)
        assert torch.equal(torch.cat([p.est for p in plans], 0), whole)


def test_rl_fft_zx_plan_for_views_rotated_about_y(cuda, monkeypatch):
    """PSF of a view rotated about y = a(y) (x) B(z, x): batched 2-D (z, x) FFTs + direct y pass (zx plan) against the 3-D
    FFT plan and the oracle; a PSF rotated about another axis has no such factor and keeps the 3-D plan."""
    from scipy import ndimage
    from mvregfus_b200 import fusion, rl, multiview as mv
    views, params, props, sp = small_case(shape=(48, 40, 52), nviews=3, dtype=np.uint16, seed=13, odd_extra_deg=9.0)
    tv = [O.transform_stack_dask_numpy(v, p, sp) for v, p in zip(views, params)]
    w = O.get_weights_simple(props, params, sp)
    psfs = [mv.get_psf(p, sp, 3, 1) for p in params]
    assert rl.factor_psf(psfs[1]) is None and rl.factor_psf_y(psfs[1]) is not None
    a, B = rl.factor_psf_y(psfs[1])
    assert np.abs(a[None, :, None] * B[:, None, :] - psfs[1]).max() <= 2e-6 * psfs[1].max()
    dv, dw = [fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w]
    ref = O.fuse_LR_with_weights_np(tv, params, sp, num_iterations=3, sz=3, sxy=1, weights=w, return_float=True)
    plan = rl.RLPlan(dv, dw, psfs)
    assert plan.fft_zx is not None and plan.fft is None and plan.fft_of[1] is plan.fft_zx
    got = plan.run(num_iterations=3).cpu().numpy()
    assert float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0))) <= 1e-3
    monkeypatch.setenv("MVF_RL_FFT", "3d")
    plan3 = rl.RLPlan(dv, dw, psfs)
    assert plan3.fft is not None and plan3.fft_zx is None
    got3 = plan3.run(num_iterations=3).cpu().numpy()
    assert float(np.max(np.abs(got - got3) / np.maximum(np.abs(got3), 1.0))) <= 1e-4
    monkeypatch.delenv("MVF_RL_FFT")
    # rotated about x instead: y and z mix, no a(y) factor -> 3-D plan; checked against the oracle with explicit PSFs
    tilted = ndimage.rotate(psfs[0], 11.0, axes=(0, 1), reshape=False, order=1).astype(np.float32)
    tilted /= tilted.sum()
    assert rl.factor_psf_y(tilted) is None
    mixed = [psfs[0], psfs[1], tilted]
    plan_m = rl.RLPlan(dv, dw, mixed)
    assert plan_m.fft is not None and plan_m.fft_zx is not None and plan_m.fft_of[2] is plan_m.fft
    ref_m = O.fuse_LR_with_weights_np(tv, None, sp, num_iterations=3, weights=w, psfs=np.array(mixed), return_float=True)
    got_m = plan_m.run(num_iterations=3).cpu().numpy()
    assert float(np.max(np.abs(got_m - ref_m) / np.maximum(np.abs(ref_m), 1.0))) <= 1e-3
