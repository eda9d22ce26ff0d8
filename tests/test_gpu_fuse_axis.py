"""GPU parity of the table-driven, TMA-staged fused kernel (axis-aligned views, csrc/fuse_axis.cuh) against the CPU
oracle and against the general per-voxel kernel, through the C-ABI.  Reference path: mvregfus/multiview.py:1458-1634
(resampling), 3469-3633 (blending weights), 4904-4912 (fusion)."""
import os

import numpy as np
import pytest

from oracle import mvregfus_oracle as O
from tests.helpers import int_mismatch, rel_err, small_case

pytestmark = pytest.mark.gpu


def _fuse(views, props, params, sp, want_f32=False, general=False, origin=(0, 0, 0), dims=None):
    import torch
    from mvregfus_b200 import fusion
    dims = [int(s) for s in sp["size"]] if dims is None else dims
    dv = [fusion.to_device(v) for v in views]
    vs = fusion.build_viewset(dv, props, params, sp, weights="blending")
    outf = torch.empty(dims, dtype=torch.float32, device="cuda") if want_f32 else None
    old = os.environ.get("MVF_FUSE_KERNEL")
    try:
        if general:
            os.environ["MVF_FUSE_KERNEL"] = "general"
        else:
            os.environ.pop("MVF_FUSE_KERNEL", None)
        out = fusion.fuse_blend(vs, dims, origin=origin, out_f32=outf).cpu().numpy()
    finally:
        if old is None:
            os.environ.pop("MVF_FUSE_KERNEL", None)
        else:
            os.environ["MVF_FUSE_KERNEL"] = old
    return out, (outf.cpu().numpy() if want_f32 else None)


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
@pytest.mark.parametrize("shape,nviews", [((40, 36, 48), 4), ((72, 40, 64), 4), ((33, 50, 56), 8), ((64, 64, 64), 2)])
def test_axis_kernel_vs_oracle(cuda, dtype, shape, nviews):
    """0/90/180/270 degree views (8 views: 45 degree steps, so the call falls back to the general kernel and the test
    pins that the dispatch stays correct) with 1 % anisotropic scale and sub-voxel shifts."""
    views, params, props, sp = small_case(shape=shape, dtype=dtype, nviews=nviews)
    tv = [O.transform_stack_dask_numpy(v, p, sp) for v, p in zip(views, params)]
    w = O.get_weights_simple(props, params, sp)
    ref = np.asarray(O.fuse_views_weights(tv, params, sp, weights=w))
    ref_f = O.fuse_views_weights_float(tv, w)
    out, outf = _fuse(views, props, params, sp, want_f32=True)
    mx, frac = int_mismatch(out, ref)
    assert mx <= nviews and frac <= nviews * 2e-3, (mx, frac)
    if dtype == np.float32:
        assert rel_err(outf, ref_f) <= 1e-5
    out2, _ = _fuse(views, props, params, sp)
    assert np.array_equal(out, out2)   # the float32 side output does not change the uint16 result


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_axis_kernel_vs_general_kernel(cuda, dtype):
    """Same inputs through both kernels: identical up to float32 rounding of the interpolation (<= 1 LSB per view)."""
    views, params, props, sp = small_case(shape=(80, 48, 96), dtype=dtype, nviews=4)
    a, af = _fuse(views, props, params, sp, want_f32=True)
    g, gf = _fuse(views, props, params, sp, want_f32=True, general=True)
    mx, frac = int_mismatch(a, g)
    assert mx <= 4 and frac <= 4e-3, (mx, frac)
    if dtype == np.float32:   # (uint16 views are rounded per view before weighting: a 1-ulp difference can flip an integer)
        assert rel_err(af, gf) <= 2e-6


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_axis_kernel_slabs_and_chunks_bit_exact(cuda, dtype):
    """A voxel's value is a closed form of its fused-frame index: ragged z-slabs and an off-grid sub-box reproduce the
    whole volume bit for bit."""
    views, params, props, sp = small_case(shape=(72, 40, 64), dtype=dtype, nviews=4)
    dims = [int(s) for s in sp["size"]]
    whole, _ = _fuse(views, props, params, sp)
    cuts = [0, 19, 37, dims[0]]
    slabs = [_fuse(views, props, params, sp, origin=(a, 0, 0), dims=[b - a, dims[1], dims[2]])[0] for a, b in zip(cuts, cuts[1:])]
    assert np.array_equal(np.concatenate(slabs, 0), whole)
    sub, _ = _fuse(views, props, params, sp, origin=(5, 9, 11), dims=[30, 21, 40])
    assert np.array_equal(sub, whole[5:35, 9:30, 11:51])


def test_axis_kernel_upsampling_and_exact_permutation(cuda):
    """2x up-sampling into a spacing-0.5 frame (configs[4] geometry) and an exact 90 degree rotation without jitter."""
    from mvregfus_b200 import synthetic
    from mvregfus_b200.image_array import ImageArray
    shape = (32, 24, 40)
    views, params, props = synthetic.make_views(shape, 4, dtype=np.float32, seed=5, jitter=False)
    views = [ImageArray(v, spacing=p["spacing"], origin=p["origin"]) for v, p in zip(views, props)]
    for spacing in (0.5, 1.0):
        sp = O.calc_stack_properties_from_views_and_params(props, params, spacing=[spacing] * 3)
        tv = [O.transform_stack_dask_numpy(v, p, sp) for v, p in zip(views, params)]
        w = O.get_weights_simple(props, params, sp)
        ref_f = O.fuse_views_weights_float(tv, w)
        out, outf = _fuse(views, props, params, sp, want_f32=True)
        assert rel_err(outf, ref_f) <= 1e-5, spacing
