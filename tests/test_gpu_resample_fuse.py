"""GPU parity: K1 (affine resampling, both boundary rules), blending weights, weighted fusion and the fused
single-pass K2 against the CPU oracle, through the C-ABI."""
import numpy as np
import pytest

from oracle import mvregfus_oracle as O
from tests.helpers import int_mismatch, rel_err, small_case

pytestmark = pytest.mark.gpu

# tolerances (north_star): 1e-5 relative on float32 voxels; integer outputs may differ by 1 LSB where the
# float32 kernel value and the float64 reference value straddle a rounding boundary.
RTOL = 1e-5
MAX_LSB_FRACTION = 2e-3


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
@pytest.mark.parametrize("odd", [0.0, 7.0])
def test_resample_scipy_rule(cuda, dtype, odd):
    from mvregfus_b200 import multiview as mv
    views, params, props, sp = small_case(dtype=dtype, nviews=3, odd_extra_deg=odd)
    for v, p in zip(views, params):
        got = np.asarray(mv.transform_stack_dask_numpy(v, p, stack_properties=sp))
        ref = O.transform_stack_dask_numpy(v, p, sp)
        assert got.shape == ref.shape and got.dtype == ref.dtype
        if dtype == np.float32:
            assert rel_err(got, ref) <= RTOL
        else:
            mx, frac = int_mismatch(got, ref)
            assert mx <= 1 and frac <= MAX_LSB_FRACTION, (mx, frac)


def test_resample_integer_translation_exact(cuda):
    """Known-answer: a pure integer translation reproduces shifted voxels exactly, zero outside."""
    from mvregfus_b200 import multiview as mv
    rng = np.random.default_rng(0)
    a = (rng.random((20, 24, 40)) * 60000).astype(np.uint16)
    got = np.asarray(mv.affine_transform_dask(a, np.eye(3), offset=[2, -3, 5], output_shape=(20, 24, 40)))
    ref = np.zeros_like(a)
    ref[0:18, 3:24, 0:35] = a[2:20, 0:21, 5:40]
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_resample_itk_rule(cuda, dtype):
    from mvregfus_b200 import multiview as mv
    views, params, props, sp = small_case(dtype=dtype, nviews=2, odd_extra_deg=7.0)
    for v, p in zip(views, params):
        got = np.asarray(mv.transform_stack_sitk(v, p, stack_properties=sp))
        ref = np.asarray(O.transform_stack_sitk(v, p, stack_properties=sp))
        assert got.shape == ref.shape and got.dtype == ref.dtype
        if dtype == np.float32:
            assert rel_err(got, ref) <= RTOL
        else:
            mx, frac = int_mismatch(got, ref)
            assert mx <= 1 and frac <= MAX_LSB_FRACTION, (mx, frac)


def test_identity_returns_input_object(cuda):
    from mvregfus_b200 import multiview as mv
    views, params, props, sp = small_case()
    assert mv.transform_stack_sitk(views[0]) is views[0]


@pytest.mark.parametrize("odd", [0.0, 7.0])
def test_get_weights_simple(cuda, odd):
    from mvregfus_b200 import multiview as mv
    views, params, props, sp = small_case(nviews=4, odd_extra_deg=odd)
    got = mv.get_weights_simple(props, params, sp)
    ref = O.get_weights_simple(props, params, sp)
    for g, r in zip(got, ref):
        assert g.dtype == np.float32 and g.shape == r.shape
        assert float(np.max(np.abs(np.asarray(g, dtype=np.float64) - r))) <= 2e-6  # weights are in [0,1]


def test_fuse_views_weights_bit_exact(cuda):
    """Same transformed views and weights in -> identical uint16 out (truncation per view, mv:4907)."""
    from mvregfus_b200 import multiview as mv
    for dtype in (np.uint16, np.float32):
        views, params, props, sp = small_case(dtype=dtype, nviews=3)
        tv = [O.transform_stack_dask_numpy(v, p, sp) for v, p in zip(views, params)]
        w = O.get_weights_simple(props, params, sp)
        got = mv.fuse_views_weights(tv, params, sp, weights=w)
        ref = O.fuse_views_weights(tv, params, sp, weights=w)
        assert got.dtype == np.uint16 and np.array_equal(np.asarray(got), np.asarray(ref))
        got = mv.fuse_views_weights(tv, params, sp, weights=None)
        ref = O.fuse_views_weights(tv, params, sp, weights=None)
        assert np.array_equal(np.asarray(got), np.asarray(ref))


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
@pytest.mark.parametrize("nviews,odd", [(2, 0.0), (4, 0.0), (8, 7.0)])
def test_fused_single_pass(cuda, dtype, nviews, odd):
    import torch
    from mvregfus_b200 import fusion
    views, params, props, sp = small_case(dtype=dtype, nviews=nviews, odd_extra_deg=odd)
    tv = [O.transform_stack_dask_numpy(v, p, sp) for v, p in zip(views, params)]
    w = O.get_weights_simple(props, params, sp)
    ref = np.asarray(O.fuse_views_weights(tv, params, sp, weights=w))
    ref_f = O.fuse_views_weights_float(tv, w)
    dims = [int(s) for s in sp["size"]]
    dv = [fusion.to_device(v) for v in views]
    vs = fusion.build_viewset(dv, props, params, sp, weights="blending")
    outf = torch.empty(dims, dtype=torch.float32, device="cuda")
    out = fusion.fuse_blend(vs, dims, out_f32=outf).cpu().numpy()
    mx, frac = int_mismatch(out, ref)
    assert mx <= nviews and frac <= nviews * MAX_LSB_FRACTION, (mx, frac)
    if dtype == np.float32:
        assert rel_err(outf.cpu().numpy(), ref_f) <= RTOL
    # z-slab consistency: fusing two slabs separately gives the same voxels as the whole box
    h = dims[0] // 2
    lo = fusion.fuse_blend(vs, [h, dims[1], dims[2]], origin=(0, 0, 0)).cpu().numpy()
    hi = fusion.fuse_blend(vs, [dims[0] - h, dims[1], dims[2]], origin=(h, 0, 0)).cpu().numpy()
    assert np.array_equal(np.concatenate([lo, hi], 0), out)
