"""Edge cases of the GPU path: ragged/tiny shapes, views outside the fused box, single view, eight views,
non-contiguous inputs, error behaviour (no silent fallback)."""
import numpy as np
import pytest

from oracle import mvregfus_oracle as O
from tests.helpers import int_mismatch, rel_err, small_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(5, 7, 9), (8, 8, 32), (9, 17, 33), (3, 40, 70)])
def test_resample_ragged_and_tiny_shapes(cuda, shape):
    """Volumes smaller than one tile and extents that are not multiples of the tile / vector width."""
    from mvregfus_b200 import multiview as mv
    rng = np.random.default_rng(0)
    a = (rng.random(shape) * 1000).astype(np.float32)
    c = (np.array(shape) - 1) / 2.0
    th = 0.3
    R = np.array([[np.cos(th), 0, -np.sin(th)], [0, 1, 0], [np.sin(th), 0, np.cos(th)]])
    off = c - R @ c + np.array([0.3, -0.2, 0.4])
    out_shape = tuple(int(s) + 3 for s in shape)
    got = np.asarray(mv.affine_transform_dask(a, R, offset=off, output_shape=out_shape))
    from scipy import ndimage
    ref = ndimage.affine_transform(a, R, off, output_shape=out_shape, order=1)
    assert got.shape == ref.shape and rel_err(got, ref) <= 1e-5


def test_view_entirely_outside_gives_zeros(cuda):
    from mvregfus_b200 import multiview as mv
    a = np.full((16, 16, 16), 7, dtype=np.uint16)
    got = np.asarray(mv.affine_transform_dask(a, np.eye(3), offset=[100.0, 0, 0], output_shape=(16, 16, 16)))
    assert got.dtype == np.uint16 and not got.any()


def test_single_view_fusion_is_the_resampled_view(cuda):
    """One view: normalised weight is 1 inside its blending field -> fused == resampled view there, 0 elsewhere."""
    from mvregfus_b200 import fusion
    views, params, props, sp = small_case(nviews=1, dtype=np.uint16)
    dims = [int(s) for s in sp["size"]]
    vs = fusion.build_viewset([fusion.to_device(views[0])], props, params, sp, weights="blending")
    out = fusion.fuse_blend(vs, dims).cpu().numpy()
    tv = O.transform_stack_dask_numpy(views[0], params[0], sp)
    w = O.get_weights_simple(props, params, sp)
    ref = np.asarray(O.fuse_views_weights([tv], params, sp, weights=w))
    mx, frac = int_mismatch(out, ref)
    assert mx <= 1 and frac <= 2e-3


def test_eight_views_and_nine_is_an_error(cuda):
    from mvregfus_b200 import fusion
    views, params, props, sp = small_case(nviews=8, dtype=np.uint16, shape=(24, 20, 28))
    dims = [int(s) for s in sp["size"]]
    vs = fusion.build_viewset([fusion.to_device(v) for v in views], props, params, sp, weights="blending")
    assert fusion.fuse_blend(vs, dims).shape == tuple(dims)
    with pytest.raises(ValueError):
        fusion.build_viewset([fusion.to_device(views[0])] * 9, props * 2, list(params) + [params[0]], sp)


def test_non_contiguous_and_float64_inputs_are_accepted(cuda):
    from mvregfus_b200 import multiview as mv
    rng = np.random.default_rng(1)
    big = (rng.random((20, 22, 48)) * 500).astype(np.float32)
    sl = big[::2, 1:, ::2]                      # strided view
    got = np.asarray(mv.affine_transform_dask(sl, np.eye(3), offset=[0.5, 0.25, 0.0], output_shape=sl.shape))
    from scipy import ndimage
    ref = ndimage.affine_transform(np.ascontiguousarray(sl), np.eye(3), [0.5, 0.25, 0.0], output_shape=sl.shape, order=1)
    assert rel_err(got, ref) <= 1e-5


def test_errors_are_raised_not_swallowed(cuda):
    import ctypes as C
    import torch
    from mvregfus_b200 import _cabi, fusion
    lib = _cabi.load()
    t = torch.zeros((4, 4, 4), dtype=torch.float32, device="cuda")
    # bad dtype code
    rc = lib.mvf_affine_resample(C.c_void_p(t.data_ptr()), 7, _cabi.i32x3(t.shape), _cabi.f64(np.eye(3).ravel(), 9),
                                 _cabi.f64([0, 0, 0], 3), C.c_void_p(t.data_ptr()), _cabi.i32x3(t.shape), _cabi.i32x3([0, 0, 0]), 0, None)
    assert rc == -1 and b"dtype" in lib.mvf_last_error()
    with pytest.raises(_cabi.MvfError):
        _cabi.check(rc)
    with pytest.raises(TypeError):
        fusion.to_device(np.zeros((2, 2, 2), dtype=np.int64))
    # RL: volume smaller than the PSF is refused, not approximated
    from mvregfus_b200 import rl
    small = torch.zeros((8, 8, 8), dtype=torch.float32, device="cuda")
    k = np.ones(19, dtype=np.float32) / 19
    with pytest.raises(_cabi.MvfError):
        rl.blur(small, [k, k, k])


def test_mixed_dtypes_in_one_fusion_call_are_refused(cuda):
    from mvregfus_b200 import _cabi, fusion
    views, params, props, sp = small_case(nviews=2, dtype=np.uint16)
    dv = [fusion.to_device(views[0]), fusion.to_device(np.asarray(views[1]).astype(np.float32))]
    vs = fusion.build_viewset(dv, props, params, sp, weights="blending")
    with pytest.raises(_cabi.MvfError):
        fusion.fuse_blend(vs, [int(s) for s in sp["size"]])


def test_more_than_eight_views_weighted_fusion_and_rl(cuda):
    """The reference accepts any number of views (mv:4879-4914, 5545-5749): beyond the 8 views of one fused pass the
    drop-in entry points work in groups - weights, weighted sum (uint16 wrap / float32 clip) and the RL start estimate."""
    import numpy as np
    from mvregfus_b200 import multiview as mv, synthetic
    from oracle import mvregfus_oracle as O
    shape, V = (40, 36, 48), 10
    views, params, props = synthetic.make_views(shape, V, dtype=np.uint16, seed=7)
    sp = {"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)}
    w = mv.get_weights_simple(props, params, sp)
    wr = O.get_weights_simple(props, params, sp)
    assert len(w) == V and max(float(np.abs(np.asarray(a) - np.asarray(b)).max()) for a, b in zip(w, wr)) <= 2e-6
    from mvregfus_b200.image_array import ImageArray
    tv = [np.asarray(O.transform_stack_dask_numpy(ImageArray(v, spacing=q["spacing"], origin=q["origin"]), p, sp))
          for v, p, q in zip(views, params, props)]
    for cast in (np.uint16, np.float32):
        tvc = [t.astype(cast) * (9 if cast == np.uint16 else 1) for t in tv]      # x9: the uint16 sum wraps modulo 2^16
        big = [np.full(shape, 0.9, dtype=np.float32) for _ in range(V)]
        got = np.asarray(mv.fuse_views_weights(tvc, params, sp, weights=big))
        ref = np.asarray(O.fuse_views_weights(tvc, params, sp, weights=big))
        assert np.array_equal(got, ref)
    got = np.asarray(mv.fuse_views_weights(tv, params, sp, weights=None))
    ref = np.asarray(O.fuse_views_weights(tv, params, sp, weights=None))
    assert np.max(np.abs(got.astype(np.int64) - ref.astype(np.int64))) <= 1
    est = np.asarray(mv.fuse_LR_with_weights_np(tv, params, sp, num_iterations=2, sz=3, sxy=1, weights=wr))
    ref = np.asarray(O.fuse_LR_with_weights_np(tv, params, sp, num_iterations=2, sz=3, sxy=1, weights=wr))
    assert np.max(np.abs(est.astype(np.int64) - ref.astype(np.int64))) <= 1
