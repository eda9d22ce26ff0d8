"""GPU parity of the DCT-entropy quality weights (K3 + coarse-grid heuristics + DCT mode of the weights kernel)."""
import os

import numpy as np
import pytest

from oracle import mvregfus_oracle as O

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))


def _case():
    from tests.golden.make_golden import blob_volume
    params = G["dct_params"]
    props = [{"size": np.array((96, 96, 96)), "spacing": np.ones(3), "origin": np.zeros(3)} for _ in range(2)]
    sp = {"size": np.array([128, 128, 128]), "spacing": np.ones(3), "origin": G["dct_origin"]}
    tviews = np.array([blob_volume((128, 128, 128), 70 + i, np.uint16, 0.5) for i in range(2)])
    return params, props, sp, tviews


def test_chunk_quality_matches_reference(cuda):
    """determine_chunk_quality against the value the unmodified reference produced (golden) and the oracle."""
    from mvregfus_b200 import multiview as mv
    params, props, sp, tviews = _case()
    blk = tviews[:, 32:64, 32:64, 32:64]
    qprops = {"size": np.array([32] * 3), "spacing": np.array([2.0] * 3), "origin": sp["origin"] + 32.0}
    got = mv.determine_chunk_quality(blk, props, params, qprops,
                                     block_info={0: {"array-location": [(0, 2), (0, 32), (0, 32), (0, 32)]}})
    assert got.shape == (2, 1, 1, 1)
    assert np.allclose(got.ravel(), G["dct_block_quality"], rtol=1e-10, atol=0)


def test_quality_grid_matches_oracle(cuda):
    from mvregfus_b200 import dct, fusion
    params, props, sp, tviews = _case()
    ref, b, size = O.dct_quality_grid(tviews, params, props, sp)
    got, b2, size2 = dct.quality_grid([fusion.to_device(t) for t in tviews], params, props, sp)
    assert (b, size) == (b2, size2)
    assert np.array_equal(np.isnan(ref), np.isnan(got))
    m = ~np.isnan(ref)
    assert np.allclose(got[m], ref[m], rtol=1e-9, atol=1e-12)


def test_dct_weights_match_reference(cuda):
    from mvregfus_b200 import multiview as mv
    params, props, sp, tviews = _case()
    wd = np.asarray(mv.get_weights_dct_dask(tviews, params, props, sp, depth=0))
    assert wd.shape == (2, 128, 128, 128) and wd.dtype == np.float32
    assert np.max(np.abs(wd[:, ::8, ::8, ::8].astype(np.float64) - G["dct_weights_sub"])) <= 5e-6
    ref = O.get_weights_dct(tviews, params, props, sp, depth=0)[(0, 0, 0)]
    assert np.max(np.abs(wd.astype(np.float64) - ref)) <= 5e-6


def test_fused_dct_mode(cuda):
    """K2 in DCT mode == oracle: transform -> DCT weights -> fuse_views_weights (one 128^3 chunk)."""
    import torch
    from mvregfus_b200 import dct, fusion
    from mvregfus_b200.image_array import ImageArray
    from tests.golden.make_golden import blob_volume
    params, props, sp, _ = _case()
    views = [ImageArray(blob_volume((96, 96, 96), 170 + i, np.uint16, 0.3)) for i in range(2)]
    tv = [O.transform_stack_dask_numpy(v, p, sp) for v, p in zip(views, params)]
    w = O.get_weights_dct(np.array(tv), params, props, sp, depth=0)[(0, 0, 0)]
    ref = np.asarray(O.fuse_views_weights(tv, params, sp, weights=list(w)))
    dv = [fusion.to_device(v) for v in views]
    dtv = [fusion.to_device(t) for t in tv]
    ws, b, size = dct.quality_grid(dtv, params, props, sp)
    coarse = dct.coarse_weights(ws, sp, b, size)
    vs = fusion.build_viewset(dv, props, params, sp, weights="dct", coarse=coarse)
    out = fusion.fuse_blend(vs, [128, 128, 128]).cpu().numpy()
    d = np.abs(out.astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 2 and np.mean(d > 0) <= 4e-3, (d.max(), np.mean(d > 0))


def test_dct_weights_multi_chunk_8_views(cuda):
    """configs[2] in miniature: 8 views (every other one 7 degrees off axis), a 2x2x2 grid of 128^3 chunks with a
    4-voxel halo - the nan-max filter, the per-cell exponent fit and the nan-Gaussian act ACROSS chunks here
    (mv:4122-4242), and every chunk multiplies its own blending weights with the up-sampled coarse grid (mv:4256-4317)."""
    from mvregfus_b200 import multiview as mv
    from mvregfus_b200 import synthetic
    from tests.golden.make_golden import blob_volume
    V, n, depth = 8, 256, 4
    params = synthetic.view_params((200, 200, 200), V, seed=21, odd_extra_deg=7.0)
    props = [{"size": np.array((200, 200, 200)), "spacing": np.ones(3), "origin": np.zeros(3)} for _ in range(V)]
    sp = {"size": np.array((n, n, n)), "spacing": np.ones(3), "origin": np.array([-28.0, -28.0, -28.0])}
    tviews = np.array([blob_volume((n, n, n), 900 + i, np.uint16, 0.5) for i in range(V)])
    ref = O.get_weights_dct(tviews, params, props, sp, depth=depth)
    got = np.asarray(mv.get_weights_dct_dask(tviews, params, props, sp, depth=depth))
    edge = 128 + 2 * depth
    assert got.shape == (V, 2 * edge, 2 * edge, 2 * edge) and got.dtype == np.float32
    assert len(ref) == 8
    worst = 0.0
    for loc, w in ref.items():
        blk = got[:, loc[0] * edge:(loc[0] + 1) * edge, loc[1] * edge:(loc[1] + 1) * edge, loc[2] * edge:(loc[2] + 1) * edge]
        worst = max(worst, float(np.max(np.abs(blk.astype(np.float64) - w))))
    assert worst <= 5e-6, worst
