"""Steps either side of the path: .ims pyramid (cumulative strided subsampling +
np.histogram per level), background subtraction, BinShrink binning.  Golden vectors come from the unmodified
reference (tests/golden/make_golden_prepost.py); the CPU tests pin the oracle, the GPU tests the device path."""
import os

import numpy as np
import pytest

from oracle import mvregfus_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "prepost_golden.npz"))


# ---- oracle vs the reference's own outputs (CPU) -----------------------------------------------------------
def test_oracle_pyramid_matches_reference():
    levels = O.pyramid_levels(G["pyr_in"])
    assert len(levels) == 4
    for r, lv in enumerate(levels):
        assert np.array_equal(lv["data"], G["pyr_data%d" % r])
        assert np.array_equal(lv["hist"].astype(np.uint64), G["pyr_hist%d" % r])
        assert np.allclose([lv["edges"][0], lv["edges"][-1]], G["pyr_minmax%d" % r], rtol=0, atol=1e-3)
    assert np.array_equal(O.subsample_data(G["pyr_in"], (3, 2, 2)), G["sub_322"])


def test_oracle_bin_stack_and_background_match_reference():
    from mvregfus_b200.image_array import ImageArray
    im = ImageArray(G["bin_in"], spacing=np.array([2.0, 0.5, 0.5]), origin=np.array([1.0, -3.0, 10.0]))
    for tag, fac in (("122", [2, 2, 1]), ("341", [1, 4, 3])):
        b = O.bin_stack(im, np.array(fac))
        assert np.array_equal(np.asarray(b), G["bin_%s_data" % tag])
        assert np.array_equal(b.spacing, G["bin_%s_spacing" % tag]) and np.array_equal(b.origin, G["bin_%s_origin" % tag])
    assert O.bin_stack(im, np.array([1, 1, 1])) is im
    assert np.array_equal(O.subtract_background(G["bg_in"], 200), G["bg_out"])


def test_value_lut_reproduces_numpy_histogram():
    """The host table the device histogram uses: built from numpy's own binning, exact for any range."""
    from mvregfus_b200.prepost import value_lut
    rng = np.random.default_rng(3)
    for lo, hi in ((0, 65535), (17, 4111), (5, 5), (100, 355), (0, 255), (3, 70)):
        data = rng.integers(lo, hi + 1, size=20000).astype(np.uint16)
        data[:2] = (lo, hi)
        lut, edges = value_lut(lo, hi)
        ref, ref_edges = np.histogram(data, 256)
        assert np.array_equal(np.bincount(lut[data], minlength=256), ref)
        assert np.array_equal(edges, ref_edges)


# ---- device path (GPU) -------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_pyramid_matches_reference_golden(cuda):
    from mvregfus_b200 import prepost
    levels = prepost.pyramid_levels(G["pyr_in"])
    for r, lv in enumerate(levels):
        assert np.array_equal(lv["data"], G["pyr_data%d" % r])
        assert np.array_equal(lv["hist"].astype(np.uint64), G["pyr_hist%d" % r])
        assert np.allclose([lv["edges"][0], lv["edges"][-1]], G["pyr_minmax%d" % r], rtol=0, atol=1e-3)
    assert np.array_equal(prepost.subsample_data(G["pyr_in"], (3, 2, 2)), G["sub_322"])
    f32 = G["pyr_in"].astype(np.float32) * 0.37
    assert np.array_equal(prepost.subsample_data(f32, (2, 3, 1)), f32[::2, ::3, ::1])


@pytest.mark.gpu
def test_gpu_bin_stack_and_background_match_reference_golden(cuda):
    from mvregfus_b200 import prepost
    from mvregfus_b200.image_array import ImageArray
    im = ImageArray(G["bin_in"], spacing=np.array([2.0, 0.5, 0.5]), origin=np.array([1.0, -3.0, 10.0]))
    for tag, fac in (("122", [2, 2, 1]), ("341", [1, 4, 3])):
        b = prepost.bin_stack(im, np.array(fac))
        assert np.array_equal(np.asarray(b), G["bin_%s_data" % tag])
        assert np.array_equal(b.spacing, G["bin_%s_spacing" % tag]) and np.array_equal(b.origin, G["bin_%s_origin" % tag])
    assert prepost.bin_stack(im, np.array([1, 1, 1])) is im
    f32 = ImageArray(G["bin_in"].astype(np.float32) * 1.37, spacing=im.spacing, origin=im.origin)
    assert np.array_equal(np.asarray(prepost.bin_stack(f32, np.array([2, 3, 2]))), np.asarray(O.bin_stack(f32, np.array([2, 3, 2]))))
    assert np.array_equal(prepost.subtract_background(G["bg_in"], 200), G["bg_out"])
    for level in (0, 1, 65535):
        assert np.array_equal(prepost.subtract_background(G["bg_in"], level), O.subtract_background(G["bg_in"], level))


@pytest.mark.gpu
def test_gpu_histogram_edge_cases_and_large(cuda):
    """Constant volume (numpy widens the range to +-0.5), full uint16 range, odd sizes (scalar tails), and a
    256^3 volume against numpy; the histogram always sums to the voxel count."""
    import torch
    from mvregfus_b200 import prepost
    rng = np.random.default_rng(11)
    cases = [np.full((3, 5, 7), 1234, np.uint16),
             rng.integers(0, 65536, size=(9, 11, 13)).astype(np.uint16),
             (rng.random((256, 256, 256)) ** 3 * 4000).astype(np.uint16)]
    cases[1].flat[:2] = (0, 65535)
    for data in cases:
        hist, edges = prepost.histogram256(data)
        ref, ref_edges = np.histogram(data, 256)
        assert np.array_equal(hist, ref) and np.array_equal(edges, ref_edges) and hist.sum() == data.size
    # device-resident input stays on the device
    t = torch.from_numpy(cases[2]).cuda()
    sub = prepost.subsample_data(t, (2, 2, 2))
    assert isinstance(sub, torch.Tensor) and torch.equal(sub.cpu(), torch.from_numpy(cases[2][::2, ::2, ::2].copy()))


@pytest.mark.gpu
def test_gpu_vectorised_x2_paths_match_oracle(cuda):
    """Extents that are multiples of 16 take the LDG.128 fast paths (x step / x factor 2): same results."""
    from mvregfus_b200 import prepost
    from mvregfus_b200.image_array import ImageArray
    rng = np.random.default_rng(21)
    vol = rng.integers(0, 65536, size=(24, 36, 64)).astype(np.uint16)
    for ss in ((2, 2, 2), (1, 3, 2), (4, 1, 2)):
        assert np.array_equal(prepost.subsample_data(vol, ss), O.subsample_data(vol, ss))
    im = ImageArray(vol)
    for fac in ([2, 2, 1], [2, 1, 3], [2, 4, 2]):   # x, y, z
        assert np.array_equal(np.asarray(prepost.bin_stack(im, np.array(fac))), np.asarray(O.bin_stack(im, np.array(fac))))


@pytest.mark.gpu
def test_gpu_illumination_fusion_matches_reference(cuda):
    """illumination_fusion_planewise (mv:391-467) on the device: the reference's own outputs (run unmodified for the golden
    file), a larger case against the oracle, tensor in / tensor out, and the reference's square-plane requirement."""
    import torch
    from mvregfus_b200 import prepost
    GR = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))
    for tag in ("a", "b"):
        got = prepost.illumination_fusion_planewise(GR["illum_%s_in" % tag], fusion_axis=2)
        ref = GR["illum_%s_out" % tag]
        assert got.dtype == np.uint16 and got.shape == ref.shape
        d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
        assert d.max() <= 1 and np.mean(d > 0) <= 1e-3, (d.max(), np.mean(d > 0))
    rng = np.random.default_rng(5)
    z, y, x = np.mgrid[:5, :160, :160]
    base = 300 + 2500 * np.abs(np.sin((x + y) / 37.0)) * np.sin((y + 1) / 170.0 * np.pi) ** 2
    pair = np.array([base * (0.2 + 0.8 * y / 159.0) + rng.random(base.shape) * 40,
                     base * (1.0 - 0.8 * y / 159.0) + rng.random(base.shape) * 40]).astype(np.uint16)
    pair[:, 2, :, 70:90] = 310            # a flat band: columns with (nearly) constant mask
    ref = O.illumination_fusion_planewise(pair, fusion_axis=2)
    got = prepost.illumination_fusion_planewise(pair, fusion_axis=2)
    d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 1 and np.mean(d > 0) <= 1e-3, (d.max(), np.mean(d > 0))
    t = prepost.illumination_fusion_planewise(torch.from_numpy(pair.view(np.int16)).cuda().view(torch.uint16), fusion_axis=2)
    assert isinstance(t, torch.Tensor) and np.array_equal(t.cpu().numpy(), got)
    with pytest.raises(IndexError):
        prepost.illumination_fusion_planewise(pair[:, :, :100, :], fusion_axis=2)
