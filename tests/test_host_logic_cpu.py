"""Host-side logic that needs no GPU: PSF factorisations, the overlap-block index arithmetic of the device-resident block
loop (run here on CPU tensors), work-balanced slab bounds."""
import numpy as np
import pytest
import torch
from scipy import ndimage

from mvregfus_b200 import multiview as mv
from mvregfus_b200 import pipeline, rl


def _gauss_psf(ksize, sig):
    f = [np.exp(-0.5 * ((np.arange(k) - (k - 1) / 2) / s) ** 2) for k, s in zip(ksize, sig)]
    p = f[0][:, None, None] * f[1][None, :, None] * f[2][None, None, :]
    return (p / p.sum()).astype(np.float32)


def test_psf_factorisations():
    """factor_psf: outer product of three 1-D kernels (axis-aligned view); factor_psf_y: a(y) (x) B(z, x) (view rotated about
    y, mv:5376-5418); a rotation that mixes y with another axis has neither."""
    sep = _gauss_psf((19, 19, 19), (6.0, 2.0, 2.0))
    kz, ky, kx = rl.factor_psf(sep)
    assert np.abs(kz[:, None, None] * ky[None, :, None] * kx[None, None, :] - sep).max() <= 2e-6 * sep.max()
    about_y = ndimage.rotate(sep, 7.0, axes=(0, 2), reshape=False, order=1).astype(np.float32)
    about_y /= about_y.sum()
    assert rl.factor_psf(about_y) is None
    a, B = rl.factor_psf_y(about_y)
    assert a.shape == (19,) and B.shape == (19, 19)
    assert np.abs(a[None, :, None] * B[:, None, :] - about_y).max() <= 2e-6 * about_y.max()
    about_x = ndimage.rotate(sep, 7.0, axes=(0, 1), reshape=False, order=1).astype(np.float32)
    assert rl.factor_psf(about_x) is None and rl.factor_psf_y(about_x) is None
    long_y = _gauss_psf((9, 45, 9), (2.0, 9.0, 2.0))      # more y taps than the direct pass holds
    assert rl.factor_psf_y(long_y) is None and rl.too_long_for_kernels(long_y)


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
@pytest.mark.parametrize("depth", [0, 5, 20])
def test_overlap_block_matches_reference_overlap(dtype, depth):
    """Device-resident overlap blocks (halo from the neighbours, reflection at the volume faces, zeros beyond the volume;
    overlap_from_sources mv:3662-3707) against the numpy block builder of the host loop, on every chunk of a ragged volume."""
    rng = np.random.default_rng(3)
    shape, chunk = (70, 33, 41), 32
    src = [(rng.random(shape) * 1000).astype(dtype) for _ in range(2)]
    tsrc = [torch.from_numpy(s.view(np.int16) if dtype == np.uint16 else s) for s in src]
    if dtype == np.uint16:
        tsrc = [t.view(torch.uint16) for t in tsrc]
    nch = [-(-s // chunk) for s in shape]
    for loc in list(np.ndindex(*nch)) + [(nch[0], 0, 0)]:
        want = mv._overlap_block(src, loc, depth, chunk)
        got = pipeline.overlap_block_device(tsrc, loc, depth, chunk)
        got = got.numpy().view(np.uint16) if dtype == np.uint16 else got.numpy()
        assert got.shape == want.shape and np.array_equal(got, want), loc


def test_gaussian_taps_are_scipys():
    """prepost.gaussian_taps restates scipy's _gaussian_kernel1d: filtering a delta must return exactly these taps."""
    from mvregfus_b200 import prepost
    for sigma in (0.88, 3.2, 20.48):
        taps, r = prepost.gaussian_taps(sigma)
        delta = np.zeros(4 * r + 1)
        delta[2 * r] = 1.0
        resp = ndimage.gaussian_filter1d(delta, sigma, mode="constant")
        assert r == int(4.0 * sigma + 0.5) and np.array_equal(resp[r:3 * r + 1], taps)


def test_chunk_decisions_follow_fuse_block():
    """Per-chunk decision words of the one-launch block loop (mv:3826-3909): bits 0-7 views taking part, bit 8 copy mode,
    bits 12-15 the view whose transformed voxels are copied."""
    data = np.array([0b000, 0b010, 0b111, 0b111, 0b110, 0b011], dtype=np.uint32)
    wbits = np.array([0b000, 0b010, 0b111, 0b100, 0b000, 0b011], dtype=np.uint32)
    probe = np.ones((6, 3), dtype=bool)
    probe[5, 0] = False            # probe-point early-out drops view 0 of the last chunk
    info = pipeline.chunk_decisions(data, wbits, probe)
    copy = lambda v: (1 << 8) | (v << 12)
    assert list(info) == [copy(0),   # no live view: tviews_block[0]
                          copy(1),   # one live view: raw copy of it
                          0b111,     # three weighted views: fuse them
                          copy(2),   # only view 2 has weight: copy it
                          copy(1),   # no weighted view: first live view
                          copy(1)]   # views 0 and 1 live, view 0 dropped by the probe: copy view 1


def test_dask_wrap_is_taken_when_dask_is_importable(monkeypatch):
    """The resampling entry points return dask arrays when dask is installed (the reference's contract, mv:1458-1519); the
    image has no dask, so the wrap is exercised with a stand-in module."""
    import sys
    import types
    calls = []
    da = types.ModuleType("dask.array")
    da.from_array = lambda x, chunks=None: (calls.append((x.shape, chunks)), ("wrapped", x))[1]
    dask = types.ModuleType("dask")
    dask.array = da
    monkeypatch.setitem(sys.modules, "dask", dask)
    monkeypatch.setitem(sys.modules, "dask.array", da)
    a = np.zeros((4, 6, 8), dtype=np.uint16)
    out = mv._maybe_dask(a, (2, 3, 4))
    assert out[0] == "wrapped" and out[1] is a and calls == [((4, 6, 8), (2, 3, 4))]
    monkeypatch.delitem(sys.modules, "dask.array")
    monkeypatch.setitem(sys.modules, "dask", None)      # import dask.array -> ImportError: plain ndarray back
    assert mv._maybe_dask(a, (2, 3, 4)) is a
