"""Steps either side of the fusion path on the device.

Drop-in counterparts, same names and argument meaning as the reference:
  * ``subsample_data(data, subsamp)``                      mvregfus/imaris.py:764-765
  * ``pyramid_levels(array, subsamp)``                     the level loop of np_to_ims / da_to_ims
                                                           (imaris.py:141-176, 316-331): cumulative
                                                           subsampling + np.histogram(data, 256) per level
  * ``subtract_background(stack, background_level)``       mvregfus/multiview.py:316
  * ``bin_stack(im, bin_factors)``                         mvregfus/mv_utils.py:288-311 (sitk.BinShrink)
  * ``illumination_fusion_planewise(stack, fusion_axis)``  mvregfus/multiview.py:391-467
numpy / ImageArray in, numpy / ImageArray out; device tensors are accepted and then returned as tensors.
The kernels live in csrc/prepost.cu behind the C-ABI; there is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi
from .fusion import _stream_ptr, dtype_code, to_device
from .image_array import ImageArray


def _as_device(a):
    if isinstance(a, torch.Tensor):
        return a.contiguous(), True
    return to_device(np.ascontiguousarray(np.asarray(a))), False


def subsample_data(data, subsamp):
    """data[0::subsamp[0], 0::subsamp[1], 0::subsamp[2]] (imaris.py:764-765)."""
    d, was_tensor = _as_device(data)
    if d.dim() != 3:
        raise ValueError("subsample_data expects a 3-D volume")
    step = [int(s) for s in subsamp]
    dims = [int(s) for s in d.shape]
    out = torch.empty([(n + s - 1) // s for n, s in zip(dims, step)], dtype=d.dtype, device=d.device)
    _cabi.check(_cabi.load().mvf_subsample(C.c_void_p(d.data_ptr()), dtype_code(d), _cabi.i32x3(dims), _cabi.i32x3(step),
                                           C.c_void_p(out.data_ptr()), _stream_ptr()))
    return out if was_tensor else out.cpu().numpy()


def value_lut(lo, hi, bins=256):
    """uint16 value -> bin of np.histogram(data, bins) for data with minimum lo and maximum hi, derived from
    numpy's own binning (so every edge decision is numpy's); values outside [lo, hi] map to bin 0 (never hit)."""
    vals = np.arange(int(lo), int(hi) + 1, dtype=np.uint16)
    counts, edges = np.histogram(vals, bins)   # same range rule as the reference call: (vals.min(), vals.max())
    lut = np.zeros(65536, dtype=np.uint8)
    lut[int(lo):int(hi) + 1] = np.repeat(np.arange(bins, dtype=np.uint8), counts)
    return lut, edges


def histogram256(data):
    """(hist, edges) == np.histogram(data, 256) for a uint16 volume (imaris.py:155, 324); counts as int64."""
    d, _ = _as_device(data)
    if d.dtype != torch.uint16:
        raise TypeError("histogram256 handles the uint16 volumes the fusion path produces")
    lib = _cabi.load()
    mm = torch.empty(2, dtype=torch.int32, device=d.device)
    _cabi.check(lib.mvf_range_u16(C.c_void_p(d.data_ptr()), d.numel(), C.c_void_p(mm.data_ptr()), _stream_ptr()))
    lo, hi = (int(x) for x in mm.cpu().numpy())
    lut, edges = value_lut(lo, hi)
    lut_d = torch.from_numpy(lut).to(d.device)
    counts = torch.empty(256, dtype=torch.int64, device=d.device)
    _cabi.check(lib.mvf_hist_u16(C.c_void_p(d.data_ptr()), d.numel(), C.c_void_p(lut_d.data_ptr()),
                                 C.c_void_p(counts.data_ptr()), _stream_ptr()))
    return counts.cpu().numpy(), edges


def pyramid_levels(array, subsamp=((1, 1, 1), (2, 2, 2), (4, 4, 4), (8, 8, 8)), histograms=True):
    """The resolution levels np_to_ims writes for one (time point, channel) volume: the subsampling factors are
    applied CUMULATIVELY (level r subsamples level r-1, imaris.py:146-148 / 321-323) and every level carries
    np.histogram(data, 256).  Returns a list of dicts {'data', 'hist', 'edges'} (host arrays)."""
    d, _ = _as_device(array)
    levels = []
    for r, ss in enumerate(subsamp):
        if len(ss) != 3:
            raise AssertionError("Only deal with 3D chunks")
        if any(int(i) > 1 for i in ss):
            d = subsample_data(d, ss)
        hist, edges = histogram256(d) if (histograms and d.dtype == torch.uint16) else (None, None)
        levels.append({"data": d.cpu().numpy(), "hist": hist, "edges": edges})
    return levels


def subtract_background(stack, background_level):
    """(stack - background_level) * (stack > background_level) on uint16 data (multiview.py:316)."""
    d, was_tensor = _as_device(stack)
    if d.dtype != torch.uint16:
        raise TypeError("subtract_background expects the raw uint16 stack")
    out = torch.empty_like(d)
    _cabi.check(_cabi.load().mvf_background_u16(C.c_void_p(d.data_ptr()), C.c_void_p(out.data_ptr()), d.numel(),
                                                int(background_level), _stream_ptr()))
    return out if was_tensor else out.cpu().numpy()


def bin_stack(im, bin_factors=np.array([1, 1, 1])):
    """mv_utils.bin_stack: sitk.BinShrink by integer factors given in x,y,z order; spacing and origin of the
    ImageArray follow mv_utils.py:297-303.  Identity factors return the input object (mv_utils.py:290-291)."""
    if np.allclose(bin_factors, [1] * len(bin_factors)):
        return im
    bin_factors = np.array(bin_factors)
    spacing, origin, rotation = im.spacing, im.origin, im.rotation
    binned_spacing = spacing * bin_factors[::-1]
    binned_origin = origin + (binned_spacing - spacing) / 2.
    d, _ = _as_device(np.asarray(im))
    f_zyx = [int(i) for i in bin_factors][::-1]
    dims = [int(s) for s in d.shape]
    out = torch.empty([n // f for n, f in zip(dims, f_zyx)], dtype=d.dtype, device=d.device)
    _cabi.check(_cabi.load().mvf_bin_shrink(C.c_void_p(d.data_ptr()), dtype_code(d), _cabi.i32x3(dims), _cabi.i32x3(f_zyx),
                                            C.c_void_p(out.data_ptr()), _stream_ptr()))
    return ImageArray(out.cpu().numpy(), spacing=binned_spacing, origin=binned_origin, rotation=rotation)


def gaussian_taps(sigma, truncate=4.0):
    """Taps and radius of scipy.ndimage.gaussian_filter1d (order 0): exp(-x^2 / (2 sigma^2)) on [-r, r], r = int(truncate *
    sigma + 0.5), normalised - the same float64 expressions as scipy's _gaussian_kernel1d."""
    sigma = float(sigma)
    radius = int(truncate * sigma + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum(), radius


def illumination_fusion_planewise(stack, fusion_axis=2):
    """Fusion of the two light-sheet illuminations of a view, plane by plane (mv:391-467): ``stack`` (2, z, y, x) uint16 ->
    (z, ., .) uint16.  Same contract as the reference, quirks included: planes (after swapping ``fusion_axis`` with the last
    axis) must be square, and the swap back is applied with the same axis number to the 3-D result (a no-op for the default
    2: the planes come back transposed)."""
    was_tensor = isinstance(stack, torch.Tensor)
    arr = stack if was_tensor else np.asarray(stack)
    if arr.ndim != 4 or arr.shape[0] != 2:
        raise ValueError("illumination_fusion_planewise expects a (2, z, y, x) stack")
    if (arr.dtype if not was_tensor else None) not in (None, np.dtype(np.uint16)) or (was_tensor and arr.dtype != torch.uint16):
        raise TypeError("illumination_fusion_planewise fuses uint16 stacks")
    d = arr if was_tensor else to_device(np.ascontiguousarray(arr))
    pair = d.view(torch.int16).swapaxes(fusion_axis, -1).contiguous().view(torch.uint16)
    nplanes, n0, n1 = (int(s) for s in pair.shape[1:])
    if n0 != n1:   # the reference indexes the rows of the (n0, n1) mask with a boolean of length n1
        raise IndexError("boolean index did not match indexed array along axis 0; size of axis is %d but size of "
                         "corresponding boolean axis is %d" % (n0, n1))
    taps, radius = gaussian_taps(n0 / 100 * 2.)
    dtaps = torch.from_numpy(taps).to(pair.device)
    px = nplanes * n0 * n0
    nbytes = 8 * px + 8 * nplanes * n0 + 16 * nplanes + 64
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=pair.device)
    out = torch.empty((nplanes, n0, n0), dtype=torch.uint16, device=pair.device)
    _cabi.check(_cabi.load().mvf_illum_fuse_planes(C.c_void_p(pair.data_ptr()), nplanes, n0, C.c_void_p(dtaps.data_ptr()), radius,
                                                   C.c_void_p(scratch.data_ptr()), nbytes, C.c_void_p(out.data_ptr()), _stream_ptr()))
    if was_tensor:
        return out.view(torch.int16).swapaxes(fusion_axis, -1).contiguous().view(torch.uint16)
    return np.swapaxes(out.cpu().numpy(), fusion_axis, -1)
