"""Drop-in replacements for the view-fusion entry points of ``mvregfus.multiview``.

Same names, positional/keyword signatures, numpy/``ImageArray`` contracts and error behaviour as the
reference functions (``mv`` = /root/reference/mvregfus/multiview.py); the voxel arithmetic runs in the
CUDA kernels of libmvregfus_b200 through the C-ABI.  All functions are plain Python ``def``s because
``fuse_block`` introspects ``weights_func.__code__.co_varnames`` (mv:3864-3874).

Host arrays in, host arrays out (the reference's contract); the device-resident API that avoids the
PCIe round trip per call is ``mvregfus_b200.fusion`` / ``mvregfus_b200.pipeline``.
"""
import os

import numpy as np
import torch

from . import fusion
from .image_array import ImageArray
from .mv_utils import (index_space_affine, invert_params, matrix_to_params, params_invert_coordinates,  # noqa: F401
                       params_to_matrix, transform_points)

__all__ = [
    "transform_stack_sitk", "transform_stack_dask_numpy", "affine_transform_dask",
    "get_weights_simple", "blocks_inside", "fuse_views_weights", "get_sample_volume", "get_union_volume",
    "calc_stack_properties_from_volume", "calc_stack_properties_from_views_and_params",
    "get_psf", "fuse_LR_with_weights_np",
    "get_dct_options", "determine_chunk_quality", "get_weights_dct_dask", "get_weights_dct",
    "fuse_block", "fuse_blockwise", "fuse_blockwise_arrays", "transform_view_dask_and_save_chunked",
]


# ---------------------------------------------------------------------------------------------------
# a3: fused-frame bounding box (host, float64 -> integer; mv:3030-3122, 4916-4932)
# ---------------------------------------------------------------------------------------------------
def get_sample_volume(stack_properties_list, params):
    """Lower/upper physical corner of the union of all views back-projected into the fused frame."""
    ndim = len(stack_properties_list[0]["spacing"])
    unit = np.array(list(np.ndindex(*([2] * ndim))))
    verts = []
    for sp, p in zip(stack_properties_list, params):
        corners = unit * np.array(sp["size"]) * np.array(sp["spacing"]) + np.array(sp["origin"])
        Minv = params_to_matrix(invert_params(p))
        verts.append(np.dot(Minv[:ndim, :ndim], corners.T).T + Minv[:ndim, ndim])
    verts = np.concatenate(verts, 0)
    return np.min(verts, 0), np.max(verts, 0)


get_union_volume = get_sample_volume  # identical computation in the reference (mv:3030-3052)


def calc_stack_properties_from_volume(volume, spacing):
    """{'size','spacing','origin'} of the fused frame; size is uint16 like the reference (mv:3115)."""
    spacing = np.array(spacing).astype(np.float64)
    size = np.ceil((volume[1] - volume[0]) / spacing).astype(np.uint16)
    return {"size": size, "spacing": spacing, "origin": volume[0]}


def calc_stack_properties_from_views_and_params(views_props, params, spacing=None, mode="sample"):
    """mv:4916-4932.  (The reference wraps this in its io_decorator; that file-cache layer is the caller's.)"""
    spacing = np.array(spacing).astype(np.float64)
    if mode in ("sample", "union"):
        volume = get_sample_volume(views_props, params)
    elif mode == "intersection":
        raise Exception("intersection mode is broken in the reference (mv:3068) and not supported")
    else:
        raise Exception("can't understand volume mode %s" % mode)
    return calc_stack_properties_from_volume(volume, spacing)


def blocks_inside(orig_stack_propertiess, params, stack_properties, n_points_per_dim=5):
    """mv:3406-3466 (host, bit-exact integer codes)."""
    return fusion.blocks_inside_codes(orig_stack_propertiess, params, stack_properties, n_points_per_dim)


# ---------------------------------------------------------------------------------------------------
# a4-a7: resampling
# ---------------------------------------------------------------------------------------------------
def _out_geometry(stack, out_shape, out_spacing, out_origin, stack_properties):
    if stack_properties is not None:
        return stack_properties["size"], stack_properties["origin"], stack_properties["spacing"]
    return (stack.shape if out_shape is None else out_shape,
            stack.origin if out_origin is None else out_origin,
            stack.spacing if out_spacing is None else out_spacing)


def _resample_host(stack, p, out_shape, out_origin, out_spacing, rule):
    matrix, offset = index_space_affine(p, stack.spacing, stack.origin, out_spacing, out_origin)
    dims = [int(np.ceil(s)) for s in out_shape]
    src = fusion.to_device(stack)
    res = fusion.affine_resample(src, matrix, offset, dims, (0, 0, 0), rule=rule)
    out = fusion.to_host(res)
    if out.dtype != np.asarray(stack).dtype and np.issubdtype(np.asarray(stack).dtype, np.floating):
        out = out.astype(np.asarray(stack).dtype)
    return out


class LazyTransformedView:
    """Result of :func:`transform_stack_dask_numpy`.  The reference returns a *lazy* dask array there (mv:1458-1519):
    nothing is resampled until a consumer pulls chunks.  This object keeps that contract without dask - it knows its
    ``shape`` / ``dtype`` / ``spacing`` / ``origin``, materialises on ``np.asarray()`` / indexing / ``compute()`` (one
    K1 launch for the whole volume, cached), and lets :func:`fuse_blockwise_arrays` fuse straight from the raw views
    when every input is still lazy, so the transformed views never exist in host memory."""

    def __init__(self, stack, p, out_shape, out_origin, out_spacing, chunksize):
        self.stack, self.p = stack, np.array(p, dtype=np.float64)
        self.shape = tuple(int(np.ceil(x)) for x in out_shape)
        self.spacing = np.array(out_spacing, dtype=np.float64)
        self.origin = np.array(out_origin, dtype=np.float64)
        self.dtype = np.asarray(stack).dtype
        self.ndim = len(self.shape)
        self.chunks = tuple([int(chunksize)] * self.ndim)
        self._value = None

    def compute(self):
        if self._value is None:
            self._value = _resample_host(self.stack, self.p, self.shape, self.origin, self.spacing, "scipy")
        return self._value

    def __array__(self, dtype=None, copy=None):
        v = self.compute()
        return v if dtype is None else v.astype(dtype)

    def __getitem__(self, idx):
        return self.compute()[idx]

    def __len__(self):
        return self.shape[0]

    def max(self, *a, **k):
        return self.compute().max(*a, **k)


def transform_stack_sitk(stack, p=None, out_shape=None, out_spacing=None, out_origin=None, interp="linear",
                         stack_properties=None):
    """ITK-rule linear resampling (mv:1771-1824 + transformStack mv:6797-6877).

    The identity short-circuit returns the input object itself (mv:1813)."""
    ndim = stack.ndim
    if ndim != 3:
        raise Exception("mvregfus_b200 resamples 3-D stacks only")
    if p is None:
        p = matrix_to_params(np.eye(ndim + 1))
    p = np.array(p)
    out_shape, out_origin, out_spacing = _out_geometry(stack, out_shape, out_spacing, out_origin, stack_properties)
    if (np.allclose(p, matrix_to_params(np.eye(ndim + 1))) and np.allclose(out_shape, stack.shape)
            and np.allclose(out_origin, stack.origin) and np.allclose(out_spacing, stack.spacing)):
        return stack
    if interp != "linear":
        raise Exception("mvregfus_b200.transform_stack_sitk implements interp='linear' (got %r); bspline/"
                        "gaussian/nearest are used only by registration code outside the fusion path" % (interp,))
    res = _resample_host(stack, p, out_shape, out_origin, out_spacing, "itk")
    return ImageArray(res, origin=out_origin, spacing=out_spacing)


def transform_stack_dask_numpy(stack, p=None, out_shape=None, out_spacing=None, out_origin=None, interp="linear",
                               stack_properties=None, chunksize=512):
    """scipy-rule order-1 resampling of a view into the fused frame (mv:1458-1519).

    The reference returns a lazy dask array that evaluates ``scipy.ndimage.affine_transform(order=1)`` per chunk;
    here a :class:`LazyTransformedView` (wrapped with ``da.from_array`` when dask is installed) that resamples the
    whole volume with one kernel launch the first time its voxels are asked for."""
    out_shape, out_origin, out_spacing = _out_geometry(stack, out_shape, out_spacing, out_origin, stack_properties)
    if not hasattr(stack, "spacing"):
        stack = ImageArray(stack)
    lazy = LazyTransformedView(stack, p, out_shape, out_origin, out_spacing, chunksize)
    try:
        import dask.array as da
    except ImportError:
        return lazy
    return da.from_array(lazy, chunks=lazy.chunks)   # dask pulls through __getitem__: still lazy


def affine_transform_dask(input, matrix, offset=0.0, output_shape=None, output_chunks=(128, 128, 128), **kwargs):
    """Index-space affine resampling, always order 1 like the reference (mv:1524-1634, order fixed at mv:1610)."""
    if output_shape is None:
        output_shape = input.shape
    matrix = np.asarray(matrix, dtype=np.float64)
    if matrix.ndim == 1:
        matrix = np.diag(matrix)
    offset = np.broadcast_to(np.asarray(offset, dtype=np.float64), (3,))
    src = fusion.to_device(input)
    res = fusion.affine_resample(src, matrix, offset, [int(s) for s in output_shape], (0, 0, 0), rule="scipy")
    return _maybe_dask(fusion.to_host(res), output_chunks)


def transform_view_dask_and_save_chunked(fn, view, params, iview, stack_properties, chunksize=128):
    """Resample view ``iview`` into the fused frame and store it as HDF5 dataset 'Data' in ``fn`` (mv:1746-1769).
    File formats stay the reference's business: this needs h5py."""
    try:
        import h5py
    except ImportError as e:
        raise Exception("transform_view_dask_and_save_chunked writes HDF5 and needs h5py (%s)" % e)
    res = transform_stack_dask_numpy(view, params[iview], stack_properties=stack_properties, interp='linear',
                                     chunksize=chunksize)
    with h5py.File(fn, 'w') as f:
        f.create_dataset('Data', data=np.asarray(res))
    return fn


def _maybe_dask(arr, chunks):
    try:
        import dask.array as da
    except ImportError:
        return arr
    return da.from_array(arr, chunks=tuple(int(c) for c in chunks))


# ---------------------------------------------------------------------------------------------------
# a9: blending weights, a13: weighted additive fusion
# ---------------------------------------------------------------------------------------------------
def get_weights_simple(orig_stack_propertiess, params, stack_properties):
    """Sigmoid-border blending weights of every view on the target grid, normalised over views
    (mv:3469-3633).  Returns a list of V float32 arrays like the reference."""
    dims = [int(s) for s in stack_properties["size"]]
    if len(params) <= 8:
        vs = fusion.build_viewset(None, orig_stack_propertiess, params, stack_properties, weights="blending")
        w = fusion.to_host(fusion.blend_weights(vs, dims, (0, 0, 0), normalise=True))
    else:   # more than 8 views: raw weights per group of 8, then the reference's normalisation (sum in view order, 0 -> 1)
        raw = []
        for i0 in range(0, len(params), 8):
            vs = fusion.build_viewset(None, orig_stack_propertiess[i0:i0 + 8], params[i0:i0 + 8], stack_properties, weights="blending")
            raw.extend(fusion.blend_weights(vs, dims, (0, 0, 0), normalise=False).unbind(0))
        wsum = torch.zeros_like(raw[0])
        for r in raw:
            wsum += r
        wsum[wsum == 0] = 1
        w = fusion.to_host(torch.stack([r / wsum for r in raw]))
    return [ImageArray(w[i], origin=stack_properties["origin"], spacing=stack_properties["spacing"])
            for i in range(len(params))]


def fuse_views_weights(views, params, stack_properties, weights=None, views_in_target_space=True):
    """f = sum_v uint16(w_v * I_v), clipped, uint16 (mv:4879-4914)."""
    if not views_in_target_space:
        views = [np.array(transform_stack_sitk(view, params[iview], out_origin=stack_properties["origin"],
                                               out_shape=stack_properties["size"],
                                               out_spacing=stack_properties["spacing"]))
                 for iview, view in enumerate(views)]
    dviews = [fusion.to_device(v) for v in views]
    dweights = [fusion.to_device(np.asarray(w, dtype=np.float32)) for w in weights] if weights is not None else None
    if len(dviews) <= 8:
        out = fusion.fuse_weighted(dviews, dweights)
    elif dweights is None:
        # mean of more than 8 views: float64 sum on the device, clipped and truncated like np.mean(...) -> astype(uint16)
        acc = torch.zeros(dviews[0].shape, dtype=torch.float64, device=dviews[0].device)
        for v in dviews:
            acc += v.to(torch.int32).to(torch.float64) if v.dtype == torch.uint16 else v.to(torch.float64)
        out = (acc / len(dviews)).clamp_(0, 65535).to(torch.int32).to(torch.uint16)
    else:
        # more than 8 views: one fused pass per group of 8.  Every term uint16(w_v * I_v) is a non-negative integer, so
        # the groups add exactly: uint16 views accumulate modulo 2^16 (the reference's uint16 `f`), float32 views saturate
        # at 65535 (a group that saturates means the total does, too).
        acc = torch.zeros(dviews[0].shape, dtype=torch.int32, device=dviews[0].device)
        for i0 in range(0, len(dviews), 8):
            acc += fusion.fuse_weighted(dviews[i0:i0 + 8], dweights[i0:i0 + 8]).to(torch.int32)
        out = (acc & 0xFFFF if dviews[0].dtype == torch.uint16 else acc.clamp_(max=65535)).to(torch.uint16)
    return ImageArray(fusion.to_host(out), spacing=stack_properties["spacing"], origin=stack_properties["origin"])


# ---------------------------------------------------------------------------------------------------
# a14/a15: PSF and multi-view Richardson-Lucy
# ---------------------------------------------------------------------------------------------------
_PSF_CACHE = {}


def get_psf(p, stack_properties, sz, sxy):
    """Gaussian PSF of a view in the fused frame (mv:5376-5418), float32, normalised to sum 1.

    Extent per axis: ``int(max(1, 3 * max(sz, sxy)) / spacing)`` made odd; a centred delta is blurred with
    ``sigma = (sz, sxy, sxy) / spacing`` (scipy's reflect mode, as the reference) and rotated about the centre by the
    linear part of the view's affine with ITK-rule linear resampling (K1).  Cached per (linear part, spacing, sigmas):
    every 128^3 block of a fusion asks for the same PSFs."""
    from scipy import ndimage
    spacing = np.asarray(stack_properties['spacing'], dtype=np.float64)
    linear = np.asarray(p, dtype=np.float64)[:9]
    key = (tuple(linear), tuple(spacing), float(sz), float(sxy))
    hit = _PSF_CACHE.get(key)
    if hit is not None:
        return hit.copy()
    sigmas = np.array([sz, sxy, sxy], dtype=np.float64)
    extent = (np.array([max(1, 3 * sigmas.max())]) / spacing).astype(np.int64)
    extent = extent + (1 - extent % 2)          # even extents grow by one voxel
    centre = extent // 2
    delta = np.zeros(extent, dtype=np.float32)
    delta[tuple(centre)] = 1
    corner = -spacing * centre                  # physical position of voxel 0: the delta sits at the origin
    gauss = ImageArray(ndimage.gaussian_filter(delta, sigmas / spacing), spacing=spacing, origin=corner)
    rotation_only = np.concatenate([linear, np.zeros(3)])
    rotated = np.asarray(transform_stack_sitk(gauss, rotation_only, out_origin=corner, out_shape=extent,
                                              out_spacing=spacing, interp='linear'))
    psf = (rotated / np.sum(rotated)).astype(np.float32)
    if len(_PSF_CACHE) > 64:
        _PSF_CACHE.clear()
    _PSF_CACHE[key] = psf
    return psf.copy()


def fuse_LR_with_weights_np(views, params, stack_properties, num_iterations=25, sz=4, sxy=0.5, tol=5e-5,
                            weights=None):
    """Multi-view Richardson-Lucy with weights (mv:5545-5749): returns the uint16 ImageArray estimate."""
    from . import rl
    psfs = [get_psf(params[ip], stack_properties, sz, sxy) for ip in range(len(params))]
    dviews = [fusion.to_device(v) for v in views]
    dweights = [fusion.to_device(np.asarray(w, dtype=np.float32)) for w in weights]
    plan = rl.RLPlan(dviews, dweights, psfs)
    est = plan.run(num_iterations=num_iterations, tol=tol)
    est = fusion.to_host(est.to(torch.int32).to(torch.uint16))   # float32 -> uint16 truncates like ndarray.astype (values are clipped to [0, 65535])
    return ImageArray(est, spacing=stack_properties['spacing'], origin=stack_properties['origin'])


# ---------------------------------------------------------------------------------------------------
# a10-a12: DCT-entropy quality weights
# ---------------------------------------------------------------------------------------------------
def get_dct_options(spacing, size=None, max_kernel=None, gaussian_kernel=None):
    """mv:3944-3958."""
    from . import dct
    return dct.get_dct_options(spacing, size, max_kernel, gaussian_kernel)


def determine_chunk_quality(vrs, orig_stack_propertiess, params, stack_properties, block_info=None):
    """DCT Shannon entropy of one (V, s, s, s) block, normalised over the views that see it; NaN for the others
    (mv:4459-4561).  ``stack_properties`` carries the quality-grid spacing/origin; the block origin comes from
    ``block_info[0]['array-location']`` exactly like the reference."""
    from . import dct
    vrs = np.asarray(vrs)
    curr_origin = []
    for i in range(3):
        pixel_offset = block_info[0]['array-location'][i + 1][0]
        curr_origin.append(stack_properties['origin'][i] + pixel_offset * stack_properties['spacing'][i])
    bprops = {'size': np.array(vrs[0].shape).astype(np.int64), 'spacing': np.array(stack_properties['spacing']),
              'origin': np.array(curr_origin)}
    inside = blocks_inside(orig_stack_propertiess, params, bprops, n_points_per_dim=2)
    nv = len(inside)
    if not np.max(inside):
        return (np.ones(nv).astype(np.float32) * np.nan)[:, None, None, None]
    if np.sum(inside > 0) == 1:
        ws = np.ones(nv).astype(np.float32) * np.nan
        ws[inside > 0] = 1.
        return ws[:, None, None, None]
    size = int(vrs.shape[1])
    if vrs.shape[1:] != (size, size, size):
        raise Exception("determine_chunk_quality expects cubic blocks")
    qs = []
    for v in np.where(inside > 0)[0]:
        q = dct.block_quality_device(fusion.to_device(vrs[v]), 1, size, [1, 1, 1], [0])
        qs.append(q[0])
    ws = np.array(qs)
    if np.sum(ws) > 0:
        ws = ws / np.sum(ws)
    full = np.ones(nv) * np.nan
    full[inside > 0] = ws
    return full[:, None, None, None]


def get_weights_dct_dask(tviews, params, orig_stack_propertiess, stack_properties, depth=0, size=None, max_kernel=None,
                         gaussian_kernel=None, how_many_best_views=2, cumulative_weight_best_views=0.9):
    """DCT-quality x blending weights of every 128^3 chunk (mv:4032-4457).

    ``tviews``: (V, Z, Y, X) transformed views (numpy or dask), extents multiples of 128.  Returns the
    (V, nz*(128+2d), ...) float32 array whose (V, 128+2*depth, ...) chunks fuse_block consumes; a dask array with
    those chunks when dask is installed."""
    from . import dct
    arr = np.asarray(tviews)
    dviews = [fusion.to_device(arr[v]) for v in range(arr.shape[0])]
    chunks, _ = dct.weights_dct_chunks(dviews, params, orig_stack_propertiess, stack_properties, depth=depth,
                                       max_kernel=max_kernel, how_many_best_views=how_many_best_views,
                                       cumulative_weight_best_views=cumulative_weight_best_views)
    edge = dct.CHUNK + 2 * depth
    nchunks = [s // dct.CHUNK for s in arr.shape[1:]]
    out = np.zeros([arr.shape[0]] + [n * edge for n in nchunks], dtype=np.float32)
    for loc, w in chunks.items():
        out[:, loc[0] * edge:(loc[0] + 1) * edge, loc[1] * edge:(loc[1] + 1) * edge, loc[2] * edge:(loc[2] + 1) * edge] = w.cpu().numpy()
    return _maybe_dask(out, (arr.shape[0], edge, edge, edge))


def get_weights_dct(views, params, orig_stack_propertiess, stack_properties, size=None, max_kernel=None,
                    gaussian_kernel=None, how_many_best_views=1, cumulative_weight_best_views=0.9, block_info=None,
                    array_info=None):
    """The graph uses this name only as an identity tag for the DCT mode (mv:3729, mv_graph.py:336); called
    directly it evaluates the DCT weights of the given (128-multiple) transformed views."""
    return get_weights_dct_dask(np.asarray(views), params, orig_stack_propertiess, stack_properties, depth=0,
                                max_kernel=max_kernel, how_many_best_views=how_many_best_views,
                                cumulative_weight_best_views=cumulative_weight_best_views)


# ---------------------------------------------------------------------------------------------------
# a16: block orchestration
# ---------------------------------------------------------------------------------------------------
def _live(values):
    """Indices whose maximum is positive: fuse_block keeps a view (or its weights) only if it has signal in the block."""
    return [i for i, x in enumerate(values) if x.max() > 0]


def _block_frame(stack_properties, block_shape, chunk_location, array_info):
    """stack_properties of one (overlapping) chunk: origin moved by ``chunk_location * chunksize - depth`` voxels."""
    shift = np.array(chunk_location, dtype=np.float64) * array_info['chunksize'] - array_info['depth']
    frame = dict(stack_properties)
    frame['size'] = np.array(block_shape)
    frame['origin'] = np.asarray(stack_properties['origin'], dtype=np.float64) + shift * np.asarray(stack_properties['spacing'])
    return frame


def fuse_block(tviews_block, weights, params, stack_properties, orig_stack_propertiess, array_info, weights_func,
               fusion_func, weights_kwargs, fusion_kwargs, block_info=None):
    """Fuse one (V, 128+2d, ...) block of transformed views (mv:3817-3916).

    Block semantics of the reference: views without signal in the block are dropped; with no or one view left the raw
    block of view 0 / of that view is the result; the block's frame follows from its chunk location; the weight and
    fusion functions are called by keyword and only receive ``orig_stack_propertiess`` when they name it; views whose
    weights vanish on the block are judged the same way a second time."""
    keep = _live(tviews_block)
    if len(keep) < 2:
        return tviews_block[keep[0] if keep else 0]
    block_views = tviews_block[keep]
    params = np.array(params)[keep]
    view_frames = [orig_stack_propertiess[i] for i in keep]
    frame = _block_frame(stack_properties, block_views[0].shape, block_info[0]['chunk-location'][1:], array_info)
    for kwargs, func in ((weights_kwargs, weights_func), (fusion_kwargs, fusion_func)):
        kwargs['stack_properties'] = frame
        kwargs['params'] = params
        if 'orig_stack_propertiess' in func.__code__.co_varnames:
            kwargs['orig_stack_propertiess'] = view_frames
    tviews = [ImageArray(t, spacing=frame['spacing'], origin=frame['origin']) for t in block_views]
    blending = weights is None and weights_func == get_weights_simple
    if blending and fusion_func in (fuse_views_weights, fuse_LR_with_weights_np) and len(keep) <= 8:
        # device-resident: same decisions, but the block's weights never travel over PCIe
        return _fuse_block_device(block_views, params, view_frames, frame, fusion_func, fusion_kwargs)
    if blending:
        weights = weights_func(**weights_kwargs)
    else:
        weights = [ImageArray(weights[i], spacing=frame['spacing'], origin=frame['origin']) for i in keep]
    weighted = _live(weights)
    if len(weighted) < 2:
        return block_views[weighted[0] if weighted else 0]
    return fusion_func(tviews, weights=weights, **fusion_kwargs)


def _fuse_block_device(block_views, params, view_frames, frame, fusion_func, fusion_kwargs):
    """The tail of fuse_block (mv:3876-3911) for blending weights, on device tensors: weights -> views whose weights
    vanish decide the trivial cases -> weighted average or Richardson-Lucy -> one D2H of the fused block."""
    from . import rl
    dims = [int(s) for s in frame['size']]
    vs = fusion.build_viewset(None, view_frames, params, frame, weights="blending")
    w = fusion.blend_weights(vs, dims, (0, 0, 0), normalise=True)
    weighted = np.where(w.reshape(w.shape[0], -1).amax(dim=1).cpu().numpy() > 0)[0]
    if len(weighted) < 2:
        return block_views[weighted[0] if len(weighted) else 0]
    dviews = [fusion.to_device(t) for t in block_views]
    dweights = list(w.unbind(0))
    if fusion_func == fuse_views_weights:
        return ImageArray(fusion.to_host(fusion.fuse_weighted(dviews, dweights)), spacing=frame['spacing'], origin=frame['origin'])
    opt = dict(fusion_kwargs)
    psfs = [get_psf(q, frame, opt.get('sz', 4), opt.get('sxy', 0.5)) for q in params]
    est = rl.RLPlan(dviews, dweights, psfs).run(num_iterations=opt.get('num_iterations', 25), tol=opt.get('tol', 5e-5))
    return ImageArray(fusion.to_host(est.to(torch.int32).to(torch.uint16)), spacing=frame['spacing'], origin=frame['origin'])


def _overlap_block(sources, loc, depth, chunk):
    """One chunk of every source with a ``depth`` halo (overlap_from_sources, mv:3662-3707): the halo comes from the
    neighbouring voxels inside the volume and from a reflection at its faces; chunks beyond the volume are zeros."""
    extent = np.array(sources[0].shape)
    first = np.array(loc) * chunk
    block = np.zeros((len(sources),) + (chunk + 2 * depth,) * 3, dtype=sources[0].dtype)
    if np.any(first >= extent):
        return block
    last = first + chunk
    window = tuple(slice(int(max(0, a - depth)), int(min(n, b + depth))) for a, b, n in zip(first, last, extent))
    # a face of the volume inside the halo -> reflect what is missing there
    missing = [(depth if a == 0 else 0, int(max(0, depth - (n - b)))) for a, b, n in zip(first, last, extent)]
    for i, src in enumerate(sources):
        part = np.array(src[window])
        block[i] = np.pad(part, missing, mode='reflect') if np.any(missing) else part
    return block


def _fuse_lazy_on_device(tviews, params, stack_properties, orig_stack_propertiess):
    """fuse_blockwise for blending weights + weighted average at overlap 0 when every transformed view is still
    lazy: the raw views go to the device once, one fused launch applies fuse_block's per-chunk decisions
    (mvregfus_b200.pipeline), the fused volume comes back once.  None if the inputs do not qualify."""
    from . import pipeline
    if not all(isinstance(t, LazyTransformedView) and t._value is None for t in tviews):
        return None
    size = tuple(int(s) for s in stack_properties['size'])
    for t, p, osp in zip(tviews, params, orig_stack_propertiess):
        same_frame = (t.shape == size and np.allclose(t.spacing, stack_properties['spacing']) and
                      np.allclose(t.origin, stack_properties['origin']))
        same_view = (np.array_equal(t.p, np.asarray(p, dtype=np.float64)) and np.allclose(t.stack.spacing, osp['spacing']) and
                     np.allclose(t.stack.origin, osp['origin']))
        if not (same_frame and same_view and t.dtype in (np.dtype(np.uint16), np.dtype(np.float32))):
            return None
    if len(tviews) > 8 or len({t.dtype for t in tviews}) != 1:
        return None
    host_dtype = np.uint16 if tviews[0].dtype == np.uint16 else np.float32
    if os.environ.get("MVF_ROW_STREAM", "1") != "0" and os.environ.get("MVF_BRICK_UPLOAD", "1") != "0":
        # volumes of several 128-plane rows: per-row source bricks on an upload stream one row ahead of the kernels, results
        # on a download stream - both PCIe directions busy at once
        fused = pipeline.fuse_volume_rows_from_host([t.stack for t in tviews], orig_stack_propertiess, params, stack_properties,
                                                    host_dtype)
        if fused is not None:
            return fused
    # Upload per view only the source brick the target box touches (the reference's per-chunk crop, mv:1552-1601): for a
    # z-slab of the fused frame that is a fraction of every view; whole views go up in one piece.
    from . import slab
    dviews, windows = [], []
    for t in tviews:
        vshape = np.asarray(np.asarray(t.stack).shape)
        matrix, offset = index_space_affine(t.p, t.stack.spacing, t.stack.origin, t.spacing, t.origin)
        lo, hi = slab.source_brick(matrix, offset, vshape, 0, size[0], size[1], size[2])
        lo = np.minimum(lo, vshape - 1)
        hi = np.maximum(hi, lo + 1)
        if os.environ.get("MVF_BRICK_UPLOAD", "1") == "0" or np.prod(hi - lo) >= 0.8 * np.prod(vshape):
            dviews.append(fusion.to_device(t.stack))
            windows.append(np.zeros(3, dtype=np.int64))
        else:
            dviews.append(fusion.upload_brick(t.stack, lo, hi))
            windows.append(lo)
    # fused in rows of 128 planes; every row leaves the device (in the dtype of the views, as the block loop assembles its
    # result) while the next one is being fused
    return pipeline.fuse_volume_blockwise(dviews, orig_stack_propertiess, params, stack_properties, windows=windows,
                                          host_dtype=host_dtype)


def _fuse_blocks_on_device(tviews, params, stack_properties, orig_stack_propertiess, depth, weights_func, fusion_func,
                           weights_kwargs, fusion_kwargs):
    """The block loop of fuse_blockwise with everything resident in HBM (mvregfus_b200.pipeline.fuse_blocks_device): any
    overlap, blending or DCT weights, weighted average or Richardson-Lucy.  None if the request is outside that set
    (custom weight / fusion functions, more than 8 views, mixed dtypes): the host loop below handles it."""
    from . import dct, pipeline
    if fusion_func not in (fuse_views_weights, fuse_LR_with_weights_np) or weights_func not in (get_weights_simple, get_weights_dct):
        return None
    if len(tviews) > 8 or len({np.dtype(t.dtype) for t in tviews}) != 1 or np.dtype(tviews[0].dtype) not in (np.dtype(np.uint16), np.dtype(np.float32)):
        return None
    if weights_func == get_weights_simple and weights_kwargs:
        return None
    extra = set(fusion_kwargs) - ({'num_iterations', 'sz', 'sxy', 'tol'} if fusion_func == fuse_LR_with_weights_np else set())
    if extra:
        return None
    dviews = []
    for t in tviews:
        if isinstance(t, LazyTransformedView) and t._value is None:   # resampled on the device, never materialised on the host
            matrix, offset = index_space_affine(t.p, t.stack.spacing, t.stack.origin, t.spacing, t.origin)
            dviews.append(fusion.affine_resample(fusion.to_device(t.stack), matrix, offset, list(t.shape), (0, 0, 0), rule="scipy"))
        else:
            dviews.append(fusion.to_device(np.asarray(t)))
    chunks = None
    if weights_func == get_weights_dct:
        shape = dviews[0].shape
        exp = [int(-(-s // dct.CHUNK)) * dct.CHUNK for s in shape]
        padded = []
        for d in dviews:
            if list(shape) == exp:
                padded.append(d)
                continue
            p = torch.zeros(exp, dtype=torch.int16 if d.dtype == torch.uint16 else d.dtype, device=d.device)
            p[:shape[0], :shape[1], :shape[2]] = d.view(torch.int16) if d.dtype == torch.uint16 else d
            padded.append(p.view(torch.uint16) if d.dtype == torch.uint16 else p)
        opts = dict(weights_kwargs)
        chunks, _ = dct.weights_dct_chunks(padded, params, orig_stack_propertiess, stack_properties, depth=depth,
                                           max_kernel=opts.get('max_kernel'),
                                           how_many_best_views=opts.get('how_many_best_views', 2),
                                           cumulative_weight_best_views=opts.get('cumulative_weight_best_views', 0.9))
    fused = pipeline.fuse_blocks_device(dviews, params, stack_properties, orig_stack_propertiess, depth,
                                        'wavg' if fusion_func == fuse_views_weights else 'rl', fusion_kwargs, chunks)
    return fusion.to_host(fused)


def fuse_blockwise_arrays(tviews, params, stack_properties, orig_stack_propertiess, fusion_block_overlap=None,
                          weights_func=None, fusion_func=None, weights_kwargs=None, fusion_kwargs=None):
    """The block loop of fuse_blockwise (mv:3709-3780) on in-memory transformed views: 128^3 chunks with a
    ``fusion_block_overlap`` halo, fuse_block per chunk, halo trimmed, cropped to the original shape."""
    weights_kwargs = dict(weights_kwargs or {})
    fusion_kwargs = dict(fusion_kwargs or {})
    weights_func = get_weights_simple if weights_func is None else weights_func
    fusion_func = fuse_views_weights if fusion_func is None else fusion_func
    depth = int(fusion_block_overlap or 0)
    chunk = 128
    if depth == 0 and weights_func == get_weights_simple and fusion_func == fuse_views_weights and not weights_kwargs \
            and not fusion_kwargs:
        fused = _fuse_lazy_on_device(tviews, params, stack_properties, orig_stack_propertiess)
        if fused is not None:
            return fused
    if os.environ.get("MVF_BLOCKWISE", "device") != "host":
        fused = _fuse_blocks_on_device(tviews, params, stack_properties, orig_stack_propertiess, depth, weights_func,
                                       fusion_func, weights_kwargs, fusion_kwargs)
        if fused is not None:
            return fused
    orig_shape = np.array(tviews[0].shape)
    nchunks = [int(-(-s // chunk)) for s in orig_shape]
    weights = None
    if weights_func == get_weights_dct:
        for k, v in zip(('size', 'max_kernel', 'gaussian_kernel'),
                        get_dct_options(stack_properties['spacing'][0], weights_kwargs.get('size'),
                                        weights_kwargs.get('max_kernel'), weights_kwargs.get('gaussian_kernel'))):
            weights_kwargs[k] = v
        expanded = np.zeros([len(tviews)] + [n * chunk for n in nchunks], dtype=np.asarray(tviews[0]).dtype)
        for iv, t in enumerate(tviews):
            expanded[iv, :orig_shape[0], :orig_shape[1], :orig_shape[2]] = np.asarray(t)
        weights = np.asarray(get_weights_dct_dask(expanded, params, orig_stack_propertiess, stack_properties,
                                                  depth=depth, **weights_kwargs))
    out = np.zeros([n * chunk for n in nchunks], dtype=np.asarray(tviews[0]).dtype)
    edge = chunk + 2 * depth
    for loc in np.ndindex(*nchunks):
        blk = _overlap_block(tviews, loc, depth, chunk)
        w = None
        if weights is not None:
            w = weights[:, loc[0] * edge:(loc[0] + 1) * edge, loc[1] * edge:(loc[1] + 1) * edge, loc[2] * edge:(loc[2] + 1) * edge]
        fused = np.asarray(fuse_block(blk, w, params, dict(stack_properties), orig_stack_propertiess,
                                      {'depth': depth, 'chunksize': chunk}, weights_func, fusion_func,
                                      dict(weights_kwargs), dict(fusion_kwargs),
                                      block_info={0: {'chunk-location': (0,) + tuple(loc)}}))
        if depth > 0:
            fused = fused[depth:-depth, depth:-depth, depth:-depth]
        out[loc[0] * chunk:(loc[0] + 1) * chunk, loc[1] * chunk:(loc[1] + 1) * chunk, loc[2] * chunk:(loc[2] + 1) * chunk] = fused
    return out[:orig_shape[0], :orig_shape[1], :orig_shape[2]]


def fuse_blockwise(fn, fns_tview, params, stack_properties, orig_stack_propertiess, fusion_block_overlap=None,
                   weights_func=None, fusion_func=None, weights_kwargs=None, fusion_kwargs=None):
    """mv:3637-3813: opens the per-view HDF5 'Data' sets, fuses block-wise and writes the result to ``fn``.

    File formats stay the reference's business (north_star): this needs h5py, and writes a plain HDF5 'Data' set
    unless the reference's ``mvregfus.imaris.da_to_ims`` is importable, in which case the .ims writer is used."""
    try:
        import h5py
    except ImportError as e:
        raise Exception("fuse_blockwise reads HDF5 transformed views and needs h5py (%s); use "
                        "fuse_blockwise_arrays for in-memory views" % e)
    tviews = [np.asarray(h5py.File(f, 'r')['Data']) if isinstance(f, str) else np.asarray(f) for f in fns_tview]
    result = fuse_blockwise_arrays(tviews, params, stack_properties, orig_stack_propertiess, fusion_block_overlap,
                                   weights_func, fusion_func, weights_kwargs, fusion_kwargs)
    import os
    if os.path.exists(fn):
        os.remove(fn)
    try:
        from mvregfus.imaris import da_to_ims
        import dask.array as da
        da_to_ims(da.from_array(result, chunks=(128, 128, 128)), fn)
    except ImportError:
        with h5py.File(fn, 'w') as f:
            f.create_dataset('Data', data=result)
    return result
