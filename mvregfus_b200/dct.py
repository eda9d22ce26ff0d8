"""DCT-entropy quality weights (reference: mvregfus/multiview.py:3944-3958, 4032-4457, 4459-4561).

The per-block DCT + entropy runs on the GPU (csrc/dct.cu); the decisions on the tiny coarse grid stay on the
host in float64 with the same scipy primitives the reference uses (nan-max filter, per-cell L-BFGS-B exponent
fit, nan-aware Gaussian), as SURVEY.md §7 recommends for bit compatibility.
"""
import ctypes as C

import numpy as np
import torch
from scipy import ndimage, optimize

from . import _cabi, _host_workers, _hostpool, fusion

CHUNK = 128  # the reference's fusion chunk size (mv:3716)


def get_dct_options(spacing, size=None, max_kernel=None, gaussian_kernel=None):
    """mv:3944-3958."""
    spacing = np.max([3, spacing])
    if size is None:
        size = np.max([4, int(50 / spacing)])
    if max_kernel is None:
        max_kernel = int(size / 2.)
    if gaussian_kernel is None:
        gaussian_kernel = int(max_kernel)
    return size, max_kernel, gaussian_kernel


def dct_binning(spacing0, chunk=CHUNK):
    """Bin factor and DCT block edge chosen at mv:4046-4072 (integers)."""
    bin_factor = 1
    relspacing = 3. / spacing0
    if relspacing > 1:
        candidates = np.array([1, 2, 4, 8, 16])
        bin_factor = int(candidates[np.where(candidates < relspacing)[0][-1]])
    size = chunk // bin_factor
    while size * (spacing0 * bin_factor) > 100 and size >= 4:
        size = size / 2
    return bin_factor, int(size)


def dct_matrix(n):
    """Orthonormal DCT-II matrix: scipy.fftpack.dct(type=2, norm='ortho') along one axis is x -> M @ x."""
    k = np.arange(n)[:, None]
    x = np.arange(n)[None, :]
    M = np.cos(np.pi * (2 * x + 1) * k / (2.0 * n)) * np.sqrt(2.0 / n)
    M[0] *= np.sqrt(0.5)
    return M


def _expanded(shape, chunk=CHUNK):
    return [int(-(-s // chunk) * chunk) for s in shape]


def block_quality_device(tview, bin_factor, size, nblocks, jobs):
    """Raw entropy q of the listed blocks of one transformed view (CUDA tensor) -> float64 numpy (flat grid)."""
    lib = _cabi.load()
    dev = tview.device
    nb = int(np.prod(nblocks))
    q = torch.zeros(nb, dtype=torch.float64, device=dev)
    if len(jobs) == 0:
        return q.cpu().numpy()
    grid = C.c_int(0)
    nbytes = lib.mvf_dct_scratch_bytes(int(size), C.byref(grid))
    if nbytes < 0:
        raise RuntimeError("DCT block size too large for the scratch buffer")
    scratch = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
    cosm = torch.from_numpy(dct_matrix(int(size))).to(dev)
    jt = torch.from_numpy(np.asarray(jobs, dtype=np.int32)).to(dev)
    _cabi.check(lib.mvf_dct_entropy(C.c_void_p(tview.data_ptr()), fusion.dtype_code(tview), _cabi.i32x3(tview.shape),
                                    int(bin_factor), int(size), _cabi.i32x3(nblocks), C.c_void_p(jt.data_ptr()), len(jobs),
                                    C.c_void_p(cosm.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                    C.c_void_p(q.data_ptr()), fusion._stream_ptr()))
    return q.cpu().numpy()


def quality_grid(tviews, params, orig_stack_propertiess, stack_properties):
    """Per-block quality of every view on the coarse grid: float64 (V, nbz, nby, nbx) with NaN for views that
    do not see a block (mv:4046-4100 + determine_chunk_quality).  ``tviews``: list of CUDA tensors."""
    spacing = np.asarray(stack_properties["spacing"], dtype=np.float64)
    b, size = dct_binning(spacing[0])
    exp = _expanded(tviews[0].shape)
    nb = [e // b // size for e in exp]
    nv = len(tviews)
    qspacing = spacing * b
    # which views see which block (2^3 probe points, bit-exact host arithmetic)
    locs = np.array(list(np.ndindex(*nb)), dtype=np.int64)
    props_np = [{k: np.asarray(v) for k, v in osp.items()} for osp in orig_stack_propertiess]
    origins = np.asarray(stack_properties["origin"], dtype=np.float64) + locs * size * np.array(qspacing)
    codes = _host_workers.block_codes_all(props_np, np.asarray(params, dtype=np.float64), origins,
                                          np.array([size] * 3).astype(np.int64), np.array(qspacing), 2)
    inside = np.ascontiguousarray(codes.T).reshape((nv,) + tuple(nb))
    seen = inside > 0
    nseen = seen.sum(0)
    ws = np.full((nv,) + tuple(nb), np.nan, dtype=np.float64)
    todo = seen & (nseen >= 2)[None]
    flat = np.arange(int(np.prod(nb))).reshape(nb)
    for v in range(nv):
        q = block_quality_device(tviews[v], b, size, nb, flat[todo[v]])
        ws[v][todo[v]] = q.reshape(nb)[todo[v]]
    # normalise over the views that see the block (mv:4548-4550); a single seeing view gets 1 (mv:4504-4508)
    tot = np.where(todo, ws, 0.0).sum(0)
    norm = np.where(tot > 0, tot, 1.0)
    ws = np.where(todo, ws / norm[None], ws)
    single = seen & (nseen == 1)[None]
    ws[single] = 1.0
    return ws, b, size


from ._host_workers import adapt_weights as _adapt_weights  # noqa: E402  (torch-free implementation, shared with the pool)


def _nan_gaussian(U, sigma):
    V = U.copy()
    V[U != U] = 0
    VV = ndimage.gaussian_filter(V, sigma=sigma, mode='nearest')
    W = 0 * U.copy() + 1
    W[U != U] = 0
    WW = ndimage.gaussian_filter(W, sigma=sigma, mode='nearest')
    Z = VV / WW
    Z[U != U] = np.nan
    return Z


def coarse_weights(ws, stack_properties, b, size, max_kernel=None, how_many_best_views=2,
                   cumulative_weight_best_views=0.9):
    """Coarse-grid heuristics (mv:4122-4253) -> (weights (V,nbz,nby,nbx), origin, spacing) of the coarse image."""
    spacing = np.asarray(stack_properties["spacing"], dtype=np.float64)
    if max_kernel is None:
        max_kernel = 100
    filter_size = np.max([1, int(max_kernel / (spacing[0] * size * b))])
    nanmask = np.isnan(ws)
    ws = np.array([ndimage.generic_filter(ws[i], function=np.nanmax, size=int(filter_size)) for i in range(len(ws))])
    ws[nanmask] = 0
    cells = np.ascontiguousarray(ws.reshape(ws.shape[0], -1).T)          # (ncells, V), one row per coarse cell
    tasks = [(cells[a:b2], how_many_best_views, cumulative_weight_best_views) for a, b2 in _hostpool.chunks(len(cells))]
    fitted = np.concatenate(_hostpool.pmap(_host_workers.adapt_cells_chunk, tasks, min_tasks=2 if len(cells) >= 256 else 10 ** 9))
    adapted = np.ascontiguousarray(fitted.T).reshape(ws.shape).astype(np.float64)
    ws = np.array([_nan_gaussian(adapted[i], filter_size / 2.) for i in range(len(adapted))])
    ws[nanmask] = 0
    origin = np.asarray(stack_properties["origin"]) + (spacing * (size * b - 1)) / 2.
    return ws, origin, spacing * size * b


def chunk_properties(stack_properties, chunk_loc, depth=0, chunk=CHUNK):
    """stack properties of one 128^3 (+2*depth) chunk as construct_weights builds them (mv:4271-4287)."""
    s0 = stack_properties["spacing"][0]
    origin = np.array([stack_properties["origin"][i] + chunk_loc[i] * s0 * chunk for i in range(3)])
    props = {"size": np.array([chunk] * 3).astype(np.int64) + 2 * depth, "spacing": np.array([s0] * 3)}
    props["origin"] = origin - depth * props["spacing"]
    return props


def weights_dct_chunks(tviews, params, orig_stack_propertiess, stack_properties, depth=0, max_kernel=None,
                       how_many_best_views=2, cumulative_weight_best_views=0.9):
    """get_weights_dct_dask (mv:4032-4457) evaluated eagerly: {chunk location: (V, 128+2d, ...) float32 CUDA tensor}."""
    ws, b, size = quality_grid(tviews, params, orig_stack_propertiess, stack_properties)
    coarse = coarse_weights(ws, stack_properties, b, size, max_kernel, how_many_best_views, cumulative_weight_best_views)
    exp = _expanded(tviews[0].shape)
    out = {}
    for loc in np.ndindex(*[e // CHUNK for e in exp]):
        bprops = chunk_properties(stack_properties, loc, depth)
        vs = fusion.build_viewset(None, orig_stack_propertiess, params, bprops, weights="dct", coarse=coarse)
        out[loc] = fusion.blend_weights(vs, [int(s) for s in bprops["size"]], normalise=True)
    return out, coarse
