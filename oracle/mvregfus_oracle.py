"""CPU oracle for the MVRegFus view-fusion hot path.  **TEST INFRASTRUCTURE — NOT PRODUCT CODE.**

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; nothing under ``mvregfus_b200/`` does.  It restates, in numpy/scipy, the
arithmetic of the reference's own CPU path (``/root/reference/mvregfus/multiview.py`` = ``mv`` below).
Where the reference calls a third-party primitive that exists in this image (scipy.ndimage,
scipy.fftpack, numpy.fft, scipy.optimize) the oracle calls the *same primitive* the same way.

Pinning status (see DESIGN.md "Oracle"):
  * scipy-path resampling, bounding boxes, blocks_inside, DCT quality, weighted fusion, RL recurrence,
    fuse_block control flow: pinned against the UNMODIFIED reference source executed in the build
    container with stub modules for its missing imports (``tests/golden/make_golden.py``; fixtures under
    ``tests/golden/``).
  * ``sitk.Resample``/``sitk.Clamp`` semantics (SimpleITK is absent from the image and is *unpinned* in
    the reference, ``requirements.txt:17``): restated from the ITK 5 algorithm
    (``itk_resample`` below) — **parity unpinned for ITK-specific rim/tie behaviour**.
  * The reference was written against numpy<1.24 (``np.float``, ``np.product``; float64 ``rfftn`` of
    float32 input).  The oracle reproduces that arithmetic: ``float`` == float64, RL FFTs in float64.
"""
import numpy as np
from scipy import ndimage, optimize
from scipy.fftpack import dctn

from mvregfus_b200.image_array import ImageArray  # value type only (metadata carrier)

CHUNK = 128  # fusion chunk size hard-wired in the reference (mv:3716, mv_graph.py:73/503)


# --------------------------------------------------------------------------------------------------
# a2: parameter algebra (mv_utils.py:96-133, 280-285)
# --------------------------------------------------------------------------------------------------
def params_to_matrix(p):
    p = np.array(p, dtype=np.float64)
    M = np.eye(4)
    M[:3, :3] = p[:9].reshape(3, 3)
    M[:3, 3] = p[9:]
    return M


def matrix_to_params(M):
    return np.concatenate([np.asarray(M)[:3, :3].ravel(), np.asarray(M)[:3, 3]])


def invert_params(p):
    M = params_to_matrix(p)
    Ai = np.linalg.inv(M[:3, :3])
    R = np.eye(4)
    R[:3, :3] = Ai
    R[:3, 3] = -np.dot(Ai, M[:3, 3])
    return matrix_to_params(R)


def params_invert_coordinates(p):
    M = params_to_matrix(p)
    M[:3, :3] = M[:3, :3][::-1, ::-1]
    M[:3, 3] = M[:3, 3][::-1]
    return matrix_to_params(M)


# --------------------------------------------------------------------------------------------------
# a3: fused-frame bounding box (mv:3030-3052 == 3080-3104, 3106-3122, 4916-4932) — integer, bit-exact
# --------------------------------------------------------------------------------------------------
def get_sample_volume(props_list, params):
    """Back-project the 8 corners of every view through A^-1 (mv:3080-3104)."""
    corners = np.array(list(np.ndindex(2, 2, 2)))
    verts = np.zeros((len(props_list) * 8, 3))
    for iv, sp in enumerate(props_list):
        tmp = corners * np.array(sp["size"]) * np.array(sp["spacing"]) + np.array(sp["origin"])
        Minv = params_to_matrix(invert_params(params[iv]))
        verts[iv * 8:(iv + 1) * 8] = np.dot(Minv[:3, :3], tmp.T).T + Minv[:3, 3]
    return np.min(verts, 0), np.max(verts, 0)


get_union_volume = get_sample_volume  # byte-identical in the reference (mv:3030-3052)


def calc_stack_properties_from_volume(volume, spacing):
    """origin = lower; size = ceil(extent/spacing).astype(uint16) (mv:3106-3122; wraps above 65535)."""
    spacing = np.array(spacing).astype(np.float64)
    size = np.ceil((volume[1] - volume[0]) / spacing).astype(np.uint16)
    return {"size": size, "spacing": spacing, "origin": volume[0]}


def calc_stack_properties_from_views_and_params(views_props, params, spacing=None, mode="sample"):
    """mv:4916-4932 ('sample' and 'union' are the same computation)."""
    if mode not in ("sample", "union"):
        raise ValueError("mode %r: the reference's 'intersection' is broken (mv:3068)" % (mode,))
    return calc_stack_properties_from_volume(get_sample_volume(views_props, params), spacing)


# --------------------------------------------------------------------------------------------------
# a4/a5: scipy-path resampling (mv:1458-1519, 1524-1634)
# --------------------------------------------------------------------------------------------------
def index_space_affine(p, in_spacing, in_origin, out_spacing, out_origin):
    """matrix', offset' of mv:1486-1496."""
    p = np.array(p, dtype=np.float64)
    A, c = p[:9].reshape(3, 3), p[9:]
    Sx, Sy = np.diag(out_spacing), np.diag(in_spacing)
    matrix = np.dot(np.linalg.inv(Sy), np.dot(A, Sx))
    offset = np.dot(np.linalg.inv(Sy), c - np.asarray(in_origin) + np.dot(A, out_origin))
    return matrix, offset


def chunk_source_bbox(matrix, offset, chunk_offset, chunk_shape, input_shape):
    """Integer source bounding box of one output chunk (mv:1552-1571) -> (lo[3], hi[3]) int64."""
    corners = np.array(list(np.ndindex(2, 2, 2))) * np.array(chunk_shape) + np.array(chunk_offset)
    src = np.dot(matrix, corners.T).T + offset
    lo, hi = np.min(src, 0), np.max(src, 0)
    for d in range(3):
        lo[d] = np.clip(lo[d] - 2, 0, input_shape[d])
        hi[d] = np.clip(hi[d] + 2, 0, input_shape[d])
    return lo.astype(np.int64), hi.astype(np.int64)


def transform_chunk(view, matrix, offset, chunk_offset, chunk_shape):
    """One output chunk exactly as the reference computes it (mv:1546-1616)."""
    lo, hi = chunk_source_bbox(matrix, offset, chunk_offset, chunk_shape, view.shape)
    crop = view[tuple(slice(int(a), int(b)) for a, b in zip(lo, hi))]
    if not np.min(crop.shape):
        return np.zeros(chunk_shape, dtype=view.dtype)
    off_mod = offset + np.dot(matrix, np.array(chunk_offset)) - lo
    return ndimage.affine_transform(np.asarray(crop), matrix, off_mod, output_shape=tuple(chunk_shape), order=1)


def affine_transform_chunked(view, matrix, offset, out_shape, chunk=CHUNK):
    """affine_transform_dask (mv:1524-1634) evaluated eagerly, chunk by chunk."""
    out = np.zeros(tuple(int(s) for s in out_shape), dtype=view.dtype)
    for z0 in range(0, out.shape[0], chunk):
        for y0 in range(0, out.shape[1], chunk):
            for x0 in range(0, out.shape[2], chunk):
                shp = [min(chunk, out.shape[d] - o) for d, o in enumerate((z0, y0, x0))]
                out[z0:z0 + shp[0], y0:y0 + shp[1], x0:x0 + shp[2]] = transform_chunk(
                    view, matrix, offset, (z0, y0, x0), shp)
    return out


def transform_stack_dask_numpy(stack, p, stack_properties, chunksize=None):
    """mv:1458-1519 with eager evaluation.  ``chunksize=None`` = one whole-volume scipy call
    (identical to the chunked result up to ~1e-13 in ``offset_modified``)."""
    matrix, offset = index_space_affine(p, stack.spacing, stack.origin,
                                        stack_properties["spacing"], stack_properties["origin"])
    out_shape = tuple(int(s) for s in stack_properties["size"])
    if chunksize is None:
        return ndimage.affine_transform(np.asarray(stack), matrix, offset, output_shape=out_shape, order=1)
    return affine_transform_chunked(np.asarray(stack), matrix, offset, out_shape, chunksize)


# --------------------------------------------------------------------------------------------------
# a7: ITK-path resampling — RESTATEMENT of itk::ResampleImageFilter + LinearInterpolateImageFunction
#     (sitk.Resample called at mv:6868-6871; SimpleITK absent -> unpinned, see module docstring)
# --------------------------------------------------------------------------------------------------
def itk_continuous_index(p, in_spacing, in_origin, out_shape, out_spacing, out_origin):
    """Continuous view index of every output voxel: x=o_out+i*s_out; y=A x+c; u=(y-o_in)/s_in."""
    p = np.array(p, dtype=np.float64)
    A, c = p[:9].reshape(3, 3), p[9:]
    out_shape = [int(np.ceil(s)) for s in out_shape]  # mv:6825
    ax = [np.asarray(out_origin, np.float64)[d] + np.arange(out_shape[d], dtype=np.float64)
          * np.asarray(out_spacing, np.float64)[d] for d in range(3)]
    Z, Y, X = np.meshgrid(*ax, indexing="ij", sparse=True)
    u = []
    for d in range(3):
        y = A[d, 0] * Z + A[d, 1] * Y + A[d, 2] * X + c[d]
        u.append((y - in_origin[d]) / in_spacing[d])
    return u


def itk_resample(img, in_spacing, in_origin, p, out_shape, out_spacing, out_origin, interp="linear"):
    """ITK resampling: inside iff -0.5 <= u_d < n_d-0.5; linear = nested lerp x->y->z in float64 with
    neighbours clamped to the edge; nearest = floor(u+0.5); outside -> 0; float32 result, then
    sitk.Clamp back to the input dtype (truncation for integer types)."""
    img = np.asarray(img)
    src = img.astype(np.float32)  # sitk.Cast(..., sitkFloat32), mv:6869
    n = src.shape
    u = itk_continuous_index(p, in_spacing, in_origin, out_shape, out_spacing, out_origin)
    u = np.broadcast_arrays(*u)
    inside = np.ones(u[0].shape, dtype=bool)
    for d in range(3):
        inside &= (u[d] >= -0.5) & (u[d] < n[d] - 0.5)
    if interp == "nearest":
        idx = [np.clip(np.floor(u[d] + 0.5).astype(np.int64), 0, n[d] - 1) for d in range(3)]
        out = src[idx[0], idx[1], idx[2]].astype(np.float64)
    elif interp == "linear":
        b0, b1, t = [], [], []
        for d in range(3):
            base = np.maximum(np.floor(u[d]).astype(np.int64), 0)
            base = np.minimum(base, n[d] - 1)
            dist = np.maximum(u[d] - base, 0.0)
            b0.append(base)
            b1.append(np.minimum(base + 1, n[d] - 1))
            t.append(dist)
        s64 = src.astype(np.float64)

        def g(iz, iy, ix):
            return s64[iz, iy, ix]

        # x (fastest) first, then y, then z  — LinearInterpolateImageFunction::EvaluateOptimized (3-D)
        vx00 = g(b0[0], b0[1], b0[2]) + (g(b0[0], b0[1], b1[2]) - g(b0[0], b0[1], b0[2])) * t[2]
        vx10 = g(b0[0], b1[1], b0[2]) + (g(b0[0], b1[1], b1[2]) - g(b0[0], b1[1], b0[2])) * t[2]
        vx01 = g(b1[0], b0[1], b0[2]) + (g(b1[0], b0[1], b1[2]) - g(b1[0], b0[1], b0[2])) * t[2]
        vx11 = g(b1[0], b1[1], b0[2]) + (g(b1[0], b1[1], b1[2]) - g(b1[0], b1[1], b0[2])) * t[2]
        vxy0 = vx00 + (vx10 - vx00) * t[1]
        vxy1 = vx01 + (vx11 - vx01) * t[1]
        out = vxy0 + (vxy1 - vxy0) * t[0]
    else:
        raise ValueError("oracle restates only linear/nearest ITK interpolation (bspline/gaussian are "
                         "registration-only callers, out of scope)")
    out = np.where(inside, out, 0.0).astype(np.float32)
    if np.issubdtype(img.dtype, np.integer):  # sitk.Clamp(stack, orig dtype): clamp then truncate
        info = np.iinfo(img.dtype)
        out = np.clip(out, info.min, info.max).astype(img.dtype)
    elif img.dtype != np.float32:
        out = out.astype(img.dtype)
    return out


def transform_stack_sitk(stack, p=None, out_shape=None, out_spacing=None, out_origin=None, interp="linear",
                         stack_properties=None):
    """mv:1771-1824 on top of the ITK restatement.  Identity returns the input object (mv:1813)."""
    if p is None:
        p = matrix_to_params(np.eye(4))
    p = np.array(p, dtype=np.float64)
    if stack_properties is not None:
        out_shape, out_origin, out_spacing = (stack_properties["size"], stack_properties["origin"],
                                              stack_properties["spacing"])
    else:
        out_shape = stack.shape if out_shape is None else out_shape
        out_origin = stack.origin if out_origin is None else out_origin
        out_spacing = stack.spacing if out_spacing is None else out_spacing
    if (np.allclose(p, matrix_to_params(np.eye(4))) and np.allclose(out_shape, stack.shape)
            and np.allclose(out_origin, stack.origin) and np.allclose(out_spacing, stack.spacing)):
        return stack
    res = itk_resample(stack, stack.spacing, stack.origin, p, out_shape, out_spacing, out_origin, interp)
    return ImageArray(res, origin=out_origin, spacing=out_spacing)


# --------------------------------------------------------------------------------------------------
# a8/a9: probe points and blending weights (mv:3406-3466, 3469-3633)
# --------------------------------------------------------------------------------------------------
def _probe_inside(orig_props, p, stack_properties, n_points):
    rel = np.linspace(0, 1, n_points)
    pts = np.array([[i, j, k] for i in rel for j in rel for k in rel])
    phys = stack_properties["origin"] + pts * stack_properties["size"] * stack_properties["spacing"]
    p = np.asarray(p, dtype=np.float64)
    flags = []
    for point in phys:
        t_point = np.dot(p[:9].reshape((3, 3)), point) + p[9:]
        pix = (t_point - orig_props["origin"]) / orig_props["spacing"]
        flags.append(bool(np.all((pix >= 0) & (pix < np.asarray(orig_props["size"])))))
    return np.array(flags)


def blocks_inside(orig_stack_propertiess, params, stack_properties, n_points_per_dim=5):
    """1 = all probe points inside the view, 0 = none, 2 = partial (mv:3406-3466)."""
    codes = []
    for iview in range(len(params)):
        ins = _probe_inside(orig_stack_propertiess[iview], params[iview], stack_properties, n_points_per_dim)
        codes.append(1 if ins.all() else (0 if not ins.any() else 2))
    return np.array(codes)


def _sigmoid(x, borderwidth):
    x0 = float(borderwidth) / 2.0
    a = 12.0 / borderwidth
    return 1 / (1 + np.exp(-a * (x - x0)))


SIG_N = 200  # mv:3566


def blending_profiles(orig_props):
    """The three 1-D border profiles f_d (float32, length 200) whose outer ``min`` is the reference's
    200^3 sigmoid field, plus the field's spacing (mv:3558-3598).  Built by the same sequence of slice
    assignments, including Python's negative-index wrap for very wide borders."""
    sigspacing = (np.array(orig_props["size"]) - 2) / SIG_N * orig_props["spacing"]
    b = int(40.0 / sigspacing[0])  # z sig-spacing used for all axes (mv:3573)
    prof = []
    for d in range(3):
        f = np.ones(SIG_N, dtype=np.float32)
        bw = b * 4 if d == 0 else b
        for bx in range(bw):
            f[bx:bx + 1] = np.minimum(np.float32(_sigmoid(bx, bw)), f[bx:bx + 1])
        for bx in range(b):
            f[SIG_N - bx - 1:SIG_N - bx] = np.minimum(np.float32(_sigmoid(bx, b)), f[SIG_N - bx - 1:SIG_N - bx])
        prof.append(f)
    return prof, sigspacing


def get_weights_simple(orig_stack_propertiess, params, stack_properties, normalise=True):
    """Blending weights of every view on the target grid (mv:3469-3633)."""
    out_shape = tuple(int(s) for s in stack_properties["size"])
    ws = []
    for iview in range(len(params)):
        osp = orig_stack_propertiess[iview]
        if not _probe_inside(osp, params[iview], stack_properties, 5).any():
            ws.append(np.zeros(out_shape, dtype=np.float32))
            continue
        (f0, f1, f2), sigspacing = blending_profiles(osp)
        sig = np.minimum(np.minimum(f0[:, None, None], f1[None, :, None]), f2[None, None, :])
        sig = ImageArray(sig, spacing=sigspacing, origin=osp["origin"] + 1 * np.asarray(osp["spacing"]))
        ws.append(np.asarray(transform_stack_sitk(sig, params[iview], out_origin=stack_properties["origin"],
                                                  out_shape=stack_properties["size"],
                                                  out_spacing=stack_properties["spacing"], interp="linear")))
    if not normalise:
        return ws
    wsum = np.sum(ws, 0)
    wsum[wsum == 0] = 1
    return [w / wsum for w in ws]


# --------------------------------------------------------------------------------------------------
# a13: weighted additive fusion (mv:4879-4914)
# --------------------------------------------------------------------------------------------------
def fuse_views_weights(views, params, stack_properties, weights=None, views_in_target_space=True):
    if not views_in_target_space:
        views = [np.array(transform_stack_sitk(v, params[i], out_origin=stack_properties["origin"],
                                               out_shape=stack_properties["size"],
                                               out_spacing=stack_properties["spacing"]))
                 for i, v in enumerate(views)]
    if weights is not None:
        f = np.zeros_like(np.asarray(views[0]))
        for iw in range(len(views)):
            f += (np.asarray(weights[iw]) * np.asarray(views[iw]).astype(float)).astype(np.uint16)
    else:
        f = np.mean(views, 0)
    f = np.clip(f, 0, 2 ** 16 - 1)
    return ImageArray(f.astype(np.uint16), spacing=stack_properties["spacing"], origin=stack_properties["origin"])


def fuse_views_weights_float(views, weights):
    """Pre-truncation float64 value  sum_v w_v * I_v  (SURVEY §7 hard part 1: exposed for the 1e-5 test)."""
    acc = np.zeros(np.asarray(views[0]).shape, dtype=np.float64)
    for w, v in zip(weights, views):
        acc += np.asarray(w) * np.asarray(v).astype(np.float64)
    return acc


# --------------------------------------------------------------------------------------------------
# a10-a12: DCT-entropy quality weights (mv:3944-3958, 4032-4457, 4459-4561)
# --------------------------------------------------------------------------------------------------
def get_dct_options(spacing, size=None, max_kernel=None, gaussian_kernel=None):
    spacing = np.max([3, spacing])
    if size is None:
        size = np.max([4, int(50 / spacing)])
    if max_kernel is None:
        max_kernel = int(size / 2.0)
    if gaussian_kernel is None:
        gaussian_kernel = int(max_kernel)
    return size, max_kernel, gaussian_kernel


def dct_binning(spacing0, chunk=CHUNK):
    """(bin_factor, dct block size) chosen at mv:4046-4072.  ``size`` may be a float in the reference
    (true division); its value is always integral."""
    bin_factor = 1
    rel = 3.0 / spacing0
    if rel > 1:
        cand = np.array([1, 2, 4, 8, 16])
        bin_factor = int(cand[np.where(cand < rel)[0][-1]])
    size = int(chunk / bin_factor)
    while size * (spacing0 * bin_factor) > 100 and size >= 4:
        size = size / 2
    return bin_factor, size


def _abslog(x):
    res = np.zeros_like(x)
    x = np.abs(x)
    res[x > 0] = np.log2(x[x > 0])
    return res


def determine_chunk_quality(vrs, orig_stack_propertiess, params, block_stack_properties):
    """DCT Shannon entropy of one (V,s,s,s) block -> float64[V] with NaN for views outside (mv:4459-4561).
    ``block_stack_properties`` is the block's own {size, spacing, origin} (mv:4478-4494)."""
    inside = blocks_inside(orig_stack_propertiess, params, block_stack_properties, n_points_per_dim=2)
    nv = len(inside)
    if not np.max(inside):
        return np.ones(nv).astype(np.float32) * np.nan
    if np.sum(inside > 0) == 1:
        ws = np.ones(nv).astype(np.float32) * np.nan
        ws[inside > 0] = 1.0
        return ws
    vrs = np.copy(vrs)[inside > 0]
    ds = []
    for v in vrs:
        if np.sum(v == 0) > np.prod(v.shape) * (4 / 5.0):
            ds.append(np.array([0.0]))
            continue
        elif v.min() < 0.0001:
            v[v == 0] = v[v > 0].min()
        ds.append(dctn(v, norm="ortho", axes=[0, 1, 2]).flatten())
    l1 = np.array([np.sum(np.abs(d)) for d in ds])
    l1[l1 == 0] = 1
    ws = np.array([-np.sum(np.abs(d) * _abslog(d / l1[i])) for i, d in enumerate(ds)])
    wssum = np.sum(ws)
    if wssum > 0:
        ws = ws / wssum
    full = np.ones(nv) * np.nan
    full[inside > 0] = ws
    return full


def adapt_weights(ws, how_many_best_views, cumulative_weight_best_views):
    """Per-coarse-cell exponent fit (mv:4155-4210).  ws: float[V] -> float[V]."""
    ws = np.asarray(ws)
    wsum = np.sum(ws, 0)
    if wsum > 0:
        ws = ws / wsum
    if not (np.sum(ws > 0) >= 2):
        return ws
    wfs = np.sort(ws.astype(np.float64), axis=0)

    def energy(exp):
        tmpw = wfs ** exp[0]
        tmpw = tmpw / np.nansum(tmpw, 0)
        nsum = np.nansum(tmpw[-int(how_many_best_views):], (-1))
        return np.abs(np.nansum(nsum) - cumulative_weight_best_views)

    res = optimize.minimize(energy, [0.5], bounds=[[0.01, 100]], method="L-BFGS-B", options={"maxiter": 10})
    ws = np.array([ws[i] ** res.x[0] for i in range(len(ws))])
    wsum = np.sum(ws, 0)
    if wsum > 0:
        ws = ws / wsum
    return ws


def _nan_gaussian_filter(U, sigma):
    V = U.copy()
    V[U != U] = 0
    VV = ndimage.gaussian_filter(V, sigma=sigma, mode="nearest")
    W = 0 * U.copy() + 1
    W[U != U] = 0
    WW = ndimage.gaussian_filter(W, sigma=sigma, mode="nearest")
    Z = VV / WW
    Z[U != U] = np.nan
    return Z


def dct_quality_grid(tviews, params, orig_stack_propertiess, stack_properties):
    """Raw per-block quality on the coarse grid: float64 (V, nbz, nby, nbx) with NaNs (mv:4046-4100).
    ``tviews``: (V, Z, Y, X) transformed views, extents multiples of 128 (mv:3721-3727)."""
    tviews = np.asarray(tviews)
    spacing = np.asarray(stack_properties["spacing"], dtype=np.float64)
    b, size = dct_binning(spacing[0])
    size = int(size)
    binned = tviews[:, ::b, ::b, ::b]
    qprops = {"spacing": np.array(spacing * b), "size": np.array([size] * 3), "origin": stack_properties["origin"]}
    nb = [binned.shape[d + 1] // size for d in range(3)]
    ws = np.zeros((tviews.shape[0],) + tuple(nb), dtype=np.float64)
    for iz in range(nb[0]):
        for iy in range(nb[1]):
            for ix in range(nb[2]):
                loc = np.array([iz, iy, ix]) * size
                blk = binned[:, loc[0]:loc[0] + size, loc[1]:loc[1] + size, loc[2]:loc[2] + size]
                bprops = {"size": np.array(blk.shape[1:]).astype(np.int64), "spacing": np.array(qprops["spacing"]),
                          "origin": np.array([qprops["origin"][i] + loc[i] * qprops["spacing"][i] for i in range(3)])}
                ws[:, iz, iy, ix] = determine_chunk_quality(blk, orig_stack_propertiess, params, bprops)
    return ws, b, size


def dct_coarse_weights(ws, stack_properties, b, size, max_kernel=None, how_many_best_views=2,
                       cumulative_weight_best_views=0.9):
    """Host-side heuristics on the coarse grid (mv:4122-4253): nan-max filter, per-cell exponent fit,
    nan-aware Gaussian.  Returns (float array (V,nbz,nby,nbx), coarse origin, coarse spacing)."""
    spacing = np.asarray(stack_properties["spacing"], dtype=np.float64)
    if max_kernel is None:
        max_kernel = 100
    filter_size = np.max([1, int(max_kernel / (spacing[0] * size * b))])
    nanmask = np.isnan(ws)
    ws = np.array([ndimage.generic_filter(ws[i], function=np.nanmax, size=int(filter_size)) for i in range(len(ws))])
    ws[nanmask] = 0
    out = np.zeros(ws.shape, dtype=np.float64)  # map_blocks' dtype=float32 is only a hint; blocks stay float64
    for idx in np.ndindex(ws.shape[1:]):
        out[(slice(None),) + idx] = adapt_weights(ws[(slice(None),) + idx], how_many_best_views,
                                                  cumulative_weight_best_views)
    ws = np.array([_nan_gaussian_filter(out[i], filter_size / 2.0) for i in range(len(out))])
    ws[nanmask] = 0
    corigin = np.asarray(stack_properties["origin"]) + (spacing * (size * b - 1)) / 2.0
    cspacing = spacing * size * b
    return ws, corigin, cspacing


def construct_block_weights(coarse, corigin, cspacing, params, orig_stack_propertiess, stack_properties,
                            chunk_loc, depth=0, chunk=CHUNK):
    """Full-resolution weights of one 128^3(+2*depth) chunk (mv:4256-4317 + normalise mv:4434-4439)."""
    spacing0 = stack_properties["spacing"][0]
    in_spacing = spacing0 * chunk
    origin = np.array([stack_properties["origin"][i] + chunk_loc[i] * in_spacing for i in range(3)])
    bprops = {"size": np.array([chunk] * 3).astype(np.int64) + 2 * depth, "spacing": np.array([spacing0] * 3)}
    bprops["origin"] = origin - depth * bprops["spacing"]
    tmpws = get_weights_simple(orig_stack_propertiess, params, bprops)
    res = []
    for iw in range(len(coarse)):
        if tmpws[iw].max():
            wim = ImageArray(coarse[iw], origin=corigin, spacing=cspacing)
            res.append(np.asarray(transform_stack_sitk(wim, stack_properties=bprops, interp="linear")))
        else:
            res.append(np.zeros(bprops["size"]).astype(np.float32))
    ws = np.array(res) * np.array(tmpws)
    wssum = np.sum(ws, 0)
    wssum[wssum == 0] = 1
    return ws / wssum


def get_weights_dct(tviews, params, orig_stack_propertiess, stack_properties, depth=0, max_kernel=None,
                    how_many_best_views=2, cumulative_weight_best_views=0.9):
    """get_weights_dct_dask (mv:4032-4457) evaluated eagerly.  Returns {chunk_loc: (V,128+2d,...) array}."""
    q, b, size = dct_quality_grid(tviews, params, orig_stack_propertiess, stack_properties)
    coarse, corigin, cspacing = dct_coarse_weights(q, stack_properties, b, size, max_kernel,
                                                   how_many_best_views, cumulative_weight_best_views)
    nchunks = [np.asarray(tviews).shape[d + 1] // CHUNK for d in range(3)]
    out = {}
    for loc in np.ndindex(*nchunks):
        out[loc] = construct_block_weights(coarse, corigin, cspacing, params, orig_stack_propertiess,
                                           stack_properties, loc, depth)
    return out


# --------------------------------------------------------------------------------------------------
# a14/a15: PSF and multi-view Richardson-Lucy (mv:5376-5418, 5545-5749)
# --------------------------------------------------------------------------------------------------
def get_psf(p, stack_properties, sz, sxy):
    spacing = np.asarray(stack_properties["spacing"], dtype=np.float64)
    shape = (np.array([np.max([1, np.max([sz, sxy, sxy]) * 3])]) / spacing).astype(np.int64)
    shape = np.array([ps + [1, 0][int(ps % 2)] for ps in shape]).astype(np.int64)
    psf = np.zeros(shape, dtype=np.float32)
    psf[shape[0] // 2, shape[1] // 2, shape[2] // 2] = 1
    psf = ndimage.gaussian_filter(psf, np.array([sz, sxy, sxy]) / spacing)
    origin = -spacing * (shape // 2)
    psf = ImageArray(psf, spacing=spacing, origin=origin)
    tmpp = np.copy(np.asarray(p, dtype=np.float64))
    tmpp[-3:] = 0
    rot = transform_stack_sitk(psf, tmpp, out_origin=origin, out_shape=shape, out_spacing=spacing, interp="linear")
    rot = np.asarray(rot)
    rot = rot / np.sum(rot)
    return rot.astype(np.float32)


def blur_fft(img, psf_ft, kshape):
    """blur_with_ftpsfind (mv:5654-5673) with the FFT forced to float64 (numpy<2 behaviour)."""
    pad = np.pad(img, [[0, kshape[i] + 1] for i in range(3)], mode="reflect")
    out_shape = pad.shape
    ft = np.fft.rfftn(pad.astype(np.float64))
    ft = psf_ft * ft
    res = np.fft.irfftn(ft, s=out_shape, axes=(0, 1, 2))
    res = res[kshape[0] // 2:-kshape[0] // 2 - 1, kshape[1] // 2:-kshape[1] // 2 - 1, kshape[2] // 2:-kshape[2] // 2 - 1]
    res[res < 0] = 0
    return res


def rl_ext_index(n, K):
    """Index map of the reference's boundary handling, closed form: positions
    t=-R..n-1+R (R=(K-1)/2) -> source index.  t>=n reflects (no edge repeat); t<0 wraps into the far
    reflected tail: x[n-3-K-t]."""
    R = (K - 1) // 2
    t = np.arange(-R, n + R)
    return np.where(t < 0, n - 3 - K - t, np.where(t >= n, 2 * n - 2 - t, t))


def blur_direct(img, psf):
    """Independent direct evaluation of the same blur: out[i]=sum_j psf[j]*ext(i+R-j), clamp<0 -> 0."""
    img = np.asarray(img, dtype=np.float64)
    K = psf.shape
    idx = [rl_ext_index(img.shape[d], K[d]) for d in range(3)]
    ext = img[np.ix_(*idx)]
    out = ndimage.convolve(ext, np.asarray(psf, dtype=np.float64), mode="constant", cval=0.0)
    R = [(k - 1) // 2 for k in K]
    out = out[R[0]:R[0] + img.shape[0], R[1]:R[1] + img.shape[1], R[2]:R[2] + img.shape[2]]
    out[out < 0] = 0
    return out


def fuse_LR_with_weights_np(views, params, stack_properties, num_iterations=25, sz=4, sxy=0.5, tol=5e-5,
                            weights=None, psfs=None, blur="fft", return_float=False, return_iterations=False):
    """Multi-view RL with weights (mv:5545-5749).  dtype flow as under numpy<2: float32 storage for
    estimate/ratio/correction, float64 FFT products.  ``blur='direct'`` swaps in the closed-form
    convolution; ``return_float`` returns the float32 estimate before the uint16 cast."""
    if psfs is None:
        psfs = np.array([get_psf(params[ip], stack_properties, sz, sxy) for ip in range(len(params))])
    data = np.array(views)
    weights = np.asarray(weights)
    estimate = np.sum(weights * data, 0)
    curr_imsum = np.sum(estimate)
    masks = weights > 1e-5
    shape = estimate.shape
    kshape = psfs[0].shape
    psfs_ft = [np.fft.rfftn(np.pad(psfs[ip], [[0, shape[i] + 1] for i in range(3)], mode="constant").astype(np.float64))
               for ip in range(len(psfs))]

    def do_blur(img, ip):
        return blur_fft(img, psfs_ft[ip], kshape) if blur == "fft" else blur_direct(img, psfs[ip])

    expected = np.zeros(data.shape, dtype=np.float32)
    i = 0
    while 1:
        for ip in range(len(psfs)):
            expected[ip] = do_blur(estimate, ip)
        expected = (data / (expected + np.float32(1e-6))).astype(np.float32)
        expected *= masks
        correction = expected[0] * np.float32(0.0)
        for ip in range(len(psfs)):
            o = do_blur(expected[ip], ip)
            o = o * weights[ip]
            correction += o
        estimate = estimate * correction
        estimate = np.clip(estimate, 0, 2 ** 16 - 1)
        new_imsum = np.sum(estimate)
        conv = np.abs(1 - new_imsum / curr_imsum)
        if conv < tol and i >= 10:
            break
        if i >= num_iterations - 1:
            break
        curr_imsum = new_imsum
        i += 1
    if return_float:
        return (estimate, i + 1) if return_iterations else estimate
    return ImageArray(estimate.astype(np.uint16), spacing=stack_properties["spacing"], origin=stack_properties["origin"])


# --------------------------------------------------------------------------------------------------
# a16: block orchestration (mv:3637-3916), in-memory (no dask / HDF5)
# --------------------------------------------------------------------------------------------------
def overlap_block(sources, loc, depth, chunk=CHUNK):
    """overlap_from_sources (mv:3662-3707) for chunk location ``loc``: (V, chunk+2d, ...) with reflect
    padding at the volume edges."""
    srcshape = np.array(sources[0].shape)
    nv = len(sources)
    ys = np.zeros([nv] + [chunk + 2 * depth] * 3, dtype=sources[0].dtype)
    pads, slices = [], []
    for d in range(3):
        l, u = loc[d] * chunk, (loc[d] + 1) * chunk
        if l >= srcshape[d]:
            return ys
        padl = depth if l == 0 else 0
        udiff = srcshape[d] - u
        padu = depth - udiff if udiff < depth else 0
        pads.append([padl, padu])
        slices.append(slice(max(0, l - depth), min(srcshape[d], u + depth)))
    pads = np.array(pads)
    for iv in range(nv):
        y = np.array(sources[iv][tuple(slices)])
        if pads.max() > 0:
            y = np.pad(y, pads, mode="reflect")
        ys[iv] = y
    return ys


def fuse_block(tviews_block, weights, params, stack_properties, orig_stack_propertiess, chunk_loc, depth,
               weights_mode, fusion_func, fusion_kwargs, chunk=CHUNK):
    """mv:3817-3916.  ``weights_mode`` 'blending' computes get_weights_simple on the block; otherwise
    ``weights`` holds the precomputed (V, ...) block weights."""
    max_vals = np.array([t.max() for t in tviews_block])
    inds = np.where(max_vals > 0)[0]
    if len(inds) == 0:
        return tviews_block[0]
    if len(inds) == 1:
        return tviews_block[inds[0]]
    tviews_block = tviews_block[inds]
    params = np.array(params)[inds]
    osp = [orig_stack_propertiess[i] for i in inds]
    origin = np.array([stack_properties["origin"][i] + (chunk_loc[i] * chunk - depth) * stack_properties["spacing"][i]
                       for i in range(3)])
    bprops = dict(stack_properties)
    bprops["size"] = np.array(tviews_block[0].shape)
    bprops["origin"] = origin
    tviews = [ImageArray(t, spacing=bprops["spacing"], origin=bprops["origin"]) for t in tviews_block]
    if weights_mode == "blending":
        w = get_weights_simple(osp, params, bprops)
    else:
        w = [np.asarray(x) for i, x in enumerate(weights) if i in inds]
    wmax = np.array([x.max() for x in w])
    winds = np.where(wmax > 0)[0]
    if len(winds) == 0:
        return tviews_block[0]
    if len(winds) == 1:
        return tviews_block[winds[0]]
    return fusion_func(tviews, params=params, stack_properties=bprops, weights=w, **fusion_kwargs)


def fuse_blockwise(tviews, params, stack_properties, orig_stack_propertiess, fusion_block_overlap=0,
                   weights_mode="blending", fusion_func=fuse_views_weights, fusion_kwargs=None,
                   weights_kwargs=None, chunk=CHUNK):
    """mv:3637-3813 evaluated eagerly on in-memory transformed views (list of V arrays)."""
    fusion_kwargs = dict(fusion_kwargs or {})
    depth = int(fusion_block_overlap)
    orig_shape = np.array(tviews[0].shape)
    nchunks = [int(-(-s // chunk)) for s in orig_shape]
    out = np.zeros([n * chunk for n in nchunks], dtype=tviews[0].dtype)
    wblocks = None
    if weights_mode == "dct":
        expanded = np.stack([overlap_full(tviews, nchunks, chunk)[iv] for iv in range(len(tviews))])
        wblocks = get_weights_dct(expanded, params, orig_stack_propertiess, stack_properties, depth=depth,
                                  **(weights_kwargs or {}))
    for loc in np.ndindex(*nchunks):
        blk = overlap_block(tviews, loc, depth, chunk)
        w = None if wblocks is None else wblocks[loc]
        fused = np.asarray(fuse_block(blk, w, params, stack_properties, orig_stack_propertiess, loc, depth,
                                      weights_mode, fusion_func, fusion_kwargs, chunk))
        if depth > 0:
            fused = fused[depth:-depth, depth:-depth, depth:-depth]
        out[loc[0] * chunk:(loc[0] + 1) * chunk, loc[1] * chunk:(loc[1] + 1) * chunk,
            loc[2] * chunk:(loc[2] + 1) * chunk] = fused
    return out[:orig_shape[0], :orig_shape[1], :orig_shape[2]]


def overlap_full(tviews, nchunks, chunk=CHUNK):
    """The depth-0 expanded stack (zero-filled up to multiples of 128), mv:3721-3727."""
    full = np.zeros([len(tviews)] + [n * chunk for n in nchunks], dtype=tviews[0].dtype)
    s = tviews[0].shape
    for iv, t in enumerate(tviews):
        full[iv, :s[0], :s[1], :s[2]] = np.asarray(t)
    return full


# ---------------------------------------------------------------------------------------------------
# steps either side of the path
# ---------------------------------------------------------------------------------------------------
def subsample_data(data, subsamp):
    """imaris.py:764-765."""
    return data[0::int(subsamp[0]), 0::int(subsamp[1]), 0::int(subsamp[2])]


def pyramid_levels(array, subsamp=((1, 1, 1), (2, 2, 2), (4, 4, 4), (8, 8, 8))):
    """Level loop of np_to_ims (imaris.py:316-331): the factors are applied cumulatively, each level gets
    np.histogram(data, 256)."""
    data = np.squeeze(np.asarray(array))
    levels = []
    for ss in subsamp:
        if any(i > 1 for i in ss):
            data = subsample_data(data, ss)
        hist, edges = np.histogram(data, 256)
        levels.append({"data": data, "hist": hist, "edges": edges})
    return levels


def subtract_background(stack, background_level):
    """multiview.py:316: (stack - background_level) * (stack > background_level), numpy dtype rules of a
    uint16 stack and a Python int level (the difference wraps, the mask zeroes the wrapped entries)."""
    stack = np.asarray(stack)
    level = np.asarray(background_level).astype(stack.dtype) if np.issubdtype(stack.dtype, np.integer) else background_level
    return (stack - level) * (stack > level)


def bin_shrink(arr, factors_zyx):
    """RESTATEMENT of itk::BinShrinkImageFilter (SimpleITK is absent: parity unpinned): mean over the bins in the
    real type, partial bins dropped, integer outputs rounded half up (itk::Math::Round)."""
    arr = np.asarray(arr)
    f = [int(x) for x in factors_zyx]
    o = [n // k for n, k in zip(arr.shape, f)]
    a = arr[:o[0] * f[0], :o[1] * f[1], :o[2] * f[2]].astype(np.float64)
    a = a.reshape(o[0], f[0], o[1], f[1], o[2], f[2])
    # accumulation order of the filter: z, then y, then x offsets inside the bin
    s = np.zeros(o, dtype=np.float64)
    for dz in range(f[0]):
        for dy in range(f[1]):
            for dx in range(f[2]):
                s += a[:, dz, :, dy, :, dx]
    m = s / float(f[0] * f[1] * f[2])
    if np.issubdtype(arr.dtype, np.integer):
        return np.floor(m + 0.5).astype(arr.dtype)
    return m.astype(arr.dtype)


def bin_stack(im, bin_factors=np.array([1, 1, 1])):
    """mv_utils.py:288-311 with bin_shrink in place of sitk.BinShrink (factors arrive in x,y,z order)."""
    if np.allclose(bin_factors, [1] * len(bin_factors)):
        return im
    bin_factors = np.array(bin_factors)
    binned_spacing = im.spacing * bin_factors[::-1]
    binned_origin = im.origin + (binned_spacing - im.spacing) / 2.
    out = bin_shrink(np.asarray(im), [int(i) for i in bin_factors][::-1])
    return ImageArray(out, spacing=binned_spacing, origin=binned_origin, rotation=im.rotation)


def illumination_fusion_planewise(stack, fusion_axis=2):
    """Dual-illumination fusion, plane by plane (mv:391-467).  ``stack``: (2, z, y, x) uint16, the two light sheets.

    Per plane (after swapping ``fusion_axis`` with the last axis of the 4-D pair): mask = log(I0 + I1 - min), rescaled to
    [0, 1]; the column-wise cumulative sum of the mask decides where the second illumination takes over (past half of the
    column total); that binary map is smoothed with a Gaussian of sigma = plane shape / 50 and blends the two planes.
    Quirks kept: rows are selected with the per-COLUMN total (``mask[total == 0] = True``) and the transposes of the
    threshold test broadcast that total along the row index as well (square planes only); the swap
    back uses the same axis number on the 3-D result, a no-op for the default 2 - the planes come back transposed."""
    from scipy import ndimage
    stack = np.swapaxes(np.array(stack), fusion_axis, -1)
    out = np.empty(stack.shape[1:], dtype=np.uint16)
    for p in range(stack.shape[1]):
        planes = stack[:, p]
        mask = np.log(np.sum(planes, 0) - planes.min())
        mask = (mask - mask.min()) / (mask.max() - mask.min())
        total = np.sum(mask, axis=-2)
        mask[total == 0] = True
        total[total == 0] = mask.shape[1]
        cumsum = np.cumsum(mask, axis=-2)
        right_weight = (cumsum.T > total.T / 2.).T
        kernel = np.array(mask.shape) / 100 * 2.
        right_weight = ndimage.gaussian_filter(right_weight.astype(np.float32), kernel)
        out[p] = (planes[0] * (1 - right_weight) + planes[1] * right_weight).astype(np.uint16)
    return np.swapaxes(out, fusion_axis, -1)
