"""Device-level API of the fusion path: torch tensors own HBM, raw pointers cross the C-ABI.

Host-side geometry (float64) lives here: index-space affine maps (reference: mvregfus/multiview.py:
1486-1496), the probe-point early-out and the 1-D border profiles of the blending field
(mv:3491-3532, 3558-3604).  All voxel arithmetic happens in the CUDA kernels behind ``_cabi``.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi
from .mv_utils import index_space_affine

SIG_N = _cabi.MVF_SIG_N
_TORCH_DT = {np.dtype(np.uint16): (torch.uint16, _cabi.MVF_U16), np.dtype(np.float32): (torch.float32, _cabi.MVF_F32)}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("mvregfus_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def dtype_code(t):
    if t.dtype == torch.uint16:
        return _cabi.MVF_U16
    if t.dtype == torch.float32:
        return _cabi.MVF_F32
    raise TypeError("views must be uint16 or float32, got %s" % (t.dtype,))


TRANSFER = {"h2d_bytes": 0}   # bytes this process has sent to the device through to_device / upload_brick (bench accounting)


def to_device(arr, device=None):
    """numpy (uint16/float32, any layout) -> contiguous CUDA tensor.  Page-locked host memory (``pinned_empty``) is
    copied asynchronously on the current stream; pageable memory goes through the driver's staging copy."""
    require_cuda()
    arr = np.ascontiguousarray(np.asarray(arr))
    if arr.dtype not in _TORCH_DT:
        if np.issubdtype(arr.dtype, np.floating):
            arr = arr.astype(np.float32)
        else:
            raise TypeError("unsupported voxel dtype %s (uint16 and float32 are supported)" % arr.dtype)
    t = torch.from_numpy(arr)
    TRANSFER["h2d_bytes"] += arr.nbytes
    return t.to(device or "cuda", non_blocking=t.is_pinned())


def pinned_empty(shape, dtype):
    """Host ndarray in page-locked memory (the tensor that owns it stays alive through ``ndarray.base``): what a caller
    should hand to the drop-in entry points when PCIe time matters."""
    require_cuda()
    return torch.empty(tuple(int(s) for s in shape), dtype=_TORCH_DT[np.dtype(dtype)][0], pin_memory=True).numpy()


_HOST_POOL = {}          # (nbytes) -> idle page-locked uint8 tensors
_HOST_POOL_LIMIT = 8 << 30


def _host_buffer(nbytes):
    free = _HOST_POOL.get(nbytes)
    if free:
        return free.pop()
    return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)


def _host_release(raw):
    if sum(k * len(v) for k, v in _HOST_POOL.items()) + raw.numel() <= _HOST_POOL_LIMIT:
        _HOST_POOL.setdefault(raw.numel(), []).append(raw)


def to_host(t):
    """CUDA tensor -> host ndarray in page-locked memory, so the D2H copy runs at link speed instead of through the
    pageable staging path.  Page-locking is expensive (~0.2 s per GB): the buffer goes back to a pool when the returned
    array (and every view of it) has been garbage collected, and the next result of that size reuses it."""
    import weakref
    nbytes = t.numel() * t.element_size()
    raw = _host_buffer(nbytes)
    buf = raw.view(t.dtype).view(t.shape)
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    arr = buf.numpy()
    weakref.finalize(arr, _host_release, raw)
    return arr


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---------------------------------------------------------------------------------------------------
# host geometry
# ---------------------------------------------------------------------------------------------------
def probe_any_inside(orig_props, p, stack_properties, n_points=5):
    """True if any of the n^3 probe points of the target box falls inside the view (mv:3491-3532)."""
    return bool(_probe_flags(orig_props, p, stack_properties, n_points).any())


from ._host_workers import blocks_inside_codes, probe_flags as _probe_flags  # noqa: E402,F401  (torch-free implementation)


_PROFILE_CACHE = {}


def border_profiles(orig_props):
    """Cached :func:`_border_profiles` (the profiles depend on the view's size and spacing only)."""
    key = (tuple(int(x) for x in orig_props["size"]), tuple(float(x) for x in orig_props["spacing"]))
    hit = _PROFILE_CACHE.get(key)
    if hit is None:
        if len(_PROFILE_CACHE) > 256:
            _PROFILE_CACHE.clear()
        hit = _PROFILE_CACHE[key] = _border_profiles(orig_props)
    return hit


def _profile_tensor(orig_props, prof, dev):
    """Device copy of a view's border profiles, uploaded once per (view geometry, device)."""
    key = (tuple(int(x) for x in orig_props["size"]), tuple(float(x) for x in orig_props["spacing"]), str(dev))
    hit = _PROFILE_CACHE.get(key)
    if hit is None:
        hit = _PROFILE_CACHE[key] = torch.from_numpy(prof.ravel().copy()).to(dev)
    return hit


def _border_profiles(orig_props):
    """1-D profiles (3, 200) float32 of the sigmoid-border field and the field spacing.

    The reference fills a 200^3 array slab by slab with ``min(sigmoid(bx), slab)`` (mv:3577-3598); the
    field is therefore the outer minimum of three 1-D profiles, built here by the same slice walk
    (Python slice semantics included: very wide borders wrap through negative indices)."""
    size = np.array(orig_props["size"])
    sigspacing = (size - 2) / SIG_N * np.asarray(orig_props["spacing"], dtype=np.float64)
    b = int(40.0 / sigspacing[0])

    def sig(x, bw):
        return 1 / (1 + np.exp(-(12.0 / bw) * (x - float(bw) / 2.0)))

    prof = np.ones((3, SIG_N), dtype=np.float32)
    for d in range(3):
        f = prof[d]
        wide = 4 * b if d == 0 else b
        for bx in range(wide):
            seg = f[bx:bx + 1]
            seg[...] = np.minimum(np.float32(sig(bx, wide)), seg)
        for bx in range(b):
            seg = f[SIG_N - bx - 1:SIG_N - bx]
            seg[...] = np.minimum(np.float32(sig(bx, b)), seg)
    return prof, sigspacing


def _ones_range(f):
    nxt = np.minimum(np.arange(SIG_N) + 1, SIG_N - 1)
    ok = np.where((f == 1.0) & (f[nxt] == 1.0))[0]
    if len(ok) == 0 or ok[-1] - ok[0] + 1 != len(ok):
        return 0, 0
    return int(ok[0]), int(ok[-1]) + 1


class ViewSet:
    """The ctypes ``mvf_view`` array of one fusion call plus the device buffers it points to."""

    def __init__(self, n):
        self.structs = (_cabi.MvfView * n)()
        self.keep = []  # tensors kept alive while the structs are in use
        self.n = n


def build_viewset(tensors, orig_stack_propertiess, params, stack_properties, weights="blending",
                  coarse=None, view_dims=None, probe=True, windows=None):
    """Fill ``mvf_view`` for every view.  ``tensors`` may be None (weights only).  ``coarse`` is an optional
    ``(grids (V,cz,cy,cx) float32, origin[3], spacing[3])`` for the DCT mode.  ``weights`` in
    {'blending', 'dct', None}.  ``windows``: optional first index (z, y, x) of every tensor inside its view when only the
    source brick of the target box was uploaded (the reference's own per-chunk crop, mv:1552-1601): the data map is
    shifted by it, weights keep the geometry of the whole view."""
    n = len(params)
    if n > _cabi.MVF_MAX_VIEWS:
        raise ValueError("at most %d views per fused pass (got %d)" % (_cabi.MVF_MAX_VIEWS, n))
    vs = ViewSet(n)
    out_sp = np.asarray(stack_properties["spacing"], dtype=np.float64)
    out_or = np.asarray(stack_properties["origin"], dtype=np.float64)
    dev = tensors[0].device if tensors is not None else torch.device("cuda")
    for i in range(n):
        s = vs.structs[i]
        osp = orig_stack_propertiess[i]
        v_sp = np.asarray(osp["spacing"], dtype=np.float64)
        v_or = np.asarray(osp["origin"], dtype=np.float64)
        M, off = index_space_affine(params[i], v_sp, v_or, out_sp, out_or)
        s.matrix = _cabi.f64(M.ravel(), 9)
        s.offset = _cabi.f64(off if windows is None else off - np.asarray(windows[i], dtype=np.float64), 3)
        if tensors is not None:
            t = tensors[i]
            if not t.is_contiguous() or t.dim() != 3:
                raise ValueError("views must be contiguous 3-D CUDA tensors")
            s.data = t.data_ptr()
            s.dtype = dtype_code(t)
            s.dims = _cabi.i32x3(t.shape)
            vs.keep.append(t)
        else:
            s.dims = _cabi.i32x3(osp["size"] if view_dims is None else view_dims[i])
        s.weight_mode = _cabi.MVF_W_OFF
        if weights is None:
            continue
        if probe and not probe_any_inside(osp, params[i], stack_properties, 5):
            continue  # zero weights for this view in this box (mv:3518-3522)
        prof, sigspacing = border_profiles(osp)
        Mw, offw = index_space_affine(params[i], sigspacing, v_or + 1 * v_sp, out_sp, out_or)
        s.wmatrix = _cabi.f64(Mw.ravel(), 9)
        s.woffset = _cabi.f64(offw, 3)
        lo_hi = [_ones_range(prof[d]) for d in range(3)]
        s.one_lo = _cabi.i32x3([a for a, _ in lo_hi])
        s.one_hi = _cabi.i32x3([b for _, b in lo_hi])
        pt = _profile_tensor(osp, prof, dev)
        vs.keep.append(pt)
        s.profiles = pt.data_ptr()
        s.weight_mode = _cabi.MVF_W_BLEND
        if weights == "dct":
            grids, c_or, c_sp = coarse
            g = torch.from_numpy(np.ascontiguousarray(grids[i], dtype=np.float32)).to(dev)
            vs.keep.append(g)
            s.cgrid = g.data_ptr()
            s.cdims = _cabi.i32x3(g.shape)
            if np.any(out_sp / np.asarray(c_sp, dtype=np.float64) > 0.25 + 1e-12):
                # the kernels cache 4 coarse planes per 8-voxel column: DCT blocks of fewer than 4 fused voxels
                # (size * bin < 4, only for spacings beyond ~25 um) are not supported
                raise ValueError("DCT weight grid too fine for the fused kernels: block edge %s fused voxels (< 4)"
                                 % (np.asarray(c_sp, dtype=np.float64) / out_sp,))
            s.cscale = _cabi.f64(out_sp / np.asarray(c_sp, dtype=np.float64), 3)
            s.coffset = _cabi.f64((out_or - np.asarray(c_or, dtype=np.float64)) / np.asarray(c_sp, dtype=np.float64), 3)
            s.weight_mode = _cabi.MVF_W_BLEND_DCT
    return vs


# ---------------------------------------------------------------------------------------------------
# kernel launches on device tensors
# ---------------------------------------------------------------------------------------------------
def affine_resample(src, matrix, offset, out_dims, out_origin=(0, 0, 0), rule="scipy", out=None):
    """K1 on a CUDA tensor; returns a tensor of ``out_dims`` with the dtype of ``src``."""
    lib = _cabi.load()
    if out is None:
        out = torch.empty(tuple(int(d) for d in out_dims), dtype=src.dtype, device=src.device)
    rc = lib.mvf_affine_resample(
        C.c_void_p(src.data_ptr()), dtype_code(src), _cabi.i32x3(src.shape),
        _cabi.f64(np.asarray(matrix, dtype=np.float64).ravel(), 9), _cabi.f64(offset, 3),
        C.c_void_p(out.data_ptr()), _cabi.i32x3(out.shape), _cabi.i32x3(out_origin),
        _cabi.MVF_RULE_SCIPY if rule == "scipy" else _cabi.MVF_RULE_ITK, _stream_ptr())
    _cabi.check(rc)
    return out


def blend_weights(viewset, dims, origin=(0, 0, 0), normalise=True, device="cuda"):
    lib = _cabi.load()
    out = torch.empty((viewset.n,) + tuple(int(d) for d in dims), dtype=torch.float32, device=device)
    rc = lib.mvf_blend_weights(viewset.n, viewset.structs, C.c_void_p(out.data_ptr()), _cabi.i32x3(dims),
                               _cabi.i32x3(origin), 1 if normalise else 0, _stream_ptr())
    _cabi.check(rc)
    return out


def fuse_blend(viewset, dims, origin=(0, 0, 0), out=None, out_f32=None, device="cuda"):
    """K2: fused resample + weights + weighted sum of the box ``[origin, origin+dims)``."""
    lib = _cabi.load()
    if out is None:
        out = torch.empty(tuple(int(d) for d in dims), dtype=torch.uint16, device=device)
    rc = lib.mvf_fuse_blend(viewset.n, viewset.structs, C.c_void_p(out.data_ptr()),
                            C.c_void_p(out_f32.data_ptr()) if out_f32 is not None else None,
                            _cabi.i32x3(dims), _cabi.i32x3(origin), _stream_ptr())
    _cabi.check(rc)
    return out


def fuse_weighted(views, weights, out=None, out_f32=None):
    """fuse_views_weights on materialised CUDA tensors (list of V views, list of V float32 weights or None)."""
    lib = _cabi.load()
    n = len(views)
    if out is None:
        out = torch.empty(views[0].shape, dtype=torch.uint16, device=views[0].device)
    vp = (C.c_void_p * n)(*[v.data_ptr() for v in views])
    wp = (C.c_void_p * n)(*[w.data_ptr() for w in weights]) if weights is not None else None
    rc = lib.mvf_fuse_weighted(n, vp, dtype_code(views[0]), wp, C.c_void_p(out.data_ptr()),
                               C.c_void_p(out_f32.data_ptr()) if out_f32 is not None else None,
                               views[0].numel(), _stream_ptr())
    _cabi.check(rc)
    return out


def upload_brick(host_array, lo, hi, device="cuda"):
    """host_array[lo0:hi0, lo1:hi1, lo2:hi2] as a contiguous CUDA tensor, moved by ONE strided DMA (cudaMemcpy3DAsync
    inside mvf_copy_brick_h2d) straight out of the caller's (ideally page-locked) array - no host-side gather."""
    a = np.asarray(host_array)
    if a.ndim != 3 or not a.flags.c_contiguous or a.dtype not in (np.dtype(np.uint16), np.dtype(np.float32)):
        a = np.ascontiguousarray(a)
        if a.dtype not in (np.dtype(np.uint16), np.dtype(np.float32)):
            a = a.astype(np.float32)
    lo, hi = [int(x) for x in lo], [int(x) for x in hi]
    out = torch.empty([h - l for l, h in zip(lo, hi)], dtype=torch.uint16 if a.dtype == np.uint16 else torch.float32, device=device)
    _cabi.check(_cabi.load().mvf_copy_brick_h2d(C.c_void_p(a.ctypes.data), dtype_code(out), _cabi.i32x3(a.shape), _cabi.i32x3(lo),
                                                _cabi.i32x3(hi), C.c_void_p(out.data_ptr()), _stream_ptr()))
    out._mvf_host_keepalive = a   # the copy is asynchronous
    TRANSFER["h2d_bytes"] += out.numel() * out.element_size()
    return out
