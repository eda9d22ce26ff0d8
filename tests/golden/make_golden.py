"""Generate golden vectors by executing the UNMODIFIED reference source in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz  (needs /root/reference)

The reference (``/root/reference/mvregfus/multiview.py``) cannot be imported as is: SimpleITK, dask, h5py,
h5pickle, skimage, ... are absent from the image and it targets numpy < 1.24.  This script imports it
anyway, with
  * inert stand-ins (``MagicMock``) for modules the fusion path never calls (h5py, skimage, networkx, ...);
  * a minimal *eager* stand-in for ``dask.array`` (zeros/empty/from_array/map_blocks/rechunk/compute with
    ``block_info`` and ``drop_axis``), enough to run ``affine_transform_dask``, ``get_weights_dct_dask`` and
    ``illumination_fusion_planewise`` unchanged;
  * a SimpleITK stand-in whose ``Resample``/``Clamp`` are the oracle's ITK RESTATEMENT
    (``oracle.mvregfus_oracle.itk_resample``).  Outputs that pass through it (blending weights, PSF rotation,
    coarse-weight upsampling) therefore pin the reference's *Python* arithmetic around ITK but NOT ITK's own
    rim/tie behaviour — that part stays "parity unpinned" (DESIGN.md);
  * numpy<1.24 compatibility: ``np.float``/``np.product``/``np.bool`` aliases, and ``numpy.fft.rfftn``
    promoted to float64 input as under numpy<2.
Nothing from the reference is copied into the repository; only its outputs on seeded inputs are stored.
"""
import importlib
import inspect
import os
import sys
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import mvregfus_oracle as O  # noqa: E402  (only for the ITK restatement behind the sitk stand-in)


# ---------------------------------------------------------------------------------------------------
# stand-ins
# ---------------------------------------------------------------------------------------------------
class _SitkImage:
    def __init__(self, arr, pixel_id=None):
        self.arr = np.asarray(arr)
        self.origin = (0.0,) * self.arr.ndim
        self.spacing = (1.0,) * self.arr.ndim
        self.pixel_id = self.arr.dtype if pixel_id is None else pixel_id

    def SetOrigin(self, o): self.origin = tuple(float(x) for x in o)
    def SetSpacing(self, s): self.spacing = tuple(float(x) for x in s)
    def GetOrigin(self): return self.origin
    def GetSpacing(self): return self.spacing
    def GetSize(self): return tuple(int(s) for s in self.arr.shape[::-1])
    def GetPixelID(self): return self.pixel_id


class _SitkTransform:
    def __init__(self, ndim, kind):
        self.params = None

    def SetParameters(self, p):
        self.params = np.array(p, dtype=np.float64)


def make_sitk_stub():
    m = types.ModuleType("SimpleITK")
    m.sitkAffine, m.sitkIdentity = "affine", "identity"
    m.sitkFloat32 = np.dtype(np.float32)
    m.sitkLinear, m.sitkNearestNeighbor, m.sitkBSpline = "linear", "nearest", "bspline"
    m.Transform = _SitkTransform
    m.CompositeTransform = lambda txs: list(txs)
    m.GetImageFromArray = lambda a: _SitkImage(a)
    m.GetArrayFromImage = lambda im: im.arr

    def cast(im, pid):
        out = _SitkImage(im.arr.astype(pid), pid)
        out.origin, out.spacing = im.origin, im.spacing
        return out

    def resample(im, shape, transf, interpolator, out_origin, out_spacing):
        assert len(transf) == 1, "hot-path callers pass exactly one affine (SURVEY appendix)"
        p_xyz = transf[0].params
        p_zyx = O.params_invert_coordinates(p_xyz)  # xyz -> zyx is the same flip
        res = O.itk_resample(im.arr.astype(np.float32), np.array(im.spacing)[::-1], np.array(im.origin)[::-1],
                             p_zyx, list(shape)[::-1], np.array(out_spacing)[::-1], np.array(out_origin)[::-1],
                             interp=interpolator)
        out = _SitkImage(res, np.dtype(np.float32))
        out.origin, out.spacing = tuple(out_origin), tuple(out_spacing)
        return out

    def clamp(im, pid):
        a = im.arr
        if np.issubdtype(np.dtype(pid), np.integer):
            info = np.iinfo(pid)
            a = np.clip(a, info.min, info.max).astype(pid)  # clamp, then truncate
        else:
            a = a.astype(pid)
        out = _SitkImage(a, pid)
        out.origin, out.spacing = im.origin, im.spacing
        return out

    m.Cast, m.Resample, m.Clamp = cast, resample, clamp
    return m


class EagerArray:
    """Just enough of dask.array.Array, evaluated eagerly."""

    def __init__(self, data, chunks):
        self.data = np.asarray(data)
        self.chunks = self._norm(chunks, self.data.shape)

    @staticmethod
    def _norm(chunks, shape):
        out = []
        for c, s in zip(chunks, shape):
            if isinstance(c, (tuple, list)):
                out.append(tuple(int(x) for x in c))
            else:
                c = int(c)
                full, rem = divmod(s, c)
                out.append(tuple([c] * full + ([rem] if rem else [])))
        return tuple(out)

    shape = property(lambda self: self.data.shape)
    dtype = property(lambda self: self.data.dtype)
    ndim = property(lambda self: self.data.ndim)
    chunksize = property(lambda self: tuple(max(c) for c in self.chunks))
    numblocks = property(lambda self: tuple(len(c) for c in self.chunks))

    def __len__(self): return len(self.data)
    def __getitem__(self, key): return self.data[key]
    def __array__(self, dtype=None, copy=None): return self.data if dtype is None else self.data.astype(dtype)
    def rechunk(self, chunks): return EagerArray(self.data, chunks)
    def compute(self, **kw): return self.data

    def map_blocks(self, func, *args, **kwargs):
        return map_blocks(func, self, *args, **kwargs)


def map_blocks(func, *args, dtype=None, chunks=None, drop_axis=None, **kwargs):
    arrays = [a for a in args if isinstance(a, EagerArray)]
    assert arrays and all(a is arrays[0] or a.numblocks == arrays[0].numblocks for a in arrays)
    first = arrays[0]
    wants_info = "block_info" in inspect.signature(func).parameters
    starts = [np.concatenate([[0], np.cumsum(c)]) for c in first.chunks]
    blocks = {}
    for loc in np.ndindex(*first.numblocks):
        sl = tuple(slice(int(starts[d][i]), int(starts[d][i + 1])) for d, i in enumerate(loc))
        call_args = [a.data[sl] if isinstance(a, EagerArray) else a for a in args]
        kw = dict(kwargs)
        if wants_info:
            loc_arr = [(s.start, s.stop) for s in sl]
            in_info = {"shape": first.shape, "num-chunks": first.numblocks, "array-location": loc_arr,
                       "chunk-location": tuple(loc)}
            out_shape = tuple(chunks) if chunks is not None else tuple(s.stop - s.start for s in sl)
            info = {i: in_info for i in range(len(arrays))}
            info[None] = {"chunk-shape": tuple(int(c) if not isinstance(c, (tuple, list)) else int(c[0]) for c in out_shape),
                          "chunk-location": tuple(loc), "num-chunks": first.numblocks, "dtype": dtype}
            kw["block_info"] = info
        blocks[loc] = np.asarray(func(*call_args, **kw))
    if drop_axis is not None:   # the function removes these axes from every block (one chunk along them): drop them from the grid
        drop = sorted([drop_axis] if np.isscalar(drop_axis) else list(drop_axis))
        assert all(first.numblocks[d] == 1 for d in drop)
        blocks = {tuple(i for d, i in enumerate(loc) if d not in drop): b for loc, b in blocks.items()}
        grid = tuple(g for d, g in enumerate(first.numblocks) if d not in drop)
        sample = blocks[(0,) * len(grid)]
        assert all(b.shape == sample.shape for b in blocks.values())
        out = np.empty(tuple(g * s for g, s in zip(grid, sample.shape)), dtype=sample.dtype)
        for loc, b in blocks.items():
            out[tuple(slice(i * s, (i + 1) * s) for i, s in zip(loc, sample.shape))] = b
        return EagerArray(out, sample.shape)
    # assemble like dask's concatenate3: dtype of the FIRST block, blocks written in place
    sample = blocks[(0,) * len(first.numblocks)]
    bshape = sample.shape
    grid = first.numblocks
    uniform = all(b.shape == bshape for b in blocks.values())
    if uniform:
        out = np.empty(tuple(g * s for g, s in zip(grid, bshape)), dtype=sample.dtype)
        for loc, b in blocks.items():
            out[tuple(slice(i * s, (i + 1) * s) for i, s in zip(loc, bshape))] = b
        return EagerArray(out, bshape)
    out = np.empty(first.shape, dtype=sample.dtype)
    for loc, b in blocks.items():
        out[tuple(slice(int(starts[d][i]), int(starts[d][i + 1])) for d, i in enumerate(loc))] = b
    return EagerArray(out, first.chunks)


def make_dask_stub():
    dask = types.ModuleType("dask")
    da = types.ModuleType("dask.array")
    da.zeros = lambda shape, dtype=float, chunks=None: EagerArray(np.zeros(tuple(int(s) for s in shape), dtype=dtype), chunks)
    da.empty = lambda shape, chunks=None, dtype=float: EagerArray(np.zeros(tuple(int(s) for s in shape), dtype=dtype), chunks)
    da.from_array = lambda x, chunks=None: EagerArray(x, chunks)
    da.map_blocks = map_blocks
    diag = types.ModuleType("dask.diagnostics")

    class ProgressBar:
        def __enter__(self): return self
        def __exit__(self, *a): return False

    diag.ProgressBar = ProgressBar
    dask.array, dask.diagnostics = da, diag
    dask.delayed = mock.MagicMock()
    return {"dask": dask, "dask.array": da, "dask.diagnostics": diag, "dask.delayed": dask.delayed}


def import_reference():
    """Import /root/reference/mvregfus/multiview.py with the stand-ins installed; returns the module."""
    for name, alias in (("float", float), ("product", np.prod), ("bool", bool), ("int", int)):
        if not hasattr(np, name):
            setattr(np, name, alias)
    _rfftn, _irfftn = np.fft.rfftn, np.fft.irfftn
    np.fft.rfftn = lambda a, *args, **kw: _rfftn(np.asarray(a, dtype=np.float64), *args, **kw)  # numpy<2 promotion
    stubs = {"SimpleITK": make_sitk_stub()}
    stubs.update(make_dask_stub())
    for name in ("h5py", "h5pickle", "skimage", "skimage.exposure", "skimage.filters", "tifffile", "czifile",
                 "aicspylibczi", "networkx", "bcolz", "distributed", "imagecodecs", "mvregfus.czifile",
                 "mvregfus.imaris", "mvregfus.tifffile"):
        stubs[name] = mock.MagicMock(name=name)
    sys.modules.update(stubs)
    sys.path.insert(0, REF)
    return importlib.import_module("mvregfus.multiview")


# ---------------------------------------------------------------------------------------------------
# seeded inputs (pure numpy so they can be regenerated anywhere)
# ---------------------------------------------------------------------------------------------------
def blob_volume(shape, seed, dtype=np.uint16, zero_fraction=0.4):
    rng = np.random.default_rng(seed)
    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, s) for s in shape], indexing="ij")
    v = 1500 * (1 + np.sin(5 * zz + seed) * np.cos(4 * yy) * np.sin(3 * xx + 0.5)) + 200 * rng.random(shape)
    r2 = (zz / 0.95) ** 2 + (yy / 0.9) ** 2 + (xx / 0.85) ** 2
    v = np.where(r2 <= np.quantile(r2, 1 - zero_fraction), v, 0.0)
    return v.astype(dtype)


def case_params(shape, nviews, seed, odd_extra_deg=7.0):
    rng = np.random.default_rng(seed)
    centre = (np.asarray(shape, dtype=np.float64) - 1) / 2
    ps = []
    for v in range(nviews):
        th = 2 * np.pi * v / nviews + (np.deg2rad(odd_extra_deg) if v % 2 else 0)
        c, s = np.cos(th), np.sin(th)
        A = np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]]) @ np.diag(1 + rng.uniform(-0.01, 0.01, 3))
        ps.append(np.concatenate([A.ravel(), centre - A @ centre + rng.uniform(-2, 2, 3)]))
    return np.array(ps)


def main():
    mv = import_reference()
    from mvregfus.image_array import ImageArray as RefImageArray
    out = {}

    # ---- G1: bounding boxes (integer, bit-exact) ----------------------------------------------------
    shape = (28, 22, 34)
    for nv, sp_out in ((2, 1.0), (4, 0.5), (3, 2.0)):
        params = case_params(shape, nv, seed=10 + nv)
        props = [{"size": np.array(shape), "spacing": np.array([1.5, 1.0, 1.0]), "origin": np.array([3.0, -2.0, 0.5])}
                 for _ in range(nv)]
        lo, hi = mv.get_sample_volume(props, params)
        sp = mv.calc_stack_properties_from_volume((lo, hi), np.array([sp_out] * 3))
        tag = "bbox_v%d" % nv
        out[tag + "_params"], out[tag + "_lower"], out[tag + "_upper"] = params, lo, hi
        out[tag + "_size"], out[tag + "_spacing"] = sp["size"], sp["spacing"]

    # ---- G2: scipy-path resampling through affine_transform_dask (chunked exactly like the graph) ------
    for dt, tag in ((np.uint16, "u16"), (np.float32, "f32")):
        params = case_params(shape, 2, seed=21)
        props = [{"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)} for _ in range(2)]
        views = [blob_volume(shape, 30 + i, dt) for i in range(2)]
        sp = mv.calc_stack_properties_from_volume(mv.get_sample_volume(props, params), np.array([1.0] * 3))
        for i in range(2):
            res = mv.transform_stack_dask_numpy(RefImageArray(views[i]), params[i], stack_properties=sp,
                                                interp="linear", chunksize=16)
            out["resample_%s_view%d" % (tag, i)] = views[i]
            out["resample_%s_out%d" % (tag, i)] = np.asarray(res.compute())
        out["resample_%s_params" % tag] = params
        out["resample_%s_size" % tag], out["resample_%s_origin" % tag] = sp["size"], sp["origin"]

    # ---- G3: blocks_inside + blending weights (ITK stand-in) ------------------------------------------
    params = case_params(shape, 3, seed=41)
    props = [{"size": np.array(shape), "spacing": np.array([2.0, 1.0, 1.0]), "origin": np.zeros(3)} for _ in range(3)]
    sp = mv.calc_stack_properties_from_volume(mv.get_sample_volume(props, params), np.array([1.0] * 3))
    out["weights_params"], out["weights_size"], out["weights_origin"] = params, sp["size"], sp["origin"]
    out["weights_blocks_inside5"] = mv.blocks_inside(props, params, sp, 5)
    out["weights_blocks_inside2"] = mv.blocks_inside(props, params, sp, 2)
    sub = dict(sp)
    sub["size"] = np.array([12, 12, 12])
    sub["origin"] = sp["origin"] + np.array([-30.0, 4.0, 4.0])
    out["weights_blocks_inside_far"] = mv.blocks_inside(props, params, sub, 5)
    ws = mv.get_weights_simple(props, params, sp)
    out["weights_simple"] = np.array([np.asarray(w) for w in ws]).astype(np.float32)

    # ---- G4: weighted fusion (truncation per view) ----------------------------------------------------
    for dt, tag in ((np.uint16, "u16"), (np.float32, "f32")):
        tviews = [blob_volume(tuple(int(s) for s in sp["size"]), 50 + i, dt, 0.3) for i in range(3)]
        f = mv.fuse_views_weights(tviews, params, sp, weights=ws)
        out["fuse_%s_tviews" % tag] = np.array(tviews)
        out["fuse_%s_out" % tag] = np.asarray(f)
        out["fuse_%s_mean" % tag] = np.asarray(mv.fuse_views_weights(tviews, params, sp, weights=None))

    # ---- G5: DCT options, chunk quality, full get_weights_dct_dask ------------------------------------
    out["dct_options"] = np.array([mv.get_dct_options(s) for s in (0.5, 1.0, 3.0, 6.0, 20.0)], dtype=np.float64)
    vshape = (96, 96, 96)
    params = case_params(vshape, 2, seed=61, odd_extra_deg=0.0)
    params[1][9:] += np.array([20.0, 0.0, 10.0])
    props = [{"size": np.array(vshape), "spacing": np.ones(3), "origin": np.zeros(3)} for _ in range(2)]
    sp = {"size": np.array([128, 128, 128]), "spacing": np.array([1.0, 1.0, 1.0]), "origin": np.array([-8.0, -16.0, -4.0])}
    tviews = np.array([blob_volume((128, 128, 128), 70 + i, np.uint16, 0.5) for i in range(2)])
    out["dct_params"], out["dct_origin"] = params, sp["origin"]
    out["dct_tviews_checksum"] = np.array([int(t.astype(np.int64).sum()) for t in tviews])
    blk = tviews[:, 32:64:1, 32:64, 32:64][:, ::1]
    bprops = {"size": np.array([32] * 3), "spacing": np.array([2.0] * 3), "origin": sp["origin"] + 32.0}
    q = mv.determine_chunk_quality(blk, props, params, bprops,
                                   block_info={0: {"array-location": [(0, 2), (0, 32), (0, 32), (0, 32)]}})
    out["dct_block_quality"] = np.asarray(q, dtype=np.float64).ravel()
    wd = mv.get_weights_dct_dask(EagerArray(tviews, (2, 128, 128, 128)), params, props, sp, depth=0,
                                 size=None, max_kernel=None, gaussian_kernel=None)
    wd = np.asarray(wd.compute())
    out["dct_weights_sub"] = wd[:, ::8, ::8, ::8].astype(np.float32)
    out["dct_weights_sum"] = np.array([float(np.sum(wd[i].astype(np.float64))) for i in range(2)])

    # ---- G6: PSF ------------------------------------------------------------------------------------------
    sp1 = {"size": np.array([40, 40, 40]), "spacing": np.array([1.0, 1.0, 1.0]), "origin": np.zeros(3)}
    p4 = case_params((40, 40, 40), 4, seed=81)
    out["psf_params"] = p4
    out["psf_sz3_sxy1"] = np.array([mv.get_psf(p, sp1, 3, 1) for p in p4])

    # ---- G7: multi-view RL ----------------------------------------------------------------------------------
    rl_shape = (36, 32, 40)
    sp_rl = {"size": np.array(rl_shape), "spacing": np.ones(3), "origin": np.zeros(3)}
    rl_params = case_params(rl_shape, 2, seed=91)
    rl_views = [blob_volume(rl_shape, 95 + i, np.uint16, 0.35) for i in range(2)]
    rng = np.random.default_rng(97)
    w0 = rng.random(rl_shape).astype(np.float32)
    w0[:6] = 0
    rl_w = [w0, (1 - w0).astype(np.float32)]
    est = mv.fuse_LR_with_weights_np(rl_views, rl_params, sp_rl, num_iterations=4, sz=2, sxy=1, tol=5e-5,
                                     weights=rl_w)
    out["rl_views"], out["rl_weights"], out["rl_params"] = np.array(rl_views), np.array(rl_w), rl_params
    out["rl_out_it4"] = np.asarray(est)

    # ---- G8: fuse_block control flow ----------------------------------------------------------------------------
    bshape = (40, 40, 40)
    params = case_params(bshape, 3, seed=101)
    props = [{"size": np.array([60, 60, 60]), "spacing": np.ones(3), "origin": np.array([-10.0, -10.0, -10.0])}
             for _ in range(3)]
    spb = {"size": np.array([128, 128, 128]), "spacing": np.ones(3), "origin": np.zeros(3)}
    tv = np.array([blob_volume(bshape, 110 + i, np.uint16, 0.3) for i in range(3)])
    tv_one = tv.copy()
    tv_one[0] = 0
    tv_one[2] = 0
    info = {0: {"chunk-location": (0, 0, 0, 0)}}
    for tag, blk in (("all", tv), ("one", tv_one)):
        for ftag, ff, fkw in (("wavg", mv.fuse_views_weights, {}),
                              ("lr", mv.fuse_LR_with_weights_np, {"num_iterations": 2, "sz": 2, "sxy": 1, "tol": 5e-5})):
            res = mv.fuse_block(blk.copy(), None, params, dict(spb), props, {"depth": 4, "chunksize": 32},
                                mv.get_weights_simple, ff, {}, dict(fkw), block_info=info)
            out["block_%s_%s" % (tag, ftag)] = np.asarray(res)
    out["block_tviews"], out["block_params"] = tv, params

    # ---- G9: dual-illumination fusion (mv:391-467, run unmodified through the eager map_blocks) ------
    import contextlib
    import io
    rng = np.random.default_rng(77)
    z, y, x = np.mgrid[:6, :44, :44]
    base = 220 + 180 * np.abs(np.sin((x + 2 * y + 5 * z) / 23.0)) * np.sin((y + 3) / 50.0 * np.pi) ** 2
    # fusion_axis counts axes of the (2, z, y, x) pair; only the default 2 runs: planes must be square (the reference indexes
    # rows with a per-column mask) and its final swap-back applies the same axis number to the 3-D result, where it is a
    # no-op for 2 (the planes come back transposed) and out of range for 3
    for tag, seed in (("a", 0), ("b", 1)):
        ramp = y / 43.0
        left = base * (0.35 + 0.65 * ramp) + rng.random(base.shape) * 12
        right = base * (1.0 - 0.65 * ramp) + rng.random(base.shape) * (12 + 30 * seed)
        pair = np.array([left, right]).astype(np.uint16)
        with contextlib.redirect_stdout(io.StringIO()):
            fused = mv.illumination_fusion_planewise(pair.copy(), fusion_axis=2)
        out["illum_%s_in" % tag], out["illum_%s_out" % tag] = pair, np.asarray(fused)

    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d arrays, %.1f KiB" % (path, len(out), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
