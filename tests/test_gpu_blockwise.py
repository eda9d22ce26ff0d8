"""fuse_block / fuse_blockwise control flow on the GPU wrappers against the oracle and the reference golden."""
import os

import numpy as np
import pytest

from oracle import mvregfus_oracle as O

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))


@pytest.mark.parametrize("tag", ["all", "one"])
@pytest.mark.parametrize("ftag", ["wavg", "lr"])
def test_fuse_block_matches_reference_golden(cuda, tag, ftag):
    from mvregfus_b200 import multiview as mv
    tv = G["block_tviews"].copy()
    if tag == "one":
        tv[0] = 0
        tv[2] = 0
    props = [{"size": np.array([60, 60, 60]), "spacing": np.ones(3), "origin": np.array([-10.0, -10.0, -10.0])}
             for _ in range(3)]
    spb = {"size": np.array([128, 128, 128]), "spacing": np.ones(3), "origin": np.zeros(3)}
    ff, kw = (mv.fuse_views_weights, {}) if ftag == "wavg" else \
        (mv.fuse_LR_with_weights_np, {"num_iterations": 2, "sz": 2, "sxy": 1, "tol": 5e-5})
    res = np.asarray(mv.fuse_block(tv, None, G["block_params"], dict(spb), props, {"depth": 4, "chunksize": 32},
                                   mv.get_weights_simple, ff, {}, dict(kw), block_info={0: {"chunk-location": (0, 0, 0, 0)}}))
    ref = G["block_%s_%s" % (tag, ftag)]
    assert res.shape == ref.shape and res.dtype == ref.dtype
    d = np.abs(res.astype(np.int64) - ref.astype(np.int64))
    if tag == "one":
        assert d.max() == 0  # raw copy of the only live view
    else:
        assert d.max() <= 3 and np.mean(d > 0) <= 5e-3, (d.max(), np.mean(d > 0))


def test_fuse_blockwise_arrays(cuda):
    """Two 128-chunks along x with an 8-voxel halo, blending weights + weighted average."""
    from mvregfus_b200 import multiview as mv
    from tests.golden.make_golden import blob_volume, case_params
    shape = (40, 48, 150)
    params = case_params((40, 48, 150), 2, seed=7, odd_extra_deg=0.0)
    params[1][:9] = params[0][:9]
    params[1][9:] = params[0][9:] + np.array([3.0, -2.0, 20.0])
    props = [{"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)} for _ in range(2)]
    sp = {"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)}
    tv = [blob_volume(shape, 300 + i, np.uint16, 0.3) for i in range(2)]
    ref = O.fuse_blockwise(tv, params, sp, props, fusion_block_overlap=8)
    got = mv.fuse_blockwise_arrays(tv, params, sp, props, fusion_block_overlap=8,
                                   weights_func=mv.get_weights_simple, fusion_func=mv.fuse_views_weights)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 2 and np.mean(d > 0) <= 5e-3, (d.max(), np.mean(d > 0))


def test_gpu_resident_blockwise_matches_block_loop(cuda):
    """One fused launch with per-chunk decisions == the reference's 128^3 block loop (oracle), including chunks
    where a view is empty (dropped / raw copy)."""
    import torch
    from mvregfus_b200 import fusion, pipeline
    from mvregfus_b200.image_array import ImageArray
    from tests.golden.make_golden import blob_volume, case_params
    vshape = (120, 100, 180)
    params = case_params(vshape, 3, seed=17, odd_extra_deg=4.0)
    params[1][9:] += np.array([10.0, 0.0, 60.0])
    props = [{"size": np.array(vshape), "spacing": np.ones(3), "origin": np.zeros(3)} for _ in range(3)]
    views = [blob_volume(vshape, 500 + i, np.uint16, 0.3) for i in range(3)]
    views[2][:, :, 90:] = 0          # view 2 has no signal in half of its volume: empty in some chunks
    views[0][:60] = 0
    sp = {"size": np.array([150, 120, 260]), "spacing": np.ones(3), "origin": np.array([-10.0, -8.0, -20.0])}
    tv = [O.transform_stack_dask_numpy(ImageArray(v), p, sp) for v, p in zip(views, params)]
    ref = O.fuse_blockwise(tv, params, sp, props, fusion_block_overlap=0)
    out = pipeline.fuse_volume_blockwise([fusion.to_device(v) for v in views], props, params, sp).cpu().numpy()
    assert out.shape == ref.shape
    d = np.abs(out.astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 3 and np.mean(d > 0) <= 5e-3, (d.max(), np.mean(d > 0))
    # the per-chunk decisions matter: whole-volume fusion without them differs visibly on this input
    vs = fusion.build_viewset([fusion.to_device(v) for v in views], props, params, sp, weights="blending")
    plain = fusion.fuse_blend(vs, [150, 120, 260]).cpu().numpy()
    assert np.mean(plain != ref) > np.mean(out != ref)


def _blockwise_case(dtype):
    from tests.golden.make_golden import blob_volume, case_params
    shape = (140, 100, 150)
    params = case_params(shape, 3, seed=11, odd_extra_deg=0.0)
    params[1][9:] = params[1][9:] + np.array([4.0, -3.0, 30.0])
    props = [{"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)} for _ in range(3)]
    sp = {"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)}
    tv = [blob_volume(shape, 700 + i, dtype, 0.3) for i in range(3)]
    tv[2][:, :, 60:] = 0        # view 2 is empty in the upper x chunk: dropped there
    tv[0][:20] = 0
    return tv, params, sp, props


@pytest.mark.parametrize("mode", ["blend_wavg", "blend_rl", "dct_wavg", "blend_rl_f32"])
def test_device_resident_block_loop_equals_host_block_loop(cuda, mode, monkeypatch):
    """fuse_blockwise_arrays with everything resident in HBM (overlap blocks with reflected halos, per-block decisions,
    weights, weighted average / RL, halo trim) == the per-block host loop around fuse_block, bit for bit, and both match
    the oracle's block loop (mv:3709-3780, 3817-3916)."""
    from mvregfus_b200 import multiview as mv
    dtype = np.float32 if mode.endswith("f32") else np.uint16
    tv, params, sp, props = _blockwise_case(dtype)
    wf = mv.get_weights_dct if mode.startswith("dct") else mv.get_weights_simple
    rl_kw = {"num_iterations": 2, "sz": 2, "sxy": 1, "tol": 5e-5}
    ff, kw = (mv.fuse_LR_with_weights_np, rl_kw) if "rl" in mode else (mv.fuse_views_weights, {})
    depth = 0 if mode.startswith("dct") else 6
    monkeypatch.setenv("MVF_BLOCKWISE", "device")
    got = np.asarray(mv.fuse_blockwise_arrays(tv, params, sp, props, fusion_block_overlap=depth, weights_func=wf,
                                              fusion_func=ff, fusion_kwargs=dict(kw)))
    monkeypatch.setenv("MVF_BLOCKWISE", "host")
    host = np.asarray(mv.fuse_blockwise_arrays(tv, params, sp, props, fusion_block_overlap=depth, weights_func=wf,
                                               fusion_func=ff, fusion_kwargs=dict(kw)))
    assert got.shape == host.shape == tuple(tv[0].shape) and got.dtype == host.dtype == dtype
    assert np.array_equal(got, host)
    if not mode.startswith("dct"):
        ref = O.fuse_blockwise(tv, params, sp, props, fusion_block_overlap=depth,
                               fusion_func=O.fuse_LR_with_weights_np if "rl" in mode else O.fuse_views_weights,
                               fusion_kwargs=dict(kw))
        d = np.abs(got.astype(np.int64) - np.asarray(ref).astype(np.int64))
        assert d.max() <= 3 and np.mean(d > 0) <= 5e-3, (d.max(), np.mean(d > 0))


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_slab_box_uploads_source_bricks_only(cuda, dtype, monkeypatch):
    """A z-slab of the fused frame through the drop-in socket (lazy transformed views -> fuse_blockwise_arrays): only the
    source brick of the slab is uploaded per view (z range of the 0/180 degree views, x range of the 90/270 degree ones -
    the reference's per-chunk crop, mv:1552-1601).  Same result, bit for bit, as with whole views on the device, and equal
    to the oracle's block loop on the slab frame."""
    from mvregfus_b200 import multiview as mv, synthetic
    from mvregfus_b200.image_array import ImageArray
    shape = (96, 80, 112)
    views, params, props = synthetic.make_views(shape, 4, dtype=dtype, seed=21)
    views = [ImageArray(v, spacing=p["spacing"], origin=p["origin"]) for v, p in zip(views, props)]
    sp = mv.calc_stack_properties_from_views_and_params(props, params, spacing=[1.0] * 3)
    size = [int(s) for s in sp["size"]]
    z0, z1 = 40, 72
    sp_slab = {"size": np.array([z1 - z0, size[1], size[2]]), "spacing": sp["spacing"],
               "origin": sp["origin"] + np.array([z0, 0, 0]) * sp["spacing"]}
    calls = []
    from mvregfus_b200 import fusion
    real = fusion.upload_brick
    monkeypatch.setattr(fusion, "upload_brick", lambda a, lo, hi, **k: (calls.append((tuple(lo), tuple(hi))), real(a, lo, hi, **k))[1])

    def run():
        tv = [mv.transform_stack_dask_numpy(v, p, stack_properties=sp_slab) for v, p in zip(views, params)]
        return np.asarray(mv.fuse_blockwise_arrays(tv, params, sp_slab, props))
    got = run()
    assert len(calls) == 4 and all(np.prod(np.array(hi) - np.array(lo)) < 0.8 * np.prod(shape) for lo, hi in calls)
    monkeypatch.setenv("MVF_BRICK_UPLOAD", "0")
    whole = run()
    assert len(calls) == 4 and got.dtype == whole.dtype == dtype and np.array_equal(got, whole)
    tvo = [O.transform_stack_dask_numpy(v, p, sp_slab) for v, p in zip(views, params)]
    ref = np.asarray(O.fuse_blockwise(tvo, params, sp_slab, props, fusion_block_overlap=0))
    d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 3 and np.mean(d > 0) <= 5e-3, (d.max(), np.mean(d > 0))


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_row_streamed_socket_equals_whole_upload(cuda, dtype, monkeypatch):
    """Volumes of several 128-plane rows through the drop-in socket: per-row source bricks on an upload stream one row
    ahead, kernels, results on a download stream (pipeline.fuse_volume_rows_from_host).  Bit for bit the result of the
    whole-view upload + one launch, and the oracle's block loop within the usual integer tolerance."""
    from mvregfus_b200 import multiview as mv, pipeline, synthetic
    from mvregfus_b200.image_array import ImageArray
    shape = (150, 40, 150)
    views, params, props = synthetic.make_views(shape, 4, dtype=dtype, seed=31)
    views[2][:, :, 100:] = 0          # a view without signal in part of the volume: per-chunk decisions differ between rows
    views = [ImageArray(v, spacing=p["spacing"], origin=p["origin"]) for v, p in zip(views, props)]
    sp = mv.calc_stack_properties_from_views_and_params(props, params, spacing=[1.0] * 3)
    assert int(sp["size"][0]) > 128
    if dtype == np.float32:   # a z-slab box of the fused frame that still spans two rows (what a rank of a 2-GPU run fuses)
        z0 = 9
        sp = {"size": np.array([int(sp["size"][0]) - z0 - 3, int(sp["size"][1]), int(sp["size"][2])]), "spacing": sp["spacing"],
              "origin": sp["origin"] + np.array([z0, 0, 0]) * sp["spacing"]}
        assert int(sp["size"][0]) > 128
    used = []
    real = pipeline.fuse_volume_rows_from_host
    monkeypatch.setattr(pipeline, "fuse_volume_rows_from_host", lambda *a, **k: (used.append(1), real(*a, **k))[1])

    def run():
        tv = [mv.transform_stack_dask_numpy(v, p, stack_properties=sp) for v, p in zip(views, params)]
        return np.asarray(mv.fuse_blockwise_arrays(tv, params, sp, props))
    got = run()
    assert used == [1]
    monkeypatch.setenv("MVF_ROW_STREAM", "0")
    whole = run()
    assert used == [1] and got.dtype == whole.dtype == dtype and np.array_equal(got, whole)
    tvo = [O.transform_stack_dask_numpy(v, p, sp) for v, p in zip(views, params)]
    ref = np.asarray(O.fuse_blockwise(tvo, params, sp, props, fusion_block_overlap=0))
    d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 3 and np.mean(d > 0) <= 5e-3, (d.max(), np.mean(d > 0))


class _RecordingH5File:
    """In-memory stand-in for h5py.File (h5py is absent from the image): datasets land in a dict keyed by file name."""
    store = {}

    def __init__(self, fn, mode="r"):
        self.fn = fn
        if mode == "w":
            _RecordingH5File.store[fn] = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def create_dataset(self, name, data=None, **kwargs):
        _RecordingH5File.store[self.fn][name] = np.array(data)
        return _RecordingH5File.store[self.fn][name]

    def __getitem__(self, name):
        return _RecordingH5File.store[self.fn][name]


def test_file_based_entry_points_with_recording_h5py(cuda, monkeypatch):
    """transform_view_dask_and_save_chunked (mv:1746-1769) and the file-based fuse_blockwise (mv:3637-3813) hand HDF5 'Data'
    sets from one to the other; with a recording stand-in for h5py both run end to end and match the oracle."""
    import sys
    import types
    from mvregfus_b200 import multiview as mv
    from tests.helpers import small_case
    stub = types.ModuleType("h5py")
    stub.File = _RecordingH5File
    monkeypatch.setitem(sys.modules, "h5py", stub)
    _RecordingH5File.store.clear()
    views, params, props, sp = small_case(shape=(40, 36, 44), nviews=2, dtype=np.uint16, seed=5)
    fns = [mv.transform_view_dask_and_save_chunked("tview_%d.h5" % i, views[i], params, i, sp) for i in range(2)]
    assert fns == ["tview_0.h5", "tview_1.h5"]
    tvo = [np.asarray(O.transform_stack_dask_numpy(v, p, sp)) for v, p in zip(views, params)]
    for fn, ref in zip(fns, tvo):
        got = _RecordingH5File.store[fn]["Data"]
        d = np.abs(got.astype(np.int64) - ref.astype(np.int64))
        assert got.dtype == ref.dtype and d.max() <= 1 and np.mean(d > 0) <= 2e-3
    res = mv.fuse_blockwise("fused.h5", fns, params, sp, props, fusion_block_overlap=4,
                            weights_func=mv.get_weights_simple, fusion_func=mv.fuse_views_weights)
    stored = [_RecordingH5File.store[fn]["Data"] for fn in fns]
    ref = np.asarray(O.fuse_blockwise(stored, params, sp, props, fusion_block_overlap=4))
    d = np.abs(np.asarray(res).astype(np.int64) - ref.astype(np.int64))
    assert d.max() <= 2 and np.mean(d > 0) <= 5e-3, (d.max(), np.mean(d > 0))
    assert np.array_equal(_RecordingH5File.store["fused.h5"]["Data"], np.asarray(res))
