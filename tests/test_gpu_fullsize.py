"""Parity at BASELINE.json's full sizes through locality and size-independent properties.

Fusion is pointwise in the fused frame, so the CPU oracle can evaluate any sub-box of the full-size result on
its own (transform_chunk / get_weights_simple / fuse_views_weights on the sub-box); RL is checked by z-slab vs
whole-volume self-consistency at full size and against the oracle on a mid-size instance."""
import numpy as np
import pytest

from oracle import mvregfus_oracle as O

pytestmark = pytest.mark.gpu


def _config2(dtype):
    """BASELINE.json configs[1]: 4 views 512^3, generated on the device like bench.py does."""
    import bench
    import torch
    views, params, props, sp = bench.make_workload(torch.device("cuda:0"), 0, dtype="u16" if dtype == np.uint16 else "f32")
    return views, params, props, sp


@pytest.mark.parametrize("dtype", [np.float32, np.uint16])
def test_config2_full_size_against_oracle_subboxes(cuda, dtype):
    import torch
    from mvregfus_b200 import fusion
    views, params, props, sp = _config2(dtype)
    dims = [int(s) for s in sp["size"]]
    vs = fusion.build_viewset(views, props, params, sp, weights="blending")
    outf = torch.empty(dims, dtype=torch.float32, device="cuda")
    out = fusion.fuse_blend(vs, dims, out_f32=outf)
    # z-slab sharding at full size: 3 ragged slabs == whole volume, bit for bit
    cuts = [0, 171, 300, dims[0]]
    parts = [fusion.fuse_blend(vs, [b - a, dims[1], dims[2]], origin=(a, 0, 0)) for a, b in zip(cuts[:-1], cuts[1:])]
    assert torch.equal(torch.cat(parts, 0).view(torch.int16), out.view(torch.int16))
    out, outf = out.cpu().numpy(), outf.cpu().numpy()
    host_views = [v.cpu().numpy() for v in views]
    maps = [O.index_space_affine(params[v], props[v]["spacing"], props[v]["origin"], sp["spacing"], sp["origin"]) for v in range(4)]
    rng = np.random.default_rng(5)
    box = 40
    worst_f, worst_i, frac, frac_f = 0.0, 0, 0.0, 0.0
    for corner in ([60, 70, 80], [230, 240, 250], [dims[0] - box - 6, 200, dims[2] - box - 9], list(rng.integers(50, 400, 3))):
        off = np.array(corner)
        tv = [O.transform_chunk(host_views[v], maps[v][0], maps[v][1], off, [box] * 3) for v in range(4)]
        bprops = {"size": np.array([box] * 3), "spacing": sp["spacing"], "origin": sp["origin"] + off * sp["spacing"]}
        w = O.get_weights_simple(props, params, bprops)
        ref = np.asarray(O.fuse_views_weights(tv, params, bprops, weights=w))
        sl = tuple(slice(int(o), int(o) + box) for o in off)
        d = np.abs(out[sl].astype(np.int64) - ref.astype(np.int64))
        worst_i, frac = max(worst_i, int(d.max())), max(frac, float(np.mean(d > 0)))
        ref_f = O.fuse_views_weights_float(tv, w)
        rel = np.abs(outf[sl] - ref_f) / np.maximum(np.abs(ref_f), 1.0)
        if dtype == np.float32:
            worst_f = max(worst_f, float(rel.max()))
        else:   # uint16 views: every resampled voxel-view is rounded to an integer first (scipy's uint16 output); a value
            # within float32 rounding of a .5 boundary may land on the other integer, i.e. the sum moves by w_v <= 1
            worst_f = max(worst_f, float(np.abs(outf[sl] - ref_f).max()))
            frac_f = max(frac_f, float(np.mean(rel > 1e-5)))
    # measured on B200 (profiles/r02/parity_report_fullsize.txt): max |d| 1 LSB on 2.4e-5 of the voxels
    assert worst_i <= 2 and frac <= 5e-4, (worst_i, frac)
    if dtype == np.float32:
        assert worst_f <= 1e-5, worst_f   # north_star: 1e-5 relative on fused float32 voxels
    else:
        assert worst_f <= 1.0 and frac_f <= 1e-3, (worst_f, frac_f)


def test_rl_midsize_against_oracle(cuda):
    """4 axis-aligned views, 96x112x128, 7^3 PSF, 10 iterations: 1e-3 relative (floor 1 grey level) vs float64 FFT."""
    from mvregfus_b200 import fusion, rl, multiview as mv
    from tests.helpers import small_case
    views, params, props, sp = small_case(shape=(96, 112, 128), nviews=4, dtype=np.uint16, seed=11)
    tv = [O.transform_stack_dask_numpy(v, p, sp) for v, p in zip(views, params)]
    w = O.get_weights_simple(props, params, sp)
    ref = O.fuse_LR_with_weights_np(tv, params, sp, num_iterations=10, sz=2, sxy=1, weights=w, return_float=True)
    psfs = [mv.get_psf(p, sp, 2, 1) for p in params]
    got = rl.RLPlan([fusion.to_device(v) for v in tv], [fusion.to_device(x) for x in w], psfs).run(num_iterations=10).cpu().numpy()
    assert float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0))) <= 1e-3
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-3


def test_rl_full_size_slab_consistency(cuda):
    """BASELINE.json configs[3] size (4 views 512x1024x1024, 19^3 PSF): 2 emulated z-slab ranks == whole volume."""
    import torch
    from mvregfus_b200 import rl, slab
    shape, V = (512, 1024, 1024), 4
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    views = [(torch.rand(shape, generator=g, device="cuda") * 3000).to(torch.int32).to(torch.uint16) for _ in range(V)]
    w = [torch.rand(shape, generator=g, device="cuda") for _ in range(V)]
    k = np.exp(-0.5 * (np.arange(-9, 10) / 6.0) ** 2)
    kz = (k / k.sum()).astype(np.float32)
    k = np.exp(-0.5 * (np.arange(-9, 10) / 2.0) ** 2)
    kxy = (k / k.sum()).astype(np.float32)
    psfs = [kz[:, None, None] * kxy[None, :, None] * kxy[None, None, :]] * V
    whole = rl.RLPlan(views, w, psfs).run(num_iterations=2).clone()
    plans = [rl.SlabRLPlan([v[a:b] for v in views], [x[a:b] for x in w], psfs, shape[0], r, 2)
             for r, (a, b) in enumerate(slab.slab_bounds(shape[0], 2))]
    rl.setup_emulated(plans)
    for p in plans:
        p.init_estimate()
    for _ in range(2):
        rl.iterate_emulated(plans)
    assert torch.equal(torch.cat([p.est for p in plans], 0), whole)
