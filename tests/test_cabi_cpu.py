"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
the wrappers keep the reference signatures, host-side geometry matches the oracle.  No kernel is launched."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "mvregfus_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mvf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mvregfus_b200 import _cabi, build
    path = build.build()  # nvcc cross-compiles for sm_100a without a GPU
    lib = ctypes.CDLL(path)
    declared = _header_functions()
    assert declared, "no functions parsed from the header"
    for name in declared:
        assert hasattr(lib, name), "library lacks %s declared in include/mvregfus_b200.h" % name
    assert sorted(_cabi.EXPORTS) == declared, "ctypes binding and header disagree"
    lib.mvf_abi_version.restype = ctypes.c_int
    assert lib.mvf_abi_version() == 1
    lib.mvf_rl_padded_size.restype = ctypes.c_int
    assert [lib.mvf_rl_padded_size(k) for k in (1, 7, 9, 19, 21, 31, 33, 43, 45)] == [7, 7, 13, 19, 25, 31, 37, 43, -1]


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof of the C structs as gcc sees the header == the ctypes mirror."""
    import subprocess
    from mvregfus_b200 import _cabi
    src = tmp_path / "layout.c"
    src.write_text("""
#include <stdio.h>
#include <stddef.h>
#include "mvregfus_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu ", sizeof(mvf_view), offsetof(mvf_view, matrix), offsetof(mvf_view, weight_mode),
         offsetof(mvf_view, wmatrix), offsetof(mvf_view, profiles), offsetof(mvf_view, cdims), offsetof(mvf_view, cscale));
  printf("%zu %zu %zu", sizeof(mvf_rl_psf), offsetof(mvf_rl_psf, ksize), offsetof(mvf_rl_psf, dense));
  return 0;
}
""")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    V, P = _cabi.MvfView, _cabi.MvfRlPsf
    assert [int(x) for x in out[:7]] == [ctypes.sizeof(V), V.matrix.offset, V.weight_mode.offset, V.wmatrix.offset,
                                         V.profiles.offset, V.cdims.offset, V.cscale.offset]
    assert [int(x) for x in out[7:]] == [ctypes.sizeof(P), P.ksize.offset, P.dense.offset]


REFERENCE_SIGNATURES = {
    "transform_stack_sitk": "(stack, p=None, out_shape=None, out_spacing=None, out_origin=None, interp='linear', stack_properties=None)",
    "transform_stack_dask_numpy": "(stack, p=None, out_shape=None, out_spacing=None, out_origin=None, interp='linear', stack_properties=None, chunksize=512)",
    "affine_transform_dask": "(input, matrix, offset=0.0, output_shape=None, output_chunks=(128, 128, 128), **kwargs)",
    "get_weights_simple": "(orig_stack_propertiess, params, stack_properties)",
    "blocks_inside": "(orig_stack_propertiess, params, stack_properties, n_points_per_dim=5)",
    "get_dct_options": "(spacing, size=None, max_kernel=None, gaussian_kernel=None)",
    "determine_chunk_quality": "(vrs, orig_stack_propertiess, params, stack_properties, block_info=None)",
    "get_weights_dct_dask": "(tviews, params, orig_stack_propertiess, stack_properties, depth=0, size=None, max_kernel=None, gaussian_kernel=None, how_many_best_views=2, cumulative_weight_best_views=0.9)",
    "get_weights_dct": "(views, params, orig_stack_propertiess, stack_properties, size=None, max_kernel=None, gaussian_kernel=None, how_many_best_views=1, cumulative_weight_best_views=0.9, block_info=None, array_info=None)",
    "fuse_views_weights": "(views, params, stack_properties, weights=None, views_in_target_space=True)",
    "get_psf": "(p, stack_properties, sz, sxy)",
    "fuse_LR_with_weights_np": "(views, params, stack_properties, num_iterations=25, sz=4, sxy=0.5, tol=5e-05, weights=None)",
    "fuse_block": "(tviews_block, weights, params, stack_properties, orig_stack_propertiess, array_info, weights_func, fusion_func, weights_kwargs, fusion_kwargs, block_info=None)",
    "fuse_blockwise": "(fn, fns_tview, params, stack_properties, orig_stack_propertiess, fusion_block_overlap=None, weights_func=None, fusion_func=None, weights_kwargs=None, fusion_kwargs=None)",
    "calc_stack_properties_from_views_and_params": "(views_props, params, spacing=None, mode='sample')",
    "transform_view_dask_and_save_chunked": "(fn, view, params, iview, stack_properties, chunksize=128)",
}


@pytest.mark.parametrize("name", sorted(REFERENCE_SIGNATURES))
def test_wrapper_signatures_are_the_reference_ones(name):
    """same names, positional/keyword signatures; plain Python defs (fuse_block reads __code__)."""
    from mvregfus_b200 import multiview as mv
    fn = getattr(mv, name)
    assert str(inspect.signature(fn)) == REFERENCE_SIGNATURES[name]
    assert hasattr(fn, "__code__")
    if name in ("get_weights_simple", "get_weights_dct"):
        assert "orig_stack_propertiess" in fn.__code__.co_varnames  # fuse_block's introspection (mv:3864-3874)


def test_blending_profiles_match_oracle():
    from mvregfus_b200 import fusion
    from oracle import mvregfus_oracle as O
    for size, spacing in (((512, 512, 512), (1.0, 1.0, 1.0)), ((60, 300, 200), (2.0, 0.5, 0.5)), ((32, 40, 36), (1.0, 1.0, 1.0))):
        props = {"size": np.array(size), "spacing": np.array(spacing), "origin": np.zeros(3)}
        prof, sigsp = fusion.border_profiles(props)
        ref, refsp = O.blending_profiles(props)
        assert np.array_equal(sigsp, refsp)
        for d in range(3):
            assert prof.dtype == np.float32 and np.array_equal(prof[d], ref[d])


def test_identity_affine_returns_same_object():
    from mvregfus_b200 import multiview as mv
    from mvregfus_b200.image_array import ImageArray
    a = ImageArray(np.zeros((4, 5, 6), dtype=np.uint16), spacing=[2, 1, 1], origin=[1, 2, 3])
    assert mv.transform_stack_sitk(a) is a  # mv:1813


def test_image_array_contract():
    import pickle
    from mvregfus_b200.image_array import ImageArray
    a = ImageArray(np.arange(24, dtype=np.uint16).reshape(2, 3, 4), spacing=[2, 1, 1], origin=[1, 2, 3])
    b = a[1:, :2]
    assert isinstance(b, ImageArray) and np.array_equal(b.spacing, [2, 1, 1]) and np.array_equal(b.origin, [1, 2, 3])
    c = pickle.loads(pickle.dumps(a))
    assert np.array_equal(c, a) and np.array_equal(c.spacing, a.spacing) and np.array_equal(c.origin, a.origin)
    info = a.get_info()
    assert set(info) == {"spacing", "origin", "rotation", "size"} and np.array_equal(info["size"], [2, 3, 4])


def test_no_product_module_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "mvregfus_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_block_probe_matches_per_block_probe():
    """The all-blocks-at-once probe of the device pipeline decides exactly like the per-block probe loop
    (get_weights_simple's early-out, mv:3491-3532), jittered and exactly aligned geometries alike."""
    from mvregfus_b200 import synthetic
    from mvregfus_b200._host_workers import probe_any_inside_blocks, probe_flags
    shape = (100, 90, 140)
    for jitter in (True, False):
        params = synthetic.view_params(shape, 4, seed=11, jitter=jitter)
        props = {"size": np.array(shape), "spacing": np.array([1.0, 0.5, 0.5]), "origin": np.array([0.0, -3.0, 2.0])}
        sp = np.array([1.0, 1.0, 1.0])
        locs = np.array(list(np.ndindex(3, 3, 4)), dtype=np.float64)
        origins = np.array([-64.0, -128.0, 0.0]) + locs * 64 * sp
        for p in params:
            fast = probe_any_inside_blocks(props, p, origins, [64] * 3, sp, 5)
            slow = np.array([probe_flags(props, p, {"size": np.array([64] * 3), "spacing": sp, "origin": o}, 5).any()
                             for o in origins])
            assert np.array_equal(fast, slow)
