"""ctypes binding of libmvregfus_b200.so (the C-ABI declared in include/mvregfus_b200.h).

There is deliberately no fallback: if the shared library has not been built, :func:`load` raises.
Build it with ``python -m mvregfus_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os
import threading

from . import build as _build

MVF_MAX_VIEWS = 8
MVF_SIG_N = 200
MVF_U16, MVF_F32 = 0, 1
MVF_RULE_SCIPY, MVF_RULE_ITK = 0, 1
MVF_W_OFF, MVF_W_BLEND, MVF_W_BLEND_DCT = 0, 1, 2

# every symbol include/mvregfus_b200.h declares (tests check the library exports all of them)
EXPORTS = (
    "mvf_abi_version", "mvf_last_error", "mvf_launch_count",
    "mvf_affine_resample", "mvf_blend_weights", "mvf_fuse_weighted", "mvf_fuse_blend",
    "mvf_chunk_probe", "mvf_fuse_blend_chunked",
    "mvf_rl_padded_size", "mvf_rl_init_estimate", "mvf_rl_blur", "mvf_rl_ratio", "mvf_rl_correct",
    "mvf_dct_scratch_bytes", "mvf_dct_entropy",
    "mvf_rl_fft_padded_dims", "mvf_rl_fft_plan_create", "mvf_rl_fft_plan_set_workspace", "mvf_rl_fft_plan_destroy",
    "mvf_rl_fft_psf_spectrum", "mvf_rl_fft_blur", "mvf_rl_fft_plan_create_zx", "mvf_rl_fft_plan_padded", "mvf_rl_fft_blur_zx",
    "mvf_rl_stream_supported", "mvf_rl_xy", "mvf_rl_blur_z", "mvf_rl_ratio_z", "mvf_rl_correct_z",
    "mvf_rl_zz", "mvf_rl_xy_correct", "mvf_copy_brick_h2d", "mvf_illum_fuse_planes",
    "mvf_subsample", "mvf_range_u16", "mvf_hist_u16", "mvf_background_u16", "mvf_bin_shrink",
)
MVF_RL_KMAX = 43


class MvfView(C.Structure):
    """struct mvf_view (include/mvregfus_b200.h)."""
    _fields_ = [
        ("data", C.c_void_p),
        ("dtype", C.c_int32),
        ("dims", C.c_int32 * 3),
        ("matrix", C.c_double * 9),
        ("offset", C.c_double * 3),
        ("weight_mode", C.c_int32),
        ("one_lo", C.c_int32 * 3),
        ("one_hi", C.c_int32 * 3),
        ("wmatrix", C.c_double * 9),
        ("woffset", C.c_double * 3),
        ("profiles", C.c_void_p),
        ("cgrid", C.c_void_p),
        ("cdims", C.c_int32 * 3),
        ("cscale", C.c_double * 3),
        ("coffset", C.c_double * 3),
    ]


class MvfRlPsf(C.Structure):
    """struct mvf_rl_psf: separable PSF factors (true-convolution taps)."""
    _fields_ = [("kz", C.c_float * 43), ("ky", C.c_float * 43), ("kx", C.c_float * 43), ("ksize", C.c_int32 * 3),
                ("dense", C.c_void_p)]


_lock = threading.Lock()
_lib = None


class MvfError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def _declare(lib):
    i32p = C.POINTER(C.c_int32)
    f64p = C.POINTER(C.c_double)
    lib.mvf_abi_version.restype = C.c_int
    lib.mvf_last_error.restype = C.c_char_p
    lib.mvf_launch_count.restype = C.c_int64
    lib.mvf_launch_count.argtypes = [C.c_int]
    lib.mvf_affine_resample.argtypes = [C.c_void_p, C.c_int, i32p, f64p, f64p, C.c_void_p, i32p, i32p, C.c_int, C.c_void_p]
    lib.mvf_blend_weights.argtypes = [C.c_int, C.POINTER(MvfView), C.c_void_p, i32p, i32p, C.c_int, C.c_void_p]
    lib.mvf_fuse_weighted.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]
    lib.mvf_fuse_blend.argtypes = [C.c_int, C.POINTER(MvfView), C.c_void_p, C.c_void_p, i32p, i32p, C.c_void_p]
    lib.mvf_chunk_probe.argtypes = [C.c_int, C.POINTER(MvfView), C.c_int, i32p, i32p, C.c_void_p, C.c_int, i32p, C.c_void_p]
    lib.mvf_fuse_blend_chunked.argtypes = [C.c_int, C.POINTER(MvfView), C.c_void_p, i32p, i32p, C.c_void_p, C.c_int, i32p,
                                           C.c_void_p]
    psfp = C.POINTER(MvfRlPsf)
    lib.mvf_rl_padded_size.argtypes = [C.c_int]
    lib.mvf_rl_init_estimate.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_void_p,
                                         C.c_size_t, C.c_void_p, C.c_void_p]
    lib.mvf_rl_blur.argtypes = [C.c_void_p, C.c_void_p, i32p, psfp, C.c_void_p, C.c_void_p]
    lib.mvf_rl_ratio.argtypes = [C.c_void_p, C.c_void_p, i32p, psfp, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mvf_rl_correct.argtypes = [C.c_void_p, C.c_void_p, i32p, psfp, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.mvf_rl_stream_supported.argtypes = [i32p, psfp]
    lib.mvf_rl_xy.argtypes = [C.c_void_p, C.c_int32, i32p, psfp, C.c_void_p, C.c_void_p]
    lib.mvf_rl_blur_z.argtypes = lib.mvf_rl_blur.argtypes
    lib.mvf_rl_ratio_z.argtypes = lib.mvf_rl_ratio.argtypes
    lib.mvf_rl_correct_z.argtypes = lib.mvf_rl_correct.argtypes
    lib.mvf_rl_zz.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, i32p, psfp, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mvf_rl_xy_correct.argtypes = [C.c_void_p, C.c_int32, i32p, psfp, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                      C.c_int64, C.c_void_p, C.c_void_p]
    lib.mvf_rl_fft_padded_dims.argtypes = [i32p, i32p, i32p]
    lib.mvf_rl_fft_plan_create.argtypes = [i32p, i32p, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    lib.mvf_rl_fft_plan_set_workspace.argtypes = [C.c_void_p, C.c_void_p]
    lib.mvf_rl_fft_plan_destroy.argtypes = [C.c_void_p]
    lib.mvf_rl_fft_psf_spectrum.argtypes = [C.c_void_p, C.c_void_p, i32p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mvf_rl_fft_blur.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, i32p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                    C.c_int64, C.c_void_p, C.c_void_p]
    lib.mvf_rl_fft_plan_create_zx.argtypes = lib.mvf_rl_fft_plan_create.argtypes
    lib.mvf_rl_fft_plan_padded.argtypes = [C.c_void_p, i32p]
    lib.mvf_rl_fft_blur_zx.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, i32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                       C.c_int64, C.c_void_p, C.c_void_p]
    lib.mvf_dct_scratch_bytes.argtypes = [C.c_int, C.POINTER(C.c_int)]
    lib.mvf_dct_entropy.argtypes = [C.c_void_p, C.c_int, i32p, C.c_int, C.c_int, i32p, C.c_void_p, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mvf_subsample.argtypes = [C.c_void_p, C.c_int, i32p, i32p, C.c_void_p, C.c_void_p]
    lib.mvf_range_u16.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.mvf_hist_u16.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mvf_background_u16.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
    lib.mvf_illum_fuse_planes.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p,
                                          C.c_void_p]
    lib.mvf_copy_brick_h2d.argtypes = [C.c_void_p, C.c_int, i32p, i32p, i32p, C.c_void_p, C.c_void_p]
    lib.mvf_bin_shrink.argtypes = [C.c_void_p, C.c_int, i32p, i32p, C.c_void_p, C.c_void_p]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("mvf_last_error", "mvf_launch_count", "mvf_abi_version", "mvf_rl_padded_size", "mvf_dct_scratch_bytes",
                        "mvf_rl_stream_supported"):
            fn.restype = C.c_int


def load():
    """Return the loaded library; raise if it was never built (no CPU fallback exists)."""
    global _lib
    with _lock:
        if _lib is None:
            path = os.environ.get("MVF_LIB") or _build.lib_path()  # MVF_LIB: kernel-variant experiments
            if not os.path.exists(path):
                if os.environ.get("MVF_AUTOBUILD") == "1":
                    _build.build()
                else:
                    raise RuntimeError(
                        "libmvregfus_b200.so is not built (%s). Run `python -m mvregfus_b200.build`; "
                        "this package has no CPU fallback." % path)
            lib = C.CDLL(path)
            _declare(lib)
            if lib.mvf_abi_version() != 1:
                raise RuntimeError("libmvregfus_b200.so ABI version mismatch")
            _lib = lib
    return _lib


def check(status):
    if status != 0:
        raise MvfError("mvregfus_b200 C-ABI error %d: %s" % (status, load().mvf_last_error().decode()))


def i32x3(v):
    return (C.c_int32 * 3)(*[int(x) for x in v])


def f64(v, n):
    return (C.c_double * n)(*[float(x) for x in v])


def launch_count(reset=False):
    return int(load().mvf_launch_count(1 if reset else 0))
