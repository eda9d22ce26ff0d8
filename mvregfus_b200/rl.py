"""Multi-view Richardson-Lucy deconvolution on device tensors (reference: mvregfus/multiview.py:5545-5749).

Host side: PSF construction helpers (separability test), the z plane maps that encode the reference's boundary
handling, and the iteration loop.  The blurs and the fused pointwise algebra run in csrc/rl.cu.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _cabi
from .fusion import _stream_ptr, dtype_code


def factor_psf(psf, tol=2e-6):
    """(kz, ky, kx) with psf ~= kz (x) ky (x) kx, or None if the PSF is not separable to float32 accuracy.

    The PSF of an axis-aligned view (get_psf, mv:5376-5418) is a Gaussian filtered delta, i.e. an outer product;
    its factors are recovered from the marginals."""
    p = np.asarray(psf, dtype=np.float64)
    total = p.sum()
    if not np.isfinite(total) or total == 0:
        return None
    kz, ky, kx = p.sum((1, 2)), p.sum((0, 2)), p.sum((0, 1))
    approx = kz[:, None, None] * ky[None, :, None] * kx[None, None, :] / total ** 2
    if np.max(np.abs(approx - p)) > tol * np.max(np.abs(p)):
        return None
    return (kz / total).astype(np.float32), (ky / total).astype(np.float32), kx.astype(np.float32)


def factor_psf_y(psf, tol=2e-6):
    """(a, B) with psf ~= a(y) (x) B(z, x), or None.  A view rotated about the y axis has such a PSF: get_psf resamples a
    separable Gaussian with a map that leaves y alone (mv:5376-5418), so only the (z, x) part loses separability - and that
    part is far from low rank.  The convolution then needs FFTs over (z, x) only."""
    if os.environ.get("MVF_RL_FFT", "zx") == "3d":
        return None
    p = np.asarray(psf, dtype=np.float64)
    total = p.sum()
    if not np.isfinite(total) or total == 0 or p.shape[1] > _cabi.MVF_RL_KMAX or p.shape[1] % 2 == 0:
        return None
    a, B = p.sum((0, 2)), p.sum(1)
    if np.max(np.abs(a[None, :, None] * B[:, None, :] / total - p)) > tol * np.max(np.abs(p)):
        return None
    return (a / total).astype(np.float32), np.ascontiguousarray(B, dtype=np.float32)


def too_long_for_kernels(psf):
    """PSFs longer than MVF_RL_KMAX taps per axis (sz / spacing large) only fit the cuFFT convolution."""
    return int(max(np.asarray(psf).shape)) > _cabi.MVF_RL_KMAX


def fft_only_psf_struct(psf):
    """PSF struct of a view that always goes through cuFFT: only the sizes are used (index map of the z halo)."""
    s = _cabi.MvfRlPsf()
    s.ksize = _cabi.i32x3(np.asarray(psf).shape)
    return s


def make_dense_psf_struct(psf, device):
    """Non-separable PSF: centred, zero padded to Kp^3 and uploaded; returns (struct, tensor to keep alive)."""
    psf = np.asarray(psf, dtype=np.float32)
    kp = _cabi.load().mvf_rl_padded_size(int(max(psf.shape)))
    if kp < 0 or any(k % 2 == 0 for k in psf.shape):
        raise ValueError("PSF sizes must be odd and <= %d for the direct kernel" % _cabi.MVF_RL_KMAX)
    full = np.zeros((kp, kp, kp), dtype=np.float32)
    o = [(kp - k) // 2 for k in psf.shape]
    full[o[0]:o[0] + psf.shape[0], o[1]:o[1] + psf.shape[1], o[2]:o[2] + psf.shape[2]] = psf
    t = torch.from_numpy(full).to(device)
    s = _cabi.MvfRlPsf()
    s.ksize = _cabi.i32x3(psf.shape)
    s.dense = t.data_ptr()
    return s, t


class FftBlur:
    """cuFFT convolution state shared by the non-separable views of one RL problem: plan, work area, one real and
    one half-spectrum buffer, and one PSF spectrum per view."""

    def __init__(self, dims, ksize, device, zx=False):
        """zx: batched 2-D (z, x) transforms + direct y pass, for PSFs a(y) (x) B(z, x) (factor_psf_y)."""
        self.lib = _cabi.load()
        self.dims, self.ksize, self.zx = tuple(int(d) for d in dims), tuple(int(k) for k in ksize), bool(zx)
        nbytes = C.c_size_t(0)
        self.plan = C.c_void_p()
        create = self.lib.mvf_rl_fft_plan_create_zx if zx else self.lib.mvf_rl_fft_plan_create
        _cabi.check(create(_cabi.i32x3(self.dims), _cabi.i32x3(self.ksize), None, C.byref(nbytes), C.byref(self.plan)))
        padded = (C.c_int32 * 3)()
        _cabi.check(self.lib.mvf_rl_fft_plan_padded(self.plan, padded))
        self.padded = tuple(int(x) for x in padded)
        self.work = torch.empty(max(int(nbytes.value), 16), dtype=torch.uint8, device=device)
        _cabi.check(self.lib.mvf_rl_fft_plan_set_workspace(self.plan, C.c_void_p(self.work.data_ptr())))
        L = self.padded
        self.real = torch.empty(L, dtype=torch.float32, device=device)
        self.cplx = torch.empty((L[0], L[1], L[2] // 2 + 1, 2), dtype=torch.float32, device=device)
        self.device = device

    def spectrum(self, psf):
        """3-D plan: spectrum of the padded PSF.  zx plan: `psf` is the pair (a, B) of factor_psf_y; returns (spectrum of
        B, taps a) - what blur() takes as `spec`."""
        L = self.padded
        if self.zx:
            a, B = psf
            psf_d = torch.from_numpy(np.ascontiguousarray(B, dtype=np.float32)).to(self.device)
            spec = torch.empty((L[0], L[2] // 2 + 1, 2), dtype=torch.float32, device=self.device)
            _cabi.check(self.lib.mvf_rl_fft_psf_spectrum(self.plan, C.c_void_p(psf_d.data_ptr()), _cabi.i32x3((B.shape[0], 1, B.shape[1])),
                                                         C.c_void_p(self.real.data_ptr()), C.c_void_p(spec.data_ptr()), _stream_ptr()))
            return spec, np.ascontiguousarray(a, dtype=np.float32)
        psf_d = torch.from_numpy(np.ascontiguousarray(psf, dtype=np.float32)).to(self.device)
        spec = torch.empty((L[0], L[1], L[2] // 2 + 1, 2), dtype=torch.float32, device=self.device)
        _cabi.check(self.lib.mvf_rl_fft_psf_spectrum(self.plan, C.c_void_p(psf_d.data_ptr()), _cabi.i32x3(psf.shape),
                                                     C.c_void_p(self.real.data_ptr()), C.c_void_p(spec.data_ptr()), _stream_ptr()))
        return spec

    def blur(self, src, zmap, rp, spec, mode, view=None, weights=None, out=None, first=0, last=0, est=None, est_off=0, sums=None):
        if self.zx:
            spec2d, ky = spec
            _cabi.check(self.lib.mvf_rl_fft_blur_zx(
                self.plan, C.c_void_p(src.data_ptr()), C.c_void_p(zmap.data_ptr()), int(rp), _cabi.i32x3(self.ksize),
                C.c_void_p(ky.ctypes.data), C.c_void_p(spec2d.data_ptr()), C.c_void_p(self.real.data_ptr()),
                C.c_void_p(self.cplx.data_ptr()), int(mode),
                C.c_void_p(view.data_ptr()) if view is not None else None, dtype_code(view) if view is not None else 0,
                C.c_void_p(weights.data_ptr()) if weights is not None else None,
                C.c_void_p(out.data_ptr()) if out is not None else None, int(first), int(last),
                C.c_void_p(est.data_ptr()) if est is not None else None, int(est_off),
                C.c_void_p(sums.data_ptr()) if sums is not None else None, _stream_ptr()))
            return
        _cabi.check(self.lib.mvf_rl_fft_blur(
            self.plan, C.c_void_p(src.data_ptr()), C.c_void_p(zmap.data_ptr()), int(rp), _cabi.i32x3(self.ksize),
            C.c_void_p(spec.data_ptr()), C.c_void_p(self.real.data_ptr()), C.c_void_p(self.cplx.data_ptr()), int(mode),
            C.c_void_p(view.data_ptr()) if view is not None else None, dtype_code(view) if view is not None else 0,
            C.c_void_p(weights.data_ptr()) if weights is not None else None,
            C.c_void_p(out.data_ptr()) if out is not None else None, int(first), int(last),
            C.c_void_p(est.data_ptr()) if est is not None else None, int(est_off),
            C.c_void_p(sums.data_ptr()) if sums is not None else None, _stream_ptr()))

    def __del__(self):
        try:
            if self.plan:
                self.lib.mvf_rl_fft_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass


def _fft_spectrum(plan, dims, psf, dev):
    """Convolution state of a view that needs cuFFT: the (z, x) plan when psf = a(y) (x) B(z, x), else the 3-D plan (both
    created on first use and shared by the views of the plan).  Returns (FftBlur, what its blur() takes as `spec`)."""
    psf = np.asarray(psf)
    fy = factor_psf_y(psf)
    if fy is not None:
        if getattr(plan, "fft_zx", None) is None:
            plan.fft_zx = FftBlur(dims, psf.shape, dev, zx=True)
        return plan.fft_zx, plan.fft_zx.spectrum(fy)
    if plan.fft is None:
        plan.fft = FftBlur(dims, psf.shape, dev)
    return plan.fft, plan.fft.spectrum(psf)


def stream_mode():
    """Separable blurs run as streaming kernels (x/y passes of all planes, z passes along plane maps) unless
    MVF_RL_KERNEL=v1 selects the fused single-kernel variant (also the fallback for rows that are not 16-byte aligned)."""
    return os.environ.get("MVF_RL_KERNEL", "v4") not in ("v1", "1")


def zz_mode():
    """Default streaming schedule (v4): per view  T1 = xy(estimate);  T2 = z(ratio(z(T1)))  in ONE kernel (the ratio
    volume never reaches HBM);  correction = epilogue(xy(T2)).  MVF_RL_KERNEL=v3 keeps the ratio volume
    (xy, z + ratio, xy, z + epilogue)."""
    return os.environ.get("MVF_RL_KERNEL", "v4") in ("v4", "4")


def zz_maps_whole(nz, ksz, rp):
    """Plane maps of mvf_rl_zz for a whole volume: (zmap1 [nz + 4rp], rmap [nz + 2rp], first side plane, side planes).
    Extended positions outside the volume take their ratio from `side` = the ratio on planes [nz-2-K, nz-2]: the
    reference's boundary map sends t < 0 to nz-3-K-t (wrap) and t >= nz to 2nz-2-t (reflect)."""
    zmap1 = ext_index(np.arange(-2 * rp, nz + 2 * rp), nz, ksz).astype(np.int32)
    t = np.arange(-rp, nz + rp)
    qa = nz - 2 - ksz
    q = np.where(t < 0, nz - 3 - ksz - t, 2 * nz - 2 - t)
    k = np.clip(q - qa, 0, ksz)
    rmap = np.where((t >= 0) & (t < nz), t, -1 - k).astype(np.int32)
    return zmap1, rmap, qa, ksz + 1


def _vector_aligned(*tensors):
    """16-byte aligned float32 / 8-byte aligned uint16 buffers: what the vector loads of the streaming kernels need."""
    return all(t is None or t.data_ptr() % (16 if t.element_size() == 4 else 8) == 0 for t in tensors)


def masked_view(view, weights):
    """Copy of a view with the voxels zeroed where the weight is <= 1e-5: I = 0 makes the ratio I / (E + 1e-6) exactly 0
    there, which is all the mask of the reference does (mv:5636, 5695-5699)."""
    keep = weights > 1e-5
    if view.dtype == torch.uint16:   # torch.where has no uint16 kernel: same bits as int16
        return torch.where(keep, view.view(torch.int16), torch.zeros((), dtype=torch.int16, device=view.device)).view(torch.uint16)
    return torch.where(keep, view, torch.zeros((), dtype=view.dtype, device=view.device))


def xy_groups(psf_structs, streamed):
    """Group index per view: streamed views with identical (ky, kx) factors share the x/y passes of the estimate."""
    keys, group = [], []
    for ps, st in zip(psf_structs, streamed):
        if not st:
            group.append(-1)
            continue
        key = (tuple(ps.ky), tuple(ps.kx), int(ps.ksize[1]), int(ps.ksize[2]))
        if key not in keys:
            keys.append(key)
        group.append(keys.index(key))
    return group


def make_psf_struct(factors):
    s = _cabi.MvfRlPsf()
    for name, k in zip(("kz", "ky", "kx"), factors):
        if len(k) > _cabi.MVF_RL_KMAX or len(k) % 2 == 0:
            raise ValueError("PSF factors must have odd length <= %d" % _cabi.MVF_RL_KMAX)
        arr = getattr(s, name)
        for j, val in enumerate(k):
            arr[j] = float(val)
    s.ksize = _cabi.i32x3([len(k) for k in factors])
    return s


def padded_radius(factors):
    kp = _cabi.load().mvf_rl_padded_size(int(max(len(k) for k in factors)))
    if kp < 0:
        raise ValueError("PSF longer than %d taps" % _cabi.MVF_RL_KMAX)
    return (kp - 1) // 2


def ext_index(t, n, K):
    """Index map of the reference's FFT convolution: reflect above, wrap into the far reflected
    tail below; clamped for positions no valid output uses."""
    t = np.asarray(t)
    i = np.where(t < 0, n - 3 - K - t, np.where(t >= n, 2 * n - 2 - t, t))
    return np.clip(i, 0, n - 1)


def zmap_whole(nz, kz_size, rp):
    """Plane map of a whole (unsharded) volume: extended position t in [-rp, nz+rp) -> plane."""
    return ext_index(np.arange(-rp, nz + rp), nz, kz_size).astype(np.int32)


class RLPlan:
    """Device buffers and PSF structs of one multi-view RL problem on one GPU (whole volume)."""

    def __init__(self, views, weights, psfs):
        self.lib = _cabi.load()
        self.views, self.weights = views, weights
        self.n = len(views)
        self.dims = tuple(int(s) for s in views[0].shape)
        dev = views[0].device
        self.psf_structs, self.zmaps, self.separable, self._keep = [], [], [], []
        self.radii, self.spectra, self.fft, self.fft_zx, self.fft_of = [], [], None, None, []
        self.dense_mode = os.environ.get("MVF_RL_DENSE", "fft")  # 'fft' (cuFFT) or 'direct' (O(K^3) kernel)
        for p in psfs:
            if any(n < k + 3 for n, k in zip(self.dims, np.asarray(p).shape)):
                raise ValueError("volume extents %s are smaller than PSF size + 3 (%s): the reference's reflect padding "
                                 "needs that much" % (self.dims, np.asarray(p).shape))
            f = None if too_long_for_kernels(p) else factor_psf(p)
            self.separable.append(f is not None)
            spec, fobj = None, None
            if f is not None:  # outer-product PSF (axis-aligned view): fused separable kernel
                self.psf_structs.append(make_psf_struct(f))
                ksz, rp = len(f[0]), padded_radius(f)
            elif too_long_for_kernels(p) and self.dense_mode == "fft":   # more taps than the convolution kernels hold
                if any(k % 2 == 0 for k in np.asarray(p).shape):
                    raise ValueError("PSF sizes must be odd")
                self.psf_structs.append(fft_only_psf_struct(p))
                ksz, rp = int(np.asarray(p).shape[0]), (int(np.asarray(p).shape[0]) - 1) // 2
                fobj, spec = _fft_spectrum(self, self.dims, p, dev)
            else:              # oblique view: dense PSF -> cuFFT convolution (or the direct kernel on request)
                st, keep = make_dense_psf_struct(p, dev)
                self.psf_structs.append(st)
                self._keep.append(keep)
                ksz, rp = int(np.asarray(p).shape[0]), (int(keep.shape[0]) - 1) // 2
                if self.dense_mode == "fft":
                    fobj, spec = _fft_spectrum(self, self.dims, p, dev)
            self.spectra.append(spec)
            self.fft_of.append(fobj if spec is not None else None)
            self.radii.append(rp)
            self.zmaps.append(torch.from_numpy(zmap_whole(self.dims[0], ksz, rp)).to(dev))
        self.est = torch.empty(self.dims, dtype=torch.float32, device=dev)
        self.ratio = torch.empty(self.dims, dtype=torch.float32, device=dev)
        self.corr = torch.empty(self.dims, dtype=torch.float32, device=dev)
        self.sums = torch.zeros(2, dtype=torch.float64, device=dev)
        self._vp = (C.c_void_p * self.n)(*[v.data_ptr() for v in views])
        self._wp = (C.c_void_p * self.n)(*[w.data_ptr() for w in weights])
        self._dims = _cabi.i32x3(self.dims)
        self.streamed = [sep and stream_mode() and bool(self.lib.mvf_rl_stream_supported(self._dims, C.byref(ps))) and
                         _vector_aligned(vw, ww, self.est, self.ratio, self.corr)
                         for sep, ps, vw, ww in zip(self.separable, self.psf_structs, views, weights)]
        self.tmp = torch.empty(self.dims, dtype=torch.float32, device=dev) if any(self.streamed) else None
        # v4 schedule for every streamed view; views that need cuFFT (or the fused fallback kernel) keep a ratio volume
        self.zz = zz_mode() and any(self.streamed)
        if self.zz:   # plane maps of the fused z-ratio-z kernel, side planes for the positions outside the volume
            self.zz_maps = []
            for ps, rp, st in zip(self.psf_structs, self.radii, self.streamed):
                if not st:
                    self.zz_maps.append(None)
                    continue
                zmap1, rmap, qa, nside = zz_maps_whole(self.dims[0], int(ps.ksize[0]), rp)
                self.zz_maps.append((np.ascontiguousarray(zmap1), np.ascontiguousarray(rmap), qa, nside))   # host arrays
            self.side = torch.empty((max(m[3] for m in self.zz_maps if m is not None),) + self.dims[1:], dtype=torch.float32, device=dev)
            if all(self.streamed):
                self.ratio = None   # never materialised
        # streamed views: (1) the mask (w > 1e-5) of the ratio step is burnt into a copy of the view once, so the ratio
        # pass does not read the weights; (2) views whose PSFs share the in-plane factors (ky, kx) share one x/y pass
        # of the estimate per iteration
        self.masked = [masked_view(vw, ww) if st else None for st, vw, ww in zip(self.streamed, views, weights)]
        self.xy_group, self.xy_tmps = xy_groups(self.psf_structs, self.streamed), []
        for _ in range(max(self.xy_group) + 1 if any(self.streamed) else 0):
            self.xy_tmps.append(torch.empty(self.dims, dtype=torch.float32, device=dev))

    def init_estimate(self):
        _init_estimate(self)

    def iterate(self):
        """One RL iteration; sum(estimate) lands in self.sums[1].  After the first (eager) call the launches of an
        iteration are replayed from a CUDA graph (MVF_RL_GRAPH=0: always eager)."""
        _graphed(self, self._iterate_eager)

    def _iterate_eager(self):
        st = _stream_ptr()
        self.sums[1:].zero_()
        done = set()
        for v in range(self.n):   # x/y passes of the estimate, once per distinct (ky, kx) pair (the estimate changes only at the end)
            g = self.xy_group[v]
            if self.streamed[v] and g not in done:
                done.add(g)
                _cabi.check(self.lib.mvf_rl_xy(C.c_void_p(self.est.data_ptr()), self.dims[0], self._dims, C.byref(self.psf_structs[v]),
                                               C.c_void_p(self.xy_tmps[g].data_ptr()), st))
        for v in range(self.n):
            last = v == self.n - 1
            if self.spectra[v] is not None:  # cuFFT path
                self.fft_of[v].blur(self.est, self.zmaps[v], self.radii[v], self.spectra[v], 1, view=self.views[v],
                              weights=self.weights[v], out=self.ratio)
                self.fft_of[v].blur(self.ratio, self.zmaps[v], self.radii[v], self.spectra[v], 2, weights=self.weights[v],
                              out=self.corr, first=1 if v == 0 else 0, last=1 if last else 0, est=self.est,
                              sums=self.sums[1:])
                continue
            zm = C.c_void_p(self.zmaps[v].data_ptr())
            if self.zz and self.streamed[v]:   # side planes, then z + ratio + z in one kernel, then x/y passes + correction epilogue
                ps = C.byref(self.psf_structs[v])
                t1 = self.xy_tmps[self.xy_group[v]]
                zmap1, rmap, qa, nside = self.zz_maps[v]
                plane = self.dims[1] * self.dims[2]
                _cabi.check(self.lib.mvf_rl_ratio_z(C.c_void_p(t1.data_ptr()), C.c_void_p(self.zmaps[v].data_ptr() + 4 * qa),
                                                    _cabi.i32x3((nside,) + self.dims[1:]), ps,
                                                    C.c_void_p(self.masked[v].data_ptr() + qa * plane * self.masked[v].element_size()),
                                                    dtype_code(self.views[v]), None, C.c_void_p(self.side.data_ptr()), st))
                _cabi.check(self.lib.mvf_rl_zz(C.c_void_p(t1.data_ptr()), C.c_void_p(zmap1.ctypes.data), C.c_void_p(rmap.ctypes.data),
                                               self._dims, ps, C.c_void_p(self.masked[v].data_ptr()), dtype_code(self.views[v]),
                                               C.c_void_p(self.side.data_ptr()), C.c_void_p(self.tmp.data_ptr()), st))
                _cabi.check(self.lib.mvf_rl_xy_correct(C.c_void_p(self.tmp.data_ptr()), self.dims[0], self._dims, ps,
                                                       C.c_void_p(self.weights[v].data_ptr()), C.c_void_p(self.corr.data_ptr()),
                                                       1 if v == 0 else 0, 1 if last else 0, C.c_void_p(self.est.data_ptr()), 0,
                                                       C.c_void_p(self.sums[1:].data_ptr()), st))
                continue
            if self.streamed[v]:  # x/y passes of every plane into tmp, then z pass + epilogue
                ps, tmp = C.byref(self.psf_structs[v]), C.c_void_p(self.tmp.data_ptr())
                _cabi.check(self.lib.mvf_rl_ratio_z(C.c_void_p(self.xy_tmps[self.xy_group[v]].data_ptr()), zm, self._dims, ps,
                                                    C.c_void_p(self.masked[v].data_ptr()), dtype_code(self.views[v]), None,
                                                    C.c_void_p(self.ratio.data_ptr()), st))
                _cabi.check(self.lib.mvf_rl_xy(C.c_void_p(self.ratio.data_ptr()), self.dims[0], self._dims, ps, tmp, st))
                _cabi.check(self.lib.mvf_rl_correct_z(tmp, zm, self._dims, ps, C.c_void_p(self.weights[v].data_ptr()),
                                                      C.c_void_p(self.corr.data_ptr()), 1 if v == 0 else 0, 1 if last else 0,
                                                      C.c_void_p(self.est.data_ptr()), 0, C.c_void_p(self.sums[1:].data_ptr()), st))
                continue
            _cabi.check(self.lib.mvf_rl_ratio(C.c_void_p(self.est.data_ptr()), zm, self._dims, C.byref(self.psf_structs[v]),
                                              C.c_void_p(self.views[v].data_ptr()), dtype_code(self.views[v]),
                                              C.c_void_p(self.weights[v].data_ptr()), C.c_void_p(self.ratio.data_ptr()), st))
            _cabi.check(self.lib.mvf_rl_correct(C.c_void_p(self.ratio.data_ptr()), zm, self._dims, C.byref(self.psf_structs[v]),
                                                C.c_void_p(self.weights[v].data_ptr()), C.c_void_p(self.corr.data_ptr()),
                                                1 if v == 0 else 0, 1 if last else 0, C.c_void_p(self.est.data_ptr()), 0,
                                                C.c_void_p(self.sums[1:].data_ptr()), st))

    def run(self, num_iterations=25, tol=5e-5):
        """The reference's loop (mv:5678-5734): stops after num_iterations, or once i >= 10 and the relative
        change of sum(estimate) drops below tol."""
        self.init_estimate()
        curr = None
        i = 0
        while True:
            self.iterate()
            if i >= 10:  # the convergence test can only fire from the 11th iteration on: no sync before that
                s = self.sums.cpu().numpy()
                curr = s[0] if curr is None else curr
                new = s[1]
                if abs(1 - new / curr) < tol:
                    break
                curr = new
            else:
                self.sums[0] = self.sums[1]
            if i >= num_iterations - 1:
                break
            i += 1
        self.iterations_run = i + 1
        return self.est


def _init_estimate(plan):
    """estimate = sum_v w_v * I_v and its sum (mv:5640-5645); more than MVF_MAX_VIEWS views go through the kernel in
    groups whose partial estimates are added."""
    plan.sums.zero_()
    code, est = dtype_code(plan.views[0]), plan.est
    cap = _cabi.MVF_MAX_VIEWS
    if plan.n <= cap:
        _cabi.check(plan.lib.mvf_rl_init_estimate(plan.n, plan._vp, code, plan._wp, C.c_void_p(est.data_ptr()), est.numel(),
                                                  C.c_void_p(plan.sums.data_ptr()), _stream_ptr()))
        return
    part = torch.empty(tuple(est.shape), dtype=torch.float32, device=est.device)
    for i0 in range(0, plan.n, cap):
        n = min(cap, plan.n - i0)
        vp = (C.c_void_p * n)(*[v.data_ptr() for v in plan.views[i0:i0 + n]])
        wp = (C.c_void_p * n)(*[w.data_ptr() for w in plan.weights[i0:i0 + n]])
        dst = est if i0 == 0 else part
        if i0 == 0 and not est.is_contiguous():
            raise ValueError("estimate buffer must be contiguous")
        _cabi.check(plan.lib.mvf_rl_init_estimate(n, vp, code, wp, C.c_void_p(dst.data_ptr()), dst.numel(), None, _stream_ptr()))
        if i0:
            est += part
    plan.sums[0] = est.sum(dtype=torch.float64)


def _graphed(plan, eager):
    """Run `eager()` once, capture it into a CUDA graph on the second call, replay from then on.  Anything that cannot be
    captured (old driver, an op that syncs) leaves the plan in eager mode."""
    state = getattr(plan, "_graph_state", 0)
    if state == 0 or os.environ.get("MVF_RL_GRAPH", "1") == "0" or not torch.cuda.is_available():
        plan._graph_state = 1 if state == 0 else state
        eager()
        return
    if state == 1:
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                eager()
            plan._graph, plan._graph_state = graph, 2
        except Exception:   # not capturable here: stay eager
            plan._graph_state = 3
            cur.wait_stream(side)
            torch.cuda.synchronize()
            eager()
            return
        cur.wait_stream(side)
    if plan._graph_state == 2:
        plan._graph.replay()
    else:
        eager()


def blur(volume, psf_factors):
    """max(blur(volume), 0) of a float32 CUDA tensor with a separable PSF (tests, PSF checks)."""
    lib = _cabi.load()
    dims = tuple(int(s) for s in volume.shape)
    zm = torch.from_numpy(zmap_whole(dims[0], len(psf_factors[0]), padded_radius(psf_factors))).to(volume.device)
    out = torch.empty_like(volume)
    ps = make_psf_struct(psf_factors)
    if stream_mode() and lib.mvf_rl_stream_supported(_cabi.i32x3(dims), C.byref(ps)) and _vector_aligned(volume, out):
        tmp = torch.empty_like(volume)
        _cabi.check(lib.mvf_rl_xy(C.c_void_p(volume.data_ptr()), dims[0], _cabi.i32x3(dims), C.byref(ps),
                                  C.c_void_p(tmp.data_ptr()), _stream_ptr()))
        _cabi.check(lib.mvf_rl_blur_z(C.c_void_p(tmp.data_ptr()), C.c_void_p(zm.data_ptr()), _cabi.i32x3(dims), C.byref(ps),
                                      C.c_void_p(out.data_ptr()), _stream_ptr()))
        return out
    _cabi.check(lib.mvf_rl_blur(C.c_void_p(volume.data_ptr()), C.c_void_p(zm.data_ptr()), _cabi.i32x3(dims), C.byref(ps),
                                C.c_void_p(out.data_ptr()), _stream_ptr()))
    return out


class SlabRLPlan:
    """Multi-view RL on one z-slab of the fused frame (one rank of a box): halo-padded estimate / ratio buffers,
    ring halo exchange over torch.distributed before every blur, scalar all-reduce for the convergence test."""

    def __init__(self, views, weights, psfs, nz_total, rank, world, group=None, bounds=None):
        """`bounds`: the z range of every rank (default: slab.slab_bounds, equal slabs; slab.rl_slab_bounds balances the
        v4 schedule by work)."""
        from . import slab
        self.bounds = bounds
        self.lib = _cabi.load()
        self.views, self.weights, self.group = views, weights, group
        self.n = len(views)
        dev = views[0].device
        nzl, ny, nx = (int(s) for s in views[0].shape)
        shapes = {tuple(np.asarray(p).shape) for p in psfs}
        if len(shapes) != 1:
            raise ValueError("all views must share one PSF shape")
        kz = next(iter(shapes))[0]
        self.psf_structs, self._keep, self.spectra = [], [], []
        self.fft, self.fft_zx, self.fft_of = None, None, []
        dense_mode = os.environ.get("MVF_RL_DENSE", "fft")
        rp = None
        facs = []
        for p in psfs:
            f = None if too_long_for_kernels(p) else factor_psf(p)
            facs.append(f)
            spec, fobj = None, None
            if f is not None:
                self.psf_structs.append(make_psf_struct(f))
                r = padded_radius(f)
            elif too_long_for_kernels(p) and dense_mode == "fft":
                self.psf_structs.append(fft_only_psf_struct(p))
                r = (int(np.asarray(p).shape[0]) - 1) // 2
                fobj, spec = _fft_spectrum(self, (nzl, ny, nx), p, dev)
            else:
                st, keep = make_dense_psf_struct(p, dev)
                self.psf_structs.append(st)
                self._keep.append(keep)
                r = (int(keep.shape[0]) - 1) // 2
                if dense_mode == "fft":
                    fobj, spec = _fft_spectrum(self, (nzl, ny, nx), p, dev)
            self.spectra.append(spec)
            self.fft_of.append(fobj if spec is not None else None)
            rp = r if rp is None else rp
            if r != rp:
                raise ValueError("PSFs need one common padded size")
        self.zz = False
        if zz_mode() and all(f is not None for f in facs) and self._init_zz(views, weights, nz_total, rank, world, kz, rp, group):
            return
        self.halo = slab.HaloPlan(nz_total, world, rank, kz, rp, self.bounds)
        if self.halo.nzl != nzl:
            raise ValueError("slab height %d does not match the planner's %d" % (nzl, self.halo.nzl))
        self.rp, self.dims = rp, (nzl, ny, nx)
        self.zmap = torch.from_numpy(self.halo.zmap()).to(dev)
        # Halo transport.  "symm" (default on CUDA with more than one rank): the padded volumes live in symmetric
        # memory and every rank pulls its halo planes from the neighbours itself (slab.SymmetricHalo); "nccl": ring of
        # isend/irecv (also what the gloo tests on CPU exercise).
        self.symm = None
        want = os.environ.get("MVF_HALO", "symm")
        if world > 1 and want == "symm" and dev.type == "cuda" and self.fft is None and self.fft_zx is None:
            try:
                self.symm = slab.SymmetricHalo(3, self.halo, ny, nx, dev, group)
            except Exception as e:   # no peer access / no symmetric memory support: the NCCL ring still works
                if os.environ.get("MVF_HALO") == "symm":
                    raise
                self.symm, self._symm_error = None, repr(e)
        if self.symm is not None:
            self.est_pad, self.ratio_pads = self.symm.bufs[0], [self.symm.bufs[1], self.symm.bufs[2]]
            self.comm_stream = torch.cuda.Stream()
            self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()
        else:
            self.est_pad = torch.zeros((nzl + 2 * rp, ny, nx), dtype=torch.float32, device=dev)
            self.ratio_pads = [torch.zeros((nzl + 2 * rp, ny, nx), dtype=torch.float32, device=dev)]
        self.ratio_pad = self.ratio_pads[0]
        self.est = self.est_pad[rp:rp + nzl]
        self.ratio = self.ratio_pad[rp:rp + nzl]
        self.corr = torch.empty(self.dims, dtype=torch.float32, device=dev)
        self.sums = torch.zeros(2, dtype=torch.float64, device=dev)
        self._vp = (C.c_void_p * self.n)(*[v.data_ptr() for v in views])
        self._wp = (C.c_void_p * self.n)(*[w.data_ptr() for w in weights])
        self._dims = _cabi.i32x3(self.dims)
        self._est_off = rp * ny * nx
        self.streamed = [sp is None and not bool(ps.dense) and stream_mode() and
                         bool(self.lib.mvf_rl_stream_supported(self._dims, C.byref(ps))) and
                         _vector_aligned(vw, ww, self.est, self.ratio, self.corr)
                         for sp, ps, vw, ww in zip(self.spectra, self.psf_structs, views, weights)]
        self.tmp_pad = torch.empty_like(self.est_pad) if any(self.streamed) else None
        # as in RLPlan: mask burnt into a copy of the streamed views, one x/y pass of the estimate per (ky, kx) pair
        self.masked = [masked_view(vw, ww) if st else None for st, vw, ww in zip(self.streamed, views, weights)]
        self.xy_group, self.xy_tmps = xy_groups(self.psf_structs, self.streamed), []
        for _ in range(max(self.xy_group) + 1 if any(self.streamed) else 0):
            self.xy_tmps.append(torch.empty_like(self.est_pad))

    # ---- v4 schedule on a slab (slab.ZZLayout): one estimate exchange per iteration, no ratio volume ---------------
    def _init_zz(self, views, weights, nz_total, rank, world, kz, rp, group):
        """Buffers and plane maps of the v4 schedule; False (nothing allocated) when it does not apply: rows that the
        vector kernels cannot take, or slabs too thin for 2*rp halo planes - the v3 schedule handles those."""
        from . import slab
        import torch.distributed as dist
        dev = views[0].device
        nzl, ny, nx = (int(s) for s in views[0].shape)
        dims = _cabi.i32x3((nzl, ny, nx))
        if dev.type != "cuda" or not all(bool(self.lib.mvf_rl_stream_supported(dims, C.byref(ps))) for ps in self.psf_structs):
            return False
        if not _vector_aligned(*weights):
            return False
        try:
            lay = slab.ZZLayout(nz_total, world, rank, kz, rp, self.bounds)
        except ValueError:
            return False
        if lay.nzl != nzl:
            raise ValueError("slab height %d does not match the planner's %d" % (nzl, lay.nzl))
        self.zz, self.layout, self.rp, self.dims = True, lay, rp, (nzl, ny, nx)
        self.halo = lay            # (.world / .rank / .nzl as for the v3 planner)
        self.spectra = [None] * self.n
        self.streamed = [True] * self.n
        self.symm = None
        want = os.environ.get("MVF_HALO", "symm")
        if world > 1 and want == "symm" and dist.is_available() and dist.is_initialized():
            try:
                self.symm = slab.SymmetricPull(lay, ny, nx, dev, group)
            except Exception as e:   # no peer access / no symmetric memory support: point-to-point transfers still work
                if os.environ.get("MVF_HALO") == "symm":
                    raise
                self.symm, self._symm_error = None, repr(e)
        if self.symm is not None:
            self.est_pad = self.symm.buf
            self.comm_stream = torch.cuda.Stream()
            self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()
        else:
            self.est_pad = torch.zeros((lay.height, ny, nx), dtype=torch.float32, device=dev)
        self.est = self.est_pad[lay.front:lay.front + nzl]
        self.corr = torch.empty(self.dims, dtype=torch.float32, device=dev)
        self.t2 = torch.empty(self.dims, dtype=torch.float32, device=dev)
        self.side = torch.zeros((kz + 1, ny, nx), dtype=torch.float32, device=dev)
        self.sums = torch.zeros(2, dtype=torch.float64, device=dev)
        self._vp = (C.c_void_p * self.n)(*[v.data_ptr() for v in views])
        self._wp = (C.c_void_p * self.n)(*[w.data_ptr() for w in weights])
        self._dims = dims
        self._est_off = lay.front * ny * nx
        # masked views with rp static halo planes (filled from the neighbours once)
        self.masked_pad = []
        for vw, ww in zip(views, weights):
            if vw.dtype == torch.uint16:   # (fill / slice copies via the int16 view: same bits, full operator support)
                mp = torch.zeros((lay.vheight, ny, nx), dtype=torch.int16, device=dev)
                mp[rp:rp + nzl] = masked_view(vw, ww).view(torch.int16)
                mp = mp.view(torch.uint16)
            else:
                mp = torch.zeros((lay.vheight, ny, nx), dtype=vw.dtype, device=dev)
                mp[rp:rp + nzl] = masked_view(vw, ww)
            self.masked_pad.append(mp)
        if world > 1 and dist.is_available() and dist.is_initialized():
            for mp in self.masked_pad:
                # (as bytes: NCCL moves neither uint16 nor int16 tensors)
                slab.run_transfers_dist(mp.view(torch.uint8) if mp.dtype == torch.uint16 else mp, lay, "view_transfers", group)
        self.xy_group = xy_groups(self.psf_structs, self.streamed)
        self.xy_tmps = [torch.zeros((lay.height, ny, nx), dtype=torch.float32, device=dev) for _ in range(max(self.xy_group) + 1)]
        self.zmap1 = np.ascontiguousarray(lay.zmap1())   # host arrays: they travel as kernel parameters
        self.rmap = np.ascontiguousarray(lay.rmap())
        self.side_jobs = [(torch.from_numpy(zm).to(dev), v0, cnt, k0) for zm, v0, cnt, k0 in lay.side_jobs()]
        return True

    def _zz_xy_estimate(self, join=None):
        """x/y passes of the padded estimate, once per group of views with equal in-plane PSF factors.  `join` (halo
        copies in flight on the side stream): the own planes of every group go first, the halo planes once they have
        arrived.  Ranks other than 0 use only the upper 2rp planes of the front block."""
        lay, st = self.layout, _stream_ptr()
        pstride = self.dims[1] * self.dims[2] * 4
        lo = 0 if (lay.rank == 0 or lay.world == 1) else lay.front - lay.back
        nback = lay.back if lay.rank + 1 < lay.world else 0      # (the last rank reflects into its own planes)
        groups, seen = [], set()
        for v in range(self.n):
            if self.xy_group[v] not in seen:
                seen.add(self.xy_group[v])
                groups.append((self.xy_group[v], C.byref(self.psf_structs[v])))

        def run(g, ps, p0, n):
            if n > 0:
                _cabi.check(self.lib.mvf_rl_xy(C.c_void_p(self.est_pad.data_ptr() + p0 * pstride), int(n), self._dims, ps,
                                               C.c_void_p(self.xy_tmps[g].data_ptr() + p0 * pstride), st))
        if join is None:
            for g, ps in groups:
                run(g, ps, lo, lay.front + lay.nzl + nback - lo)
            return
        # only the first two groups are split (their own planes cover the halo copies); the others run as ONE launch over the whole
        # contiguous plane range afterwards - launches of 18 halo planes fill half the GPU (measured at N = 8: 3.9 us per
        # plane and view against 2.9 us in the 64-plane launches)
        split = groups[:2]   # (two groups of own planes outlast the pull of 5 * rp planes on rank 0)
        for g, ps in split:
            run(g, ps, lay.front, lay.nzl)
        join()
        for g, ps in split:
            run(g, ps, lo, lay.front - lo)
            run(g, ps, lay.front + lay.nzl, nback)
        for g, ps in groups[2:]:
            run(g, ps, lo, lay.front + lay.nzl + nback - lo)

    def _zz_view(self, v, before_update=None):
        """Boundary ratios, then z + ratio + z in one kernel, then x/y passes + correction epilogue of view v
        (`before_update` runs ahead of the launch that overwrites the estimate)."""
        st, ps = _stream_ptr(), C.byref(self.psf_structs[v])
        nzl, ny, nx = self.dims
        plane = ny * nx
        t1, mp = self.xy_tmps[self.xy_group[v]], self.masked_pad[v]
        code = dtype_code(self.views[v])
        for zm, v0, cnt, k0 in self.side_jobs:
            _cabi.check(self.lib.mvf_rl_ratio_z(C.c_void_p(t1.data_ptr()), C.c_void_p(zm.data_ptr()), _cabi.i32x3((cnt, ny, nx)), ps,
                                                C.c_void_p(mp.data_ptr() + v0 * plane * mp.element_size()), code, None,
                                                C.c_void_p(self.side.data_ptr() + k0 * plane * 4), st))
        _cabi.check(self.lib.mvf_rl_zz(C.c_void_p(t1.data_ptr()), C.c_void_p(self.zmap1.ctypes.data), C.c_void_p(self.rmap.ctypes.data),
                                       self._dims, ps, C.c_void_p(mp.data_ptr()), code, C.c_void_p(self.side.data_ptr()),
                                       C.c_void_p(self.t2.data_ptr()), st))
        last = v == self.n - 1
        if last and before_update is not None:
            before_update()
        _cabi.check(self.lib.mvf_rl_xy_correct(C.c_void_p(self.t2.data_ptr()), nzl, self._dims, ps,
                                               C.c_void_p(self.weights[v].data_ptr()), C.c_void_p(self.corr.data_ptr()),
                                               1 if v == 0 else 0, 1 if last else 0, C.c_void_p(self.est_pad.data_ptr()),
                                               self._est_off, C.c_void_p(self.sums[1:].data_ptr()), st))

    def iterate_compute(self, join=None, before_update=None):
        """Everything of a v4 iteration after the estimate halos have been requested / filled."""
        self.sums[1:].zero_()
        self._zz_xy_estimate(join)
        for v in range(self.n):
            self._zz_view(v, before_update)

    def _iterate_zz_symm(self):
        main = torch.cuda.current_stream()
        self.symm.barrier()
        self._ev_fork.record(main)
        self.comm_stream.wait_event(self._ev_fork)
        with torch.cuda.stream(self.comm_stream):
            self.symm.pull()
            self._ev_join.record(self.comm_stream)
        # second barrier: no rank overwrites its estimate (last launch of the iteration) before every peer has pulled the
        # old planes; the first one orders the pulls after the previous update
        self.iterate_compute(lambda: main.wait_event(self._ev_join), self.symm.barrier)

    def init_estimate(self):
        _init_estimate(self)

    def _xy_planes(self, src_pad, v, pending=(), dst=None):
        """x/y passes of src_pad into dst (default tmp_pad).  With a halo exchange in flight (`pending`) the own planes
        go first and the halo planes follow once they have arrived, so the transfer runs beside the bulk of the work."""
        from . import slab
        ps, st = C.byref(self.psf_structs[v]), _stream_ptr()
        nzl, rp = self.dims[0], self.rp
        pstride = self.dims[1] * self.dims[2] * 4
        dst = self.tmp_pad if dst is None else dst

        def run(p0, n):
            _cabi.check(self.lib.mvf_rl_xy(C.c_void_p(src_pad.data_ptr() + p0 * pstride), int(n), self._dims, ps,
                                           C.c_void_p(dst.data_ptr() + p0 * pstride), st))
        if callable(pending):   # symmetric-memory pull in flight on the side stream: `pending()` joins it
            run(rp, nzl)
            pending()
            run(0, rp)
            run(rp + nzl, rp)
            return
        if not pending or rp == 0:
            slab.exchange_halos_end(pending)
            run(0, nzl + 2 * rp)
            return
        run(rp, nzl)
        slab.exchange_halos_end(pending)
        run(0, rp)
        run(rp + nzl, rp)

    def step_ratio(self, v, pending=()):
        """R_v = I_v / (blur(est) + 1e-6) * mask, read from the halo-padded estimate (`pending`: requests of an
        estimate halo exchange still in flight)."""
        from . import slab
        if self.spectra[v] is not None:
            slab.exchange_halos_end(pending)
            self.fft_of[v].blur(self.est_pad, self.zmap, self.rp, self.spectra[v], 1, view=self.views[v],
                          weights=self.weights[v], out=self.ratio)
            return
        if self.streamed[v]:   # the x/y passes of the estimate were done once per (ky, kx) group (xy_estimate)
            slab.exchange_halos_end(pending)
            _cabi.check(self.lib.mvf_rl_ratio_z(C.c_void_p(self.xy_tmps[self.xy_group[v]].data_ptr()),
                                                C.c_void_p(self.zmap.data_ptr()), self._dims, C.byref(self.psf_structs[v]),
                                                C.c_void_p(self.masked[v].data_ptr()), dtype_code(self.views[v]), None,
                                                C.c_void_p(self.ratio.data_ptr()), _stream_ptr()))
            return
        slab.exchange_halos_end(pending)
        _cabi.check(self.lib.mvf_rl_ratio(C.c_void_p(self.est_pad.data_ptr()), C.c_void_p(self.zmap.data_ptr()), self._dims,
                                          C.byref(self.psf_structs[v]), C.c_void_p(self.views[v].data_ptr()),
                                          dtype_code(self.views[v]), C.c_void_p(self.weights[v].data_ptr()),
                                          C.c_void_p(self.ratio.data_ptr()), _stream_ptr()))

    def xy_estimate(self, pending=()):
        """x/y passes of the halo-padded estimate, once per group of streamed views with equal in-plane PSF factors;
        returns the requests that are still pending (none once a group has consumed the halos)."""
        done = set()
        for v in range(self.n):
            g = self.xy_group[v]
            if self.streamed[v] and g not in done:
                done.add(g)
                self._xy_planes(self.est_pad, v, pending, dst=self.xy_tmps[g])
                pending = ()
        return pending

    def step_correct(self, v, pending=()):
        """C (+)= blur(R_v) * w_v; the last view applies the multiplicative update to the estimate (`pending`:
        requests of a ratio halo exchange still in flight)."""
        from . import slab
        last = v == self.n - 1
        if self.spectra[v] is not None:
            slab.exchange_halos_end(pending)
            self.fft_of[v].blur(self.ratio_pad, self.zmap, self.rp, self.spectra[v], 2, weights=self.weights[v], out=self.corr,
                          first=1 if v == 0 else 0, last=1 if last else 0, est=self.est_pad, est_off=self._est_off,
                          sums=self.sums[1:])
            return
        if self.streamed[v]:
            ps, tmp = C.byref(self.psf_structs[v]), C.c_void_p(self.tmp_pad.data_ptr())
            self._xy_planes(self.ratio_pad, v, pending)
            _cabi.check(self.lib.mvf_rl_correct_z(tmp, C.c_void_p(self.zmap.data_ptr()), self._dims, ps,
                                                  C.c_void_p(self.weights[v].data_ptr()), C.c_void_p(self.corr.data_ptr()),
                                                  1 if v == 0 else 0, 1 if last else 0, C.c_void_p(self.est_pad.data_ptr()),
                                                  self._est_off, C.c_void_p(self.sums[1:].data_ptr()), _stream_ptr()))
            return
        slab.exchange_halos_end(pending)
        _cabi.check(self.lib.mvf_rl_correct(C.c_void_p(self.ratio_pad.data_ptr()), C.c_void_p(self.zmap.data_ptr()), self._dims,
                                            C.byref(self.psf_structs[v]), C.c_void_p(self.weights[v].data_ptr()),
                                            C.c_void_p(self.corr.data_ptr()), 1 if v == 0 else 0, 1 if last else 0,
                                            C.c_void_p(self.est_pad.data_ptr()), self._est_off,
                                            C.c_void_p(self.sums[1:].data_ptr()), _stream_ptr()))

    def _select_ratio(self, i):
        """Ratio volumes alternate between two symmetric buffers: a rank may start the next view's ratio while a
        neighbour is still pulling the halo planes of the previous one."""
        self.ratio_pad = self.ratio_pads[i % len(self.ratio_pads)]
        self.ratio = self.ratio_pad[self.rp:self.rp + self.dims[0]]

    def _pull_async(self, buf_index):
        """Barrier on the compute stream (every rank has written its own planes), then the two halo copies on the side
        stream; returns the join that the first consumer of the halo planes calls."""
        main = torch.cuda.current_stream()
        self.symm.barrier()
        self._ev_fork.record(main)
        self.comm_stream.wait_event(self._ev_fork)
        with torch.cuda.stream(self.comm_stream):
            self.symm.pull(buf_index)
            self._ev_join.record(self.comm_stream)
        return lambda: main.wait_event(self._ev_join)

    def _iterate_symm(self):
        self.sums[1:].zero_()
        join = self._pull_async(0)
        if any(self.streamed):
            join = self.xy_estimate(join)
        if callable(join):
            join()
        for v in range(self.n):
            self._select_ratio(v)
            self.step_ratio(v)
            join = self._pull_async(1 + v % 2)
            if self.streamed[v]:
                self.step_correct(v, join)
            else:
                join()
                self.step_correct(v)

    def iterate(self):
        from . import slab
        if self.zz:
            if self.symm is not None:
                _graphed(self, self._iterate_zz_symm)
            else:
                slab.run_transfers_dist(self.est_pad, self.layout, "est_transfers", self.group)
                self.iterate_compute()
            return
        if self.symm is not None:
            _graphed(self, self._iterate_symm)
            return
        self.sums[1:].zero_()
        # one estimate exchange serves all views' first blur; every exchange runs beside the x/y passes of the own planes
        pending = slab.exchange_halos_begin(self.est_pad, self.halo, self.group)
        pending = self.xy_estimate(pending)
        for v in range(self.n):
            self.step_ratio(v, pending)
            pending = slab.exchange_halos_begin(self.ratio_pad, self.halo, self.group)
            self.step_correct(v, pending)
            pending = ()

    def run(self, num_iterations=25, tol=5e-5):
        import torch.distributed as dist
        self.init_estimate()
        curr = None
        i = 0
        while True:
            self.iterate()
            if i >= 10:
                s = self.sums.clone()
                if self.halo.world > 1:
                    dist.all_reduce(s, group=self.group)
                s = s.cpu().numpy()
                curr = s[0] if curr is None else curr
                if abs(1 - s[1] / curr) < tol:
                    break
                curr = s[1]
            else:
                self.sums[0] = self.sums[1]
            if i >= num_iterations - 1:
                break
            i += 1
        self.iterations_run = i + 1
        return self.est


def setup_emulated(plans):
    """Ranks emulated in ONE process (tests on a single GPU): fill the static view halos of the v4 schedule."""
    from . import slab
    if plans[0].zz:
        for v in range(plans[0].n):
            bufs = [p.masked_pad[v].view(torch.int16) if p.masked_pad[v].dtype == torch.uint16 else p.masked_pad[v] for p in plans]
            slab.run_transfers_local(bufs, [p.layout for p in plans], "view_transfers")


def iterate_emulated(plans):
    """One RL iteration of all emulated ranks with the same plane movement as the NVLink / NCCL exchanges."""
    from . import slab
    if plans[0].zz:
        slab.run_transfers_local([p.est_pad for p in plans], [p.layout for p in plans], "est_transfers")
        for p in plans:
            p.iterate_compute()
        return
    for p in plans:
        p.sums[1:].zero_()
    slab.exchange_halos_local([p.est_pad for p in plans], [p.halo for p in plans])
    for p in plans:
        p.xy_estimate()
    for v in range(plans[0].n):
        for p in plans:
            p.step_ratio(v)
        slab.exchange_halos_local([p.ratio_pad for p in plans], [p.halo for p in plans])
        for p in plans:
            p.step_correct(v)
