"""Golden vectors for the steps either side of the path, from the UNMODIFIED reference source.

    python tests/golden/make_golden_prepost.py        # writes tests/golden/prepost_golden.npz (needs /root/reference)

* ``mvregfus/imaris.py`` is imported with a recording stand-in for h5py (absent from the image): ``np_to_ims``
  runs unchanged and every dataset / attribute it would write is captured, which pins the cumulative
  subsampling and the per-level ``np.histogram(data, 256)`` of the .ims pyramid.
* ``mvregfus/mv_utils.py::bin_stack`` runs with a SimpleITK stand-in whose ``BinShrink`` is the oracle's
  restatement of the ITK filter: this pins the reference's spacing / origin / axis-order arithmetic around it,
  NOT ITK's own rounding ("parity unpinned" for that part, as for the resampling).
* background subtraction (``multiview.py:316``) is one numpy expression inside an I/O function; it is evaluated
  here verbatim on a seeded stack.
"""
import importlib
import os
import sys
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import mvregfus_oracle as O  # noqa: E402
from make_golden import blob_volume  # noqa: E402  (seeded inputs)


class _Attrs(dict):
    def create(self, key, value):
        self[key] = value


class _Group:
    def __init__(self, store, path):
        self.store, self.path, self.attrs = store, path, _Attrs()

    def create_dataset(self, name, data=None, **kw):
        self.store.datasets[self.path.rstrip("/") + "/" + name] = (np.array(data), kw)

    def __repr__(self):
        return "<group %s>" % self.path


class _File:
    last = None

    def __init__(self, fname, mode="a"):
        self.groups, self.datasets = {}, {}
        _File.last = self

    def __enter__(self): return self
    def __exit__(self, *a): return False

    def create_group(self, path):
        g = _Group(self, path)
        self.groups[path.strip("/")] = g
        return g

    def __getitem__(self, path):
        if path == "/":
            return self.groups.setdefault("", _Group(self, "/"))
        return self.groups[path.strip("/")]

    def create_dataset(self, name, data=None, **kw):
        self.datasets[name] = (np.array(data), kw)


def main():
    for name, alias in (("float", float), ("product", np.prod), ("bool", bool), ("int", int)):
        if not hasattr(np, name):
            setattr(np, name, alias)
    h5 = types.ModuleType("h5py")
    h5.File = _File
    sitk = types.ModuleType("SimpleITK")

    class _Img:
        def __init__(self, a): self.a = np.asarray(a)
    sitk.GetImageFromArray = lambda a: _Img(a)
    sitk.GetArrayFromImage = lambda im: im.a
    sitk.BinShrink = lambda im, f_xyz: _Img(O.bin_shrink(im.a, list(f_xyz)[::-1]))   # restatement (unpinned)
    sys.modules.update({"h5py": h5, "SimpleITK": sitk})
    for name in ("dask", "dask.array", "dask.delayed", "networkx", "skimage", "skimage.filters", "scipy.spatial"):
        sys.modules.setdefault(name, mock.MagicMock(name=name))
    sys.path.insert(0, REF)
    imaris = importlib.import_module("mvregfus.imaris")
    mv_utils = importlib.import_module("mvregfus.mv_utils")
    from mvregfus.image_array import ImageArray as RefImageArray

    out = {}
    # ---- .ims pyramid: np_to_ims on a seeded uint16 volume with ragged extents ------------------------------
    vol = blob_volume((70, 53, 91), 301, np.uint16, 0.35)
    imaris.np_to_ims(vol, fname="/tmp/_golden_unused.ims", subsamp=((1, 1, 1), (2, 2, 2), (4, 4, 4), (8, 8, 8)),
                     chunks=((16, 128, 128), (64, 64, 64), (32, 32, 32), (16, 16, 16)))
    f = _File.last
    out["pyr_in"] = vol
    for r in range(4):
        base = "/DataSet/ResolutionLevel %d/TimePoint 0/Channel 0" % r
        out["pyr_data%d" % r] = f.datasets[base + "/Data"][0]
        out["pyr_hist%d" % r] = f.datasets[base + "/Histogram"][0]
        g = f.groups[base.strip("/")]
        out["pyr_minmax%d" % r] = np.array([float(bytes(g.attrs["HistogramMin"]).decode() if not isinstance(g.attrs["HistogramMin"], str) else g.attrs["HistogramMin"]),
                                           float(bytes(g.attrs["HistogramMax"]).decode() if not isinstance(g.attrs["HistogramMax"], str) else g.attrs["HistogramMax"])])
    out["sub_322"] = imaris.subsample_data(vol, (3, 2, 2))
    # ---- bin_stack (axis order, spacing, origin) -----------------------------------------------------------------
    im = RefImageArray(blob_volume((20, 31, 45), 302, np.uint16, 0.2), spacing=np.array([2.0, 0.5, 0.5]),
                       origin=np.array([1.0, -3.0, 10.0]))
    for tag, fac in (("122", [2, 2, 1]), ("341", [1, 4, 3])):   # x, y, z order as in the reference call
        b = mv_utils.bin_stack(im, np.array(fac))
        out["bin_%s_data" % tag] = np.asarray(b)
        out["bin_%s_spacing" % tag], out["bin_%s_origin" % tag] = np.asarray(b.spacing), np.asarray(b.origin)
    out["bin_in"] = np.asarray(im)
    # ---- background subtraction, multiview.py:316 verbatim ----------------------------------------------------
    stack = blob_volume((12, 40, 40), 303, np.uint16, 0.1)
    background_level = 200
    out["bg_in"] = stack
    out["bg_out"] = (stack - background_level) * (stack > background_level)
    np.savez_compressed(os.path.join(HERE, "prepost_golden.npz"), **out)
    print("wrote prepost_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
