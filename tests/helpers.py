"""Shared helpers of the parity tests."""
import numpy as np

from mvregfus_b200.image_array import ImageArray
from mvregfus_b200 import synthetic
from oracle import mvregfus_oracle as O


def small_case(shape=(40, 36, 44), nviews=2, dtype=np.uint16, seed=3, odd_extra_deg=0.0, spacing=1.0):
    """Views, params, view props and the fused-frame stack properties of a small synthetic case."""
    views, params, props = synthetic.make_views(shape, nviews, dtype=dtype, seed=seed, odd_extra_deg=odd_extra_deg)
    views = [ImageArray(v, spacing=p["spacing"], origin=p["origin"]) for v, p in zip(views, props)]
    sp = O.calc_stack_properties_from_views_and_params(props, params, spacing=[spacing] * 3)
    return views, params, props, sp


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor)  (SURVEY §8d parity norm)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def int_mismatch(a, b):
    a = np.asarray(a).astype(np.int64)
    b = np.asarray(b).astype(np.int64)
    d = np.abs(a - b)
    return int(d.max()), float(np.mean(d > 0))
