"""ImageArray: the boundary value type of the view-fusion path.

An ``np.ndarray`` subclass that carries ``spacing``/``origin`` (float64, **z,y,x** order) and
``rotation`` and keeps them through slicing, ufuncs and pickling.  Mirrors the contract of the
reference type (reference: mvregfus/image_array.py:6-57) so arrays produced by this package can be
handed to unmodified reference code and vice versa.
"""
import numpy as np

_DEFAULTS = (("spacing", (1.0, 1.0, 1.0)), ("origin", (0.0, 0.0, 0.0)))


class ImageArray(np.ndarray):
    """ndarray + physical metadata (reference: mvregfus/image_array.py:6)."""

    def __new__(cls, input_array, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), rotation=0.0, meta=None):
        arr = np.asarray(input_array).view(cls)
        if meta is not None:
            arr.spacing, arr.origin, arr.rotation = meta["spacing"], meta["origin"], meta["rotation"]
        else:
            arr.spacing = np.array(spacing).astype(np.float64)
            arr.origin = np.array(origin).astype(np.float64)
            arr.rotation = float(rotation)
        return arr

    def __array_finalize__(self, parent):
        if parent is None:
            return
        for name, default in _DEFAULTS:
            setattr(self, name, getattr(parent, name, np.array(default, dtype=np.float64)))
        self.rotation = getattr(parent, "rotation", 0.0)

    # pickling: append the three metadata fields to ndarray's own state tuple
    def __reduce__(self):
        ctor, args, state = super().__reduce__()
        return ctor, args, state + (self.spacing, self.origin, self.rotation)

    def __setstate__(self, state):
        self.spacing, self.origin, self.rotation = state[-3:]
        super().__setstate__(state[:-3])

    def get_meta_dict(self):
        return {"spacing": self.spacing[:], "origin": self.origin[:], "rotation": float(self.rotation)}

    def get_info(self):
        """``{spacing, origin, rotation, size}`` — the ``stack_properties`` of this array."""
        return {
            "spacing": np.array(self.spacing[:]),
            "origin": np.array(self.origin),
            "rotation": float(self.rotation),
            "size": np.array(self.shape),
        }
