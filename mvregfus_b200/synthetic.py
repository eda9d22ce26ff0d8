"""Seeded synthetic multi-view data: a smooth random ground truth inside an ellipsoid
(about half the voxels are exact zeros), observed by V views rotated about the y axis with slight
anisotropic scaling and sub-voxel translation.  Host-side data generation only (numpy/scipy)."""
import numpy as np
from scipy import ndimage


def rotation_about_y(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, 0.0, -s], [0.0, 1.0, 0.0], [s, 0.0, c]])  # z,y,x order; y is the Z1 rotation axis


def view_params(shape, nviews, seed=0, odd_extra_deg=0.0, jitter=True):
    """12-vectors p_v (fused physical -> view physical) for views of ``shape`` (spacing 1, origin 0)."""
    rng = np.random.default_rng(seed)
    centre = (np.asarray(shape, dtype=np.float64) - 1) / 2.0
    params = []
    for v in range(nviews):
        theta = 2 * np.pi * v / nviews + (np.deg2rad(odd_extra_deg) if v % 2 else 0.0)
        eps = rng.uniform(-0.01, 0.01, 3) if jitter else np.zeros(3)
        A = rotation_about_y(theta) @ np.diag(1.0 + eps)
        t = rng.uniform(-3, 3, 3) if jitter else np.zeros(3)
        c = centre - A @ centre + t
        params.append(np.concatenate([A.ravel(), c]))
    return np.array(params)


def ground_truth(shape, seed=0, scale=4000.0):
    rng = np.random.default_rng(seed)
    shape = tuple(int(s) for s in shape)
    fine = rng.random(tuple(max(2, s // 16) for s in shape))
    coarse = rng.random(tuple(max(2, s // 128) for s in shape))
    g = ndimage.zoom(fine, [s / f for s, f in zip(shape, fine.shape)], order=1)
    g = g * ndimage.zoom(coarse, [s / f for s, f in zip(shape, coarse.shape)], order=1)
    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, s) for s in shape], indexing="ij", sparse=True)
    mask = (zz / 0.9) ** 2 + (yy / 0.9) ** 2 + (xx / 0.9) ** 2 <= 1.0
    return (g * scale * mask).astype(np.float32)


def make_views(shape, nviews, dtype=np.uint16, seed=0, odd_extra_deg=0.0, jitter=True):
    """Returns (views list of ndarray, params (V,12), view stack-properties list).

    Every view is the ground truth resampled into the view's own frame (order-1, done here on the host
    with scipy because this is input generation, not the measured path)."""
    params = view_params(shape, nviews, seed=seed, odd_extra_deg=odd_extra_deg, jitter=jitter)
    gt = ground_truth(shape, seed=seed)
    views, props = [], []
    for v in range(nviews):
        A, c = params[v][:9].reshape(3, 3), params[v][9:]
        Ainv = np.linalg.inv(A)
        # view index y -> fused index x = A^-1 (y - c); sample the ground truth there
        img = ndimage.affine_transform(gt, Ainv, -Ainv @ c, output_shape=tuple(shape), order=1)
        if np.dtype(dtype) == np.uint16:
            img = np.clip(np.rint(img), 0, 4095).astype(np.uint16)
        else:
            img = img.astype(np.float32)
        views.append(img)
        props.append({"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)})
    return views, params, props
