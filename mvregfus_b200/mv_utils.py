"""Host-side affine-parameter algebra of the fusion path (float64, z,y,x order).

A view's registration is a 12-vector ``p = [A (9, row-major), c (3)]`` mapping *fused-frame physical
coordinates* to *view physical coordinates*: ``y_view = A @ x_fused + c``
(reference: mvregfus/mv_utils.py:96-133, 280-285).
"""
import numpy as np


def _ndim_of(params):
    n = len(params)
    if n == 12:
        return 3
    if n == 6:
        return 2
    raise ValueError("affine parameter vector must have 12 (3-D) or 6 (2-D) entries, got %d" % n)


def matrix_to_params(M):
    """(ndim+1)x(ndim+1) homogeneous matrix -> parameter vector (reference: mv_utils.py:96)."""
    M = np.asarray(M)
    nd = M.shape[0] - 1
    return np.concatenate([M[:nd, :nd].ravel(), M[:nd, nd]]).astype(np.float64)


def params_to_matrix(params):
    """Parameter vector -> homogeneous matrix (reference: mv_utils.py:108)."""
    params = np.asarray(params, dtype=np.float64)
    nd = _ndim_of(params)
    M = np.eye(nd + 1)
    M[:nd, :nd] = params[: nd * nd].reshape(nd, nd)
    M[:nd, nd] = params[nd * nd:]
    return M


def params_invert_coordinates(params):
    """Flip the axis order (zyx <-> xyz) of a parameter vector (reference: mv_utils.py:122)."""
    M = params_to_matrix(params)
    nd = M.shape[0] - 1
    M[:nd, :nd] = M[:nd, :nd][::-1, ::-1]
    M[:nd, nd] = M[:nd, nd][::-1]
    return matrix_to_params(M)


def invert_params(params):
    """Parameters of the inverse affine map (reference: mv_utils.py:129)."""
    M = params_to_matrix(params)
    nd = M.shape[0] - 1
    Ainv = np.linalg.inv(M[:nd, :nd])
    out = np.eye(nd + 1)
    out[:nd, :nd] = Ainv
    out[:nd, nd] = -np.dot(Ainv, M[:nd, nd])
    return matrix_to_params(out)


def transform_points(pts, p):
    """Apply ``A @ pt + c`` to a list of points (reference: mv_utils.py:280)."""
    p = np.asarray(p, dtype=np.float64)
    nd = _ndim_of(p)
    A = p[: nd * nd].reshape(nd, nd)
    return np.asarray(pts, dtype=np.float64) @ A.T + p[nd * nd:]


def index_space_affine(p, in_spacing, in_origin, out_spacing, out_origin):
    """Index-space map of the resampling: ``u_view_index = matrix @ i_out + offset``.

    ``matrix' = S_v^-1 A S_out``, ``offset' = S_v^-1 (c - o_v + A o_out)``
    (reference: mvregfus/multiview.py:1486-1496).
    """
    p = np.asarray(p, dtype=np.float64)
    nd = _ndim_of(p)
    A = p[: nd * nd].reshape(nd, nd)
    c = p[nd * nd:]
    Sx = np.diag(np.asarray(out_spacing, dtype=np.float64))
    Sy_inv = np.linalg.inv(np.diag(np.asarray(in_spacing, dtype=np.float64)))
    matrix = np.dot(Sy_inv, np.dot(A, Sx))
    offset = np.dot(Sy_inv, c - np.asarray(in_origin, dtype=np.float64) + np.dot(A, np.asarray(out_origin, dtype=np.float64)))
    return matrix, offset


def bin_stack(im, bin_factors=np.array([1, 1, 1])):
    """mv_utils.bin_stack (reference: mv_utils.py:288-311) on the device: see mvregfus_b200.prepost.bin_stack."""
    from .prepost import bin_stack as _bin_stack
    return _bin_stack(im, bin_factors)
