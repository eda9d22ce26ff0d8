"""Pure numpy/scipy pieces of the DCT-weight host heuristics, importable without torch so that worker processes
stay light.  The functions are the single implementation: fusion.py / dct.py call them directly (serial) or through
_hostpool (one chunk of blocks / cells per task).  Per block and per cell the arithmetic is exactly the reference's
(mv:3406-3466, 4155-4210), so distributing the loop over processes cannot change a bit of the result."""
import numpy as np
from scipy import optimize


def probe_flags(orig_props, p, stack_properties, n_points):
    rel = np.linspace(0, 1, n_points)
    grid = np.stack(np.meshgrid(rel, rel, rel, indexing="ij"), -1).reshape(-1, 3)
    p = np.asarray(p, dtype=np.float64)
    A, c = p[:9].reshape(3, 3), p[9:]
    flags = np.empty(len(grid), dtype=bool)
    size = np.asarray(orig_props["size"])
    for i, g in enumerate(grid):  # same operation order as the reference loop (bit-exact flags)
        phys = stack_properties["origin"] + np.array(g) * stack_properties["size"] * stack_properties["spacing"]
        pix = (np.dot(A, phys) + c - orig_props["origin"]) / orig_props["spacing"]
        flags[i] = bool(np.all((pix >= 0) & (pix < size)))
    return flags


def probe_inside_blocks(orig_props, p, origins, size, spacing, n_points):
    """probe_flags for many target boxes of one size at once (origins: (B, 3) physical corners) -> bool (B, n^3).  Same
    arithmetic per point as the per-box loop, evaluated column-wise on all points of all boxes."""
    rel = np.linspace(0, 1, n_points)
    grid = np.stack(np.meshgrid(rel, rel, rel, indexing="ij"), -1).reshape(-1, 3)              # (P, 3)
    origins = np.asarray(origins, dtype=np.float64)
    phys = origins[:, None, :] + grid[None, :, :] * np.asarray(size) * np.asarray(spacing)     # (B, P, 3)
    p = np.asarray(p, dtype=np.float64)
    A, c = p[:9].reshape(3, 3), p[9:]
    inside = np.ones(phys.shape[:2], dtype=bool)
    vsize = np.asarray(orig_props["size"])
    for d in range(3):
        y = A[d, 0] * phys[..., 0] + A[d, 1] * phys[..., 1] + A[d, 2] * phys[..., 2]
        pix = (y + c[d] - orig_props["origin"][d]) / orig_props["spacing"][d]
        inside &= (pix >= 0) & (pix < vsize[d])
    return inside


def probe_any_inside_blocks(orig_props, p, origins, size, spacing, n_points):
    """The probe-point early-out of get_weights_simple (mv:3491-3532) for every block of a volume: True where any of
    the n^3 probe points of the block falls inside the view."""
    return probe_inside_blocks(orig_props, p, origins, size, spacing, n_points).any(axis=1)


def block_codes_all(orig_stack_propertiess, params, origins, size, spacing, n_points=2):
    """blocks_inside codes (1 all probe points inside / 2 some / 0 none, mv:3406-3466) of many boxes -> int64 (B, V)."""
    codes = np.zeros((len(origins), len(params)), dtype=np.int64)
    for v, (osp, p) in enumerate(zip(orig_stack_propertiess, params)):
        f = probe_inside_blocks(osp, p, origins, size, spacing, n_points)
        codes[:, v] = np.where(f.all(axis=1), 1, np.where(f.any(axis=1), 2, 0))
    return codes


def blocks_inside_codes(orig_stack_propertiess, params, stack_properties, n_points_per_dim=5):
    """1 all probe points inside / 0 none / 2 partial, per view (mv:3406-3466)."""
    codes = []
    for osp, p in zip(orig_stack_propertiess, params):
        f = probe_flags(osp, p, stack_properties, n_points_per_dim)
        codes.append(1 if f.all() else (2 if f.any() else 0))
    return np.array(codes)


def block_codes_chunk(task):
    """Codes of a list of coarse-grid blocks: task = (locs, origin, size, qspacing, orig_props, params)."""
    locs, sp_origin, size, qspacing, orig_props, params = task
    out = np.zeros((len(locs), len(params)), dtype=np.int64)
    for n, loc in enumerate(locs):
        origin = np.array([sp_origin[i] + loc[i] * size * qspacing[i] for i in range(3)])
        bprops = {"size": np.array([size] * 3).astype(np.int64), "spacing": np.array(qspacing), "origin": origin}
        out[n] = blocks_inside_codes(orig_props, params, bprops, 2)
    return out


def adapt_weights(ws, how_many_best_views, cumulative_weight_best_views):
    """Exponent fit of one coarse cell (mv:4155-4210)."""
    wsum = np.sum(ws, 0)
    if wsum > 0:
        ws = ws / wsum
    if not (np.sum(ws > 0) >= 2):
        return ws
    wfs = np.sort(ws.astype(np.float64), axis=0)

    def energy(exp):
        tmpw = wfs ** exp[0]
        tmpw = tmpw / np.nansum(tmpw, 0)
        nsum = np.nansum(tmpw[-int(how_many_best_views):], (-1))
        return np.abs(np.nansum(nsum) - cumulative_weight_best_views)

    res = optimize.minimize(energy, [0.5], bounds=[[0.01, 100]], method='L-BFGS-B', options={'maxiter': 10})
    ws = np.array([ws[i] ** res.x[0] for i in range(len(ws))])
    wsum = np.sum(ws, 0)
    if wsum > 0:
        ws = ws / wsum
    return ws


def adapt_cells_chunk(task):
    """task = (cells (n, V) float64, how_many_best_views, cumulative_weight_best_views) -> (n, V) float64."""
    cells, how_many, cumulative = task
    out = np.zeros(cells.shape, dtype=np.float64)
    for n in range(len(cells)):
        out[n] = adapt_weights(cells[n], how_many, cumulative)
    return out


def _serve():
    """Worker loop of mvregfus_b200._hostpool: length-prefixed pickles (function name, task) in, (ok, result) out."""
    import pickle
    import struct
    import sys
    inp, out = sys.stdin.buffer, sys.stdout.buffer
    sys.stdout = sys.stderr   # stray prints must not corrupt the pipe
    table = {"block_codes_chunk": block_codes_chunk, "adapt_cells_chunk": adapt_cells_chunk}
    while True:
        head = inp.read(8)
        if len(head) != 8:
            return
        (n,) = struct.unpack("<Q", head)
        name, task = pickle.loads(inp.read(n))
        try:
            res = (True, table[name](task))
        except Exception as exc:  # noqa: BLE001
            res = (False, "%s: %s" % (type(exc).__name__, exc))
        data = pickle.dumps(res, protocol=pickle.HIGHEST_PROTOCOL)
        out.write(struct.pack("<Q", len(data)) + data)
        out.flush()


if __name__ == "__main__":
    _serve()
