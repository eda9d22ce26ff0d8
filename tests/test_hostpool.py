"""Host worker processes of the DCT-weight heuristics (opt-in, MVF_HOST_WORKERS): identical results, clean fallback."""
import numpy as np


def _cells(n=96, v=4, seed=0):
    rng = np.random.default_rng(seed)
    c = rng.random((n, v))
    c[rng.random((n, v)) < 0.3] = 0
    return c


def test_worker_processes_reproduce_the_serial_loop(monkeypatch):
    from mvregfus_b200 import _host_workers, _hostpool
    cells = _cells()
    serial = _host_workers.adapt_cells_chunk((cells, 2, 0.9))
    monkeypatch.setenv("MVF_HOST_WORKERS", "3")
    try:
        tasks = [(cells[a:b], 2, 0.9) for a, b in _hostpool.chunks(len(cells))]
        assert len(tasks) > 1
        par = np.concatenate(_hostpool.pmap(_host_workers.adapt_cells_chunk, tasks))
    finally:
        _hostpool.shutdown()
    assert np.array_equal(serial, par)


def test_block_probe_chunks_match_direct_calls(monkeypatch):
    from mvregfus_b200 import _host_workers, _hostpool
    from tests.helpers import small_case
    views, params, props, sp = small_case(shape=(24, 20, 28), nviews=3, dtype=np.uint16, seed=2)
    locs = list(np.ndindex(3, 3, 3))
    size, qsp = 8, np.array([1.0, 1.0, 1.0])
    direct = _host_workers.block_codes_chunk((locs, np.asarray(sp["origin"], dtype=np.float64), size, qsp, props, np.asarray(params)))
    monkeypatch.setenv("MVF_HOST_WORKERS", "2")
    try:
        tasks = [(locs[a:b], np.asarray(sp["origin"], dtype=np.float64), size, qsp, props, np.asarray(params))
                 for a, b in _hostpool.chunks(len(locs))]
        par = np.concatenate(_hostpool.pmap(_host_workers.block_codes_chunk, tasks))
    finally:
        _hostpool.shutdown()
    assert np.array_equal(direct, par)


def test_dead_workers_fall_back_to_serial(monkeypatch):
    from mvregfus_b200 import _host_workers, _hostpool
    cells = _cells(24)
    monkeypatch.setenv("MVF_HOST_WORKERS", "2")
    try:
        procs = _hostpool._start(2)
        for p in procs:
            p.kill()
            p.wait()
        tasks = [(cells[a:b], 2, 0.9) for a, b in _hostpool.chunks(len(cells))]
        out = np.concatenate(_hostpool.pmap(_host_workers.adapt_cells_chunk, tasks))
    finally:
        _hostpool.shutdown()
    assert np.array_equal(out, _host_workers.adapt_cells_chunk((cells, 2, 0.9)))


def test_block_codes_all_equals_per_block_loop():
    """The all-blocks-at-once view probe of the DCT quality grid gives the per-block loop's codes (mv:3406-3466, 4496-4508)."""
    import numpy as np
    from mvregfus_b200 import _host_workers as hw
    from mvregfus_b200 import synthetic
    shape = (96, 80, 120)
    params = synthetic.view_params(shape, 4, seed=3, odd_extra_deg=7.0)
    props = [{"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)} for _ in range(4)]
    sp_origin, size, qspacing = np.array([-20.0, -10.0, -30.0]), 16, np.array([2.0, 2.0, 2.0])
    locs = list(np.ndindex(5, 4, 6))
    loop = hw.block_codes_chunk((locs, sp_origin, size, qspacing, props, params))
    origins = sp_origin + np.array(locs, dtype=np.int64) * size * qspacing
    fast = hw.block_codes_all(props, params, origins, np.array([size] * 3).astype(np.int64), qspacing, 2)
    assert np.array_equal(loop, fast)
    assert set(np.unique(fast)) == {0, 1, 2}
