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
