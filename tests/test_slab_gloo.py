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


# ---- v4 schedule (slab.ZZLayout): one estimate exchange of 2rp planes per iteration, static view halos -----------------
def _ext_ref(g, nz, K):
    g = np.asarray(g)
    return np.clip(np.where(g < 0, nz - 3 - K - g, np.where(g >= nz, 2 * nz - 2 - g, g)), 0, nz - 1)


def _zz_check(lay, est, view, nz, K, rp):
    """est / view hold GLOBAL plane numbers: every plane the two chained blurs of this rank touch must be the plane the
    reference's boundary map names for the whole volume."""
    zm, rm = lay.zmap1(), lay.rmap()
    for t in range(-rp, lay.nzl + rp):
        g = lay.z0 + t
        if 0 <= g < nz:
            if view[rm[t + rp]] != g:
                return False
            for d in range(-rp, rp + 1):
                if est[zm[t + d + 2 * rp]] != _ext_ref(g + d, nz, K):
                    return False
        else:
            q = nz - 3 - K - g if g < 0 else 2 * nz - 2 - g
            if -1 - rm[t + rp] != np.clip(q - lay.qa, 0, K):
                return False
    for zmj, v0, cnt, k0 in lay.side_jobs():
        for i in range(cnt):
            q = lay.qa + k0 + i
            if view[v0 + i] != q or any(est[zmj[i + d + rp]] != _ext_ref(q + d, nz, K) for d in range(-rp, rp + 1)):
                return False
    return True


@pytest.mark.parametrize("world,nz,K,rp", [(1, 40, 7, 3), (2, 72, 7, 3), (3, 90, 9, 6), (8, 512, 19, 9)])
def test_zz_layout_emulated_ranks(world, nz, K, rp):
    lays = [slab.ZZLayout(nz, world, r, K, rp) for r in range(world)]
    est = [np.full(l.height, -1) for l in lays]
    view = [np.full(l.vheight, -1) for l in lays]
    for l, e, v in zip(lays, est, view):
        e[l.front:l.front + l.nzl] = np.arange(l.z0, l.z1)
        v[l.rp:l.rp + l.nzl] = np.arange(l.z0, l.z1)
    slab.run_transfers_local(est, lays, "est_transfers")
    slab.run_transfers_local(view, lays, "view_transfers")
    assert all(_zz_check(l, e, v, nz, K, rp) for l, e, v in zip(lays, est, view))


def test_zz_layout_rejects_thin_slabs():
    with pytest.raises(ValueError):
        slab.ZZLayout(64, 4, 0, 19, 9)      # 16-plane slabs cannot lend 18 halo planes


def _zz_worker(rank, world, port, nz, K, rp, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lay = slab.ZZLayout(nz, world, rank, K, rp)
    est = torch.full((lay.height, 2, 2), -1.0)
    est[lay.front:lay.front + lay.nzl] = torch.arange(lay.z0, lay.z1, dtype=torch.float32)[:, None, None]
    view = torch.full((lay.vheight, 2, 2), -1, dtype=torch.int16)
    view[rp:rp + lay.nzl] = torch.arange(lay.z0, lay.z1, dtype=torch.int16)[:, None, None]
    slab.run_transfers_dist(est, lay, "est_transfers")
    slab.run_transfers_dist(view, lay, "view_transfers")
    out[rank] = _zz_check(lay, est[:, 0, 0].numpy().astype(np.int64), view[:, 0, 0].numpy().astype(np.int64), nz, K, rp)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nz,K,rp", [(2, 72, 7, 3), (3, 96, 9, 6)])
def test_zz_transfers_gloo(world, nz, K, rp):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_zz_worker, args=(world, port, nz, K, rp, out), nprocs=world, join=True)
    assert all(out[r] is True for r in range(world)), dict(out)


def test_rl_slab_bounds_balance_and_layout():
    """Work-balanced slabs of the v4 schedule: they tile the volume, rank 0 (extra wrap-around planes + boundary ratios)
    and the last rank (boundary ratios) get fewer planes, thin volumes fall back to equal slabs, and the plane bookkeeping holds for unequal slabs."""
    b = slab.rl_slab_bounds(512, 8, 19, 9)
    sizes = [z1 - z0 for z0, z1 in b]
    assert b[0][0] == 0 and b[-1][1] == 512 and all(x[1] == y[0] for x, y in zip(b[:-1], b[1:]))
    assert sizes[0] < sizes[-1] < min(sizes[1:-1]) and max(sizes[1:-1]) - min(sizes[1:-1]) <= 1
    assert slab.rl_slab_bounds(64, 4, 19, 9) == slab.slab_bounds(64, 4)
    assert slab.rl_slab_bounds(40, 1, 7, 3) == [(0, 40)]
    nz, world, K, rp = 200, 3, 9, 6
    bounds = slab.rl_slab_bounds(nz, world, K, rp)
    lays = [slab.ZZLayout(nz, world, r, K, rp, bounds) for r in range(world)]
    assert [l.nzl for l in lays] == [z1 - z0 for z0, z1 in bounds]
    est = [np.full(l.height, -1) for l in lays]
    view = [np.full(l.vheight, -1) for l in lays]
    for l, e, v in zip(lays, est, view):
        e[l.front:l.front + l.nzl] = np.arange(l.z0, l.z1)
        v[l.rp:l.rp + l.nzl] = np.arange(l.z0, l.z1)
    slab.run_transfers_local(est, lays, "est_transfers")
    slab.run_transfers_local(view, lays, "view_transfers")
    assert all(_zz_check(l, e, v, nz, K, rp) for l, e, v in zip(lays, est, view))
    with pytest.raises(ValueError):
        slab.ZZLayout(nz, world, 0, K, rp, [(0, 60), (70, 140), (140, 200)])
