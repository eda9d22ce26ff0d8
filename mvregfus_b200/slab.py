"""z-slab sharding of the fused frame across the GPUs of one box.

Resampling, weights and weighted fusion are pointwise in the fused frame, so a rank simply produces the index
box ``[z0, z1) x ny x nx`` (``out_origin``/``out_dims`` of the C-ABI): no data-path collective.  Multi-view RL
needs the PSF-radius halo planes of its convolution inputs: a ring exchange (rank 0's lower halo comes from the
last rank because of the reference's wrap-around boundary) plus one scalar all-reduce for the
convergence test.  Everything here is host logic + torch.distributed; it runs on CPU tensors with gloo in the
tests and on CUDA tensors with NCCL on the box.
"""
import numpy as np
import torch


def slab_bounds(nz, world):
    """Contiguous z ranges [(z0, z1)] of `world` slabs, sizes differing by at most one plane."""
    base, rem = divmod(int(nz), int(world))
    bounds, z = [], 0
    for r in range(world):
        h = base + (1 if r < rem else 0)
        bounds.append((z, z + h))
        z += h
    return bounds


def rl_slab_bounds(nz, world, ksize_z, rp):
    """z ranges for the v4 RL schedule, balanced by WORK rather than planes.  Every rank blurs its halo planes once per
    x/y group besides its own planes; rank 0 holds 5*rp halo planes (3*rp below the top of the volume for the wrap-around,
    2*rp above its slab) against 4*rp in the middle and 2*rp on the last rank, and ranks 0 and last compute the boundary
    ratios.  Measured on 8 B200 (tools/rl_slab_phases.py, configs[3]): a halo plane costs 0.3 own planes (x/y passes
    only), the boundary ratios 10.5 own planes (0.1 ms per view) - with equal slabs rank 0 keeps everybody waiting at the
    barrier for 0.3 ms of 3.6."""
    nz, world = int(nz), int(world)
    if world == 1:
        return [(0, nz)]
    halo = np.array([5 * rp] + [4 * rp] * (world - 2) + [2 * rp], dtype=np.float64)
    side = np.array([10.5] + [0.0] * (world - 2) + [10.5])
    fixed = 0.3 * halo + side
    own = (nz + fixed.sum()) / world - fixed            # equal total cost
    n = np.floor(own).astype(int)
    order = np.argsort(-(own - n))
    for i in order[:nz - int(n.sum())]:
        n[i] += 1
    need = np.full(world, 2 * rp)
    need[-1] = max(need[-1], ksize_z + 2 + rp)
    if np.any(n < need) or n.sum() != nz:                # thin slabs: the plain partition (the planners reject what cannot work)
        return slab_bounds(nz, world)
    z = np.concatenate([[0], np.cumsum(n)])
    return [(int(a), int(b)) for a, b in zip(z[:-1], z[1:])]


def source_brick(matrix, offset, view_shape, z0, z1, ny, nx):
    """Index box of a view that the slab [z0,z1) touches: the reference's own chunk bounding box (corners
    through matrix', +-2, clipped; mv:1552-1571) applied to the slab.  Returns (lo[3], hi[3]) int64."""
    corners = np.array([[z, y, x] for z in (z0, z1) for y in (0, ny) for x in (0, nx)], dtype=np.float64)
    src = corners @ np.asarray(matrix, dtype=np.float64).T + np.asarray(offset, dtype=np.float64)
    lo = np.clip(src.min(0) - 2, 0, view_shape).astype(np.int64)
    hi = np.clip(src.max(0) + 2, 0, view_shape).astype(np.int64)
    return lo, hi


def _ext(t, n, K):
    """Reference boundary index map (reflect above, wrap into the far reflected tail below), clamped."""
    i = np.where(t < 0, n - 3 - K - t, np.where(t >= n, 2 * n - 2 - t, t))
    return np.clip(i, 0, n - 1)


class HaloPlan:
    """Plane bookkeeping of one rank's halo-padded slab buffer: planes [0,rp) = lower halo, [rp, rp+nzl) = own
    planes, [rp+nzl, nzl+2rp) = upper halo."""

    def __init__(self, nz, world, rank, ksize_z, rp, bounds=None):
        self.nz, self.world, self.rank, self.K, self.rp = int(nz), int(world), int(rank), int(ksize_z), int(rp)
        self.bounds = slab_bounds(nz, world) if bounds is None else [(int(a), int(b)) for a, b in bounds]
        self.z0, self.z1 = self.bounds[rank]
        self.nzl = self.z1 - self.z0
        if world > 1 and min(b - a for a, b in self.bounds) < max(2 * rp + 1, ksize_z + 3 + rp):
            raise ValueError("slabs of %d planes are too thin for a halo of %d planes" % (min(b - a for a, b in self.bounds), rp))
        self.lower_src = (rank - 1) % world   # ring: rank 0's lower halo comes from the last rank
        self.upper_src = rank + 1 if rank + 1 < world else None

    def zmap(self):
        """Extended local position t in [-rp, nzl+rp) -> plane of the padded buffer."""
        t = np.arange(-self.rp, self.nzl + self.rp)
        g = self.z0 + t
        plane = t + self.rp                                   # default: own planes / halos filled by neighbours
        if self.world == 1:
            return (_ext(g, self.nz, self.K) + self.rp).astype(np.int32)
        above = g >= self.nz                                  # top rank: reflection inside its own slab
        plane = np.where(above, _ext(g, self.nz, self.K) - self.z0 + self.rp, plane)
        # g < 0 (rank 0): the lower halo planes hold what the last rank sent (wrap), same positions
        return np.clip(plane, 0, self.nzl + 2 * self.rp - 1).astype(np.int32)

    def planes_for_upper_neighbour(self):
        """Own-plane indices (0-based within the slab) that fill the LOWER halo of rank+1 (or of rank 0 when this
        is the last rank), in halo order."""
        if self.rank + 1 < self.world:
            return np.arange(self.nzl - self.rp, self.nzl)
        g = np.arange(-self.rp, 0)                             # rank 0's positions below the volume
        return _ext(g, self.nz, self.K) - self.z0              # wrap into this (last) rank's slab

    def planes_for_lower_neighbour(self):
        """Own-plane indices that fill the UPPER halo of rank-1 (none for rank 0)."""
        return np.arange(0, self.rp) if self.rank > 0 else None


class _StagedRecv:
    """Request of a receive that lands in a host buffer first (backend without device P2P, i.e. gloo with CUDA
    volumes in the single-GPU tests): wait() finishes the transfer and copies the planes to the device."""

    def __init__(self, req, host, dst):
        self.req, self.host, self.dst = req, host, dst

    def wait(self):
        self.req.wait()
        self.dst.copy_(self.host)


def exchange_halos_begin(buf, plan, group=None):
    """Start filling the halo planes of `buf` ((nzl+2rp, ny, nx), own planes already written) from the ring
    neighbours; returns the pending requests (empty for a single rank).  Work that only touches own planes may be
    queued before :func:`exchange_halos_end` - with NCCL the transfers then run beside it."""
    import torch.distributed as dist
    rp, nzl = plan.rp, plan.nzl
    if plan.world == 1:
        return []
    staged = buf.is_cuda and dist.get_backend(group) == "gloo"   # gloo moves host memory only
    up_idx = torch.as_tensor(plan.planes_for_upper_neighbour() + rp, device=buf.device, dtype=torch.long)
    send_up = buf.index_select(0, up_idx).contiguous()
    recvs = [(buf[0:rp], plan.lower_src)]
    sends = [(send_up, (plan.rank + 1) % plan.world)]
    if plan.rank > 0:
        sends.append((buf[rp:2 * rp].contiguous(), plan.rank - 1))
    if plan.upper_src is not None:
        recvs.append((buf[rp + nzl:rp + nzl + rp], plan.upper_src))
    if not staged:
        ops = [dist.P2POp(dist.isend, t, peer, group=group) for t, peer in sends] + \
              [dist.P2POp(dist.irecv, t, peer, group=group) for t, peer in recvs]
        return list(dist.batch_isend_irecv(ops))
    reqs = [dist.isend(t.cpu(), peer, group=group) for t, peer in sends]
    for t, peer in recvs:
        host = torch.empty(t.shape, dtype=t.dtype)
        reqs.append(_StagedRecv(dist.irecv(host, peer, group=group), host, t))
    return reqs


def exchange_halos_end(reqs):
    for req in reqs:
        req.wait()


def exchange_halos(buf, plan, group=None):
    """Blocking form: fill the halo planes of `buf` from the ring neighbours."""
    exchange_halos_end(exchange_halos_begin(buf, plan, group))


def exchange_halos_local(bufs, plans):
    """The same plane movement as :func:`exchange_halos` between the buffers of ranks emulated in ONE process
    (tests on a single GPU): bufs[r] / plans[r] belong to emulated rank r."""
    world = len(plans)
    if world == 1:
        return
    for r, (buf, plan) in enumerate(zip(bufs, plans)):
        rp, nzl = plan.rp, plan.nzl
        src = plan.lower_src
        idx = torch.as_tensor(plans[src].planes_for_upper_neighbour() + plans[src].rp, device=buf.device, dtype=torch.long)
        buf[0:rp] = bufs[src].index_select(0, idx)
        if plan.upper_src is not None:
            up = plan.upper_src
            buf[rp + nzl:rp + nzl + rp] = bufs[up][plans[up].rp:2 * plans[up].rp]


class SymmetricHalo:
    """Halo planes pulled straight out of the neighbours' buffers over NVLink / NVSwitch: the halo-padded volumes live
    in torch symmetric memory (every rank maps its ring neighbours' allocations), a halo fill is two device-to-device
    copies issued by the receiving rank plus a device-side barrier on the stream - no NCCL call, no host round trip
    and nothing that cannot be captured in a CUDA graph.  One allocation holds `nbuf` padded volumes of the same
    (maximal) height on every rank."""

    def __init__(self, nbuf, plan, ny, nx, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.plan, self.nbuf = plan, nbuf
        group = dist.group.WORLD if group is None else group
        height = max(b - a for a, b in plan.bounds) + 2 * plan.rp      # same shape on every rank
        self.shape = (nbuf, height, int(ny), int(nx))
        self.store = symm.empty(self.shape, dtype=torch.float32, device=device)
        self.store.zero_()
        self.hdl = symm.rendezvous(self.store, group)
        self.bufs = [self.store[i][:plan.nzl + 2 * plan.rp] for i in range(nbuf)]
        rp = plan.rp
        # source planes inside the neighbours' padded buffers (their own planes start at rp)
        lower = HaloPlan(plan.nz, plan.world, plan.lower_src, plan.K, rp, plan.bounds)
        self.lower_peer = self.hdl.get_buffer(plan.lower_src, self.shape, torch.float32)
        idx = lower.planes_for_upper_neighbour() + rp
        self.lower_slice = slice(int(idx[0]), int(idx[-1]) + 1) if np.all(np.diff(idx) == 1) else None
        self.lower_idx = torch.as_tensor(idx, device=device, dtype=torch.long)
        self.upper_peer = None
        if plan.upper_src is not None:
            self.upper_peer = self.hdl.get_buffer(plan.upper_src, self.shape, torch.float32)

    def barrier(self):
        """All ranks of the group meet on their current streams (signal pads in symmetric memory)."""
        self.hdl.barrier(channel=0)

    def pull(self, i):
        """Fill the halo planes of buffer i from the neighbours' own planes (current stream)."""
        rp, nzl = self.plan.rp, self.plan.nzl
        dst = self.bufs[i]
        if self.lower_slice is not None:
            dst[0:rp].copy_(self.lower_peer[i][self.lower_slice])
        else:   # rank 0: the reference's wrap-around picks reflected planes of the last rank
            torch.index_select(self.lower_peer[i], 0, self.lower_idx, out=dst[0:rp])
        if self.upper_peer is not None:
            dst[rp + nzl:rp + nzl + rp].copy_(self.upper_peer[i][rp:2 * rp])


# ---------------------------------------------------------------------------------------------------------------------
# v4 schedule (rl.zz_mode): per view  T1 = xy(estimate), T2 = z(ratio(z(T1))), correction = epilogue(xy(T2)).
# The ratio volume is never materialised, so it is never exchanged either: ONE halo exchange per iteration (the
# estimate, 2*rp planes on each side; rank 0 additionally pulls the planes near the top of the volume that its
# wrap-around positions refer to), against 1 + V exchanges of the v3 schedule.  The views need rp static halo planes.
# ---------------------------------------------------------------------------------------------------------------------
class ZZLayout:
    """Plane bookkeeping of one rank for the v4 schedule.

    Padded estimate / T1 buffers (same height on every rank): [front: 3rp planes][own: nzl][back: 2rp planes].  The
    lower halo sits right-aligned in `front`; on rank 0 (world > 1) `front` holds the 3rp planes starting at
    nz-2-K-rp instead (sources of the wrap-around).  Padded (masked) views: [rp][own][rp]; on rank 0 the front block
    holds planes [nz-2-K, nz-2-K+rp)."""

    def __init__(self, nz, world, rank, ksize_z, rp, bounds=None):
        self.nz, self.world, self.rank, self.K, self.rp = int(nz), int(world), int(rank), int(ksize_z), int(rp)
        self.bounds = slab_bounds(nz, world) if bounds is None else [(int(a), int(b)) for a, b in bounds]
        if len(self.bounds) != self.world or self.bounds[0][0] != 0 or self.bounds[-1][1] != self.nz or \
                any(a[1] != b[0] for a, b in zip(self.bounds[:-1], self.bounds[1:])):
            raise ValueError("slab bounds must tile [0, nz) in rank order")
        self.z0, self.z1 = self.bounds[rank]
        self.nzl = self.z1 - self.z0
        self.front, self.back = 3 * self.rp, 2 * self.rp
        self.height = self.front + self.nzl + self.back
        self.height_max = self.front + max(b - a for a, b in self.bounds) + self.back
        self.vheight = self.nzl + 2 * self.rp
        self.qa = self.nz - 2 - self.K                    # first plane whose ratio the positions outside the volume use
        self.far0 = self.qa - self.rp
        if self.nz < self.K + 3:
            raise ValueError("volume of %d planes is thinner than PSF size + 3" % self.nz)
        if world > 1:
            thin = min(b - a for a, b in self.bounds)
            if thin < 2 * self.rp or self.bounds[-1][1] - self.bounds[-1][0] < self.K + 2 + self.rp or self.far0 < 0:
                raise ValueError("slabs of %d planes are too thin for halos of %d planes" % (thin, 2 * self.rp))

    # ---- global plane -> local plane ---------------------------------------------------------------------------------
    def est_local(self, g):
        """Local plane of global plane(s) g in the padded estimate / T1 buffer, -1 where the rank does not hold it."""
        g = np.asarray(g)
        near = (g >= max(0, self.z0 - self.back)) & (g < min(self.nz, self.z1 + self.back)) & (g >= self.z0 - self.back)
        if self.world > 1 and self.rank == 0:
            near &= g >= 0
        loc = np.where(near, self.front + g - self.z0, -1)
        if self.world > 1 and self.rank == 0:
            far = (g >= self.far0) & (g < min(self.nz, self.far0 + self.front)) & (loc < 0)
            loc = np.where(far, g - self.far0, loc)
        if self.world > 1 and self.rank == self.world - 1:
            loc = np.where(g >= self.nz, -1, loc)
        return loc

    def view_local(self, g):
        g = np.asarray(g)
        near = (g >= max(0, self.z0 - self.rp)) & (g < min(self.nz, self.z1 + self.rp))
        loc = np.where(near, self.rp + g - self.z0, -1)
        if self.world > 1 and self.rank == 0:
            far = (g >= self.qa) & (g < self.qa + self.rp) & (loc < 0)
            loc = np.where(far, g - self.qa, loc)
        return loc

    # ---- plane maps of mvf_rl_zz ------------------------------------------------------------------------------------
    def _computed(self):
        """Extended local positions t in [-rp, nzl+rp) whose ratio this rank computes itself (inside the volume)."""
        t = np.arange(-self.rp, self.nzl + self.rp)
        g = self.z0 + t
        return t, (g >= 0) & (g < self.nz)

    def zmap1(self):
        rp = self.rp
        s = np.arange(-2 * rp, self.nzl + 2 * rp)
        loc = self.est_local(_ext(self.z0 + s, self.nz, self.K))
        t, comp = self._computed()
        needed = np.zeros(len(s), dtype=bool)
        for tt in t[comp]:
            needed[tt + rp:tt + 3 * rp + 1] = True      # s in [t - rp, t + rp]  (index s + 2rp)
        if np.any(needed & (loc < 0)):
            raise ValueError("slab layout does not cover the planes the first blur needs")
        return np.where(loc < 0, self.front, loc).astype(np.int32)

    def rmap(self):
        t, comp = self._computed()
        g = self.z0 + t
        q = np.where(g < 0, self.nz - 3 - self.K - g, 2 * self.nz - 2 - g)     # reference boundary map, unclipped
        side = -1 - np.clip(q - self.qa, 0, self.K)
        loc = self.view_local(g)
        if np.any(comp & (loc < 0)):
            raise ValueError("slab layout does not cover the view planes the ratio needs")
        return np.where(comp, loc, side).astype(np.int32)

    def side_jobs(self):
        """[(zmap [cnt + 2rp], first local view plane, cnt, first side plane)]: ratio planes this rank computes up front
        with mvf_rl_ratio_z because positions outside the volume repeat them (low edge: wrap, rank 0; top: reflect)."""
        rp, jobs = self.rp, []
        ranges = []
        if self.world == 1:
            ranges.append((self.qa, self.K + 1))
        else:
            if self.rank == 0:
                ranges.append((self.qa, rp))
            if self.rank == self.world - 1:
                ranges.append((self.nz - 1 - rp, rp))
        for q0, cnt in ranges:
            tt = np.arange(-rp, cnt + rp)
            loc = self.est_local(_ext(q0 + tt, self.nz, self.K))
            if np.any(loc < 0):
                raise ValueError("slab layout does not cover the planes of the boundary ratios")
            v0 = self.view_local(np.arange(q0, q0 + cnt))
            if np.any(v0 < 0) or np.any(np.diff(v0) != 1):
                raise ValueError("slab layout does not hold the view planes of the boundary ratios")
            jobs.append((loc.astype(np.int32), int(v0[0]), int(cnt), int(q0 - self.qa)))
        return jobs

    # ---- plane transfers: (src_rank, first src local plane, count, first dst local plane) --------------------------------
    def _peer(self, r):
        return ZZLayout(self.nz, self.world, r, self.K, self.rp, self.bounds)

    def est_transfers(self):
        """Halo fills of the padded estimate, every iteration (sources are OWN planes of the peers)."""
        out = []
        if self.world == 1:
            return out
        rp2 = self.back
        if self.rank > 0:
            p = self._peer(self.rank - 1)
            out.append((p.rank, p.front + p.nzl - rp2, rp2, self.front - rp2))
        if self.rank + 1 < self.world:
            p = self._peer(self.rank + 1)
            n = min(rp2, p.nzl)
            out.append((p.rank, p.front, n, self.front + self.nzl))
        if self.rank == 0:
            p = self._peer(self.world - 1)
            n = min(self.front, self.nz - self.far0)
            out.append((p.rank, p.front + self.far0 - p.z0, n, 0))
        return out

    def view_transfers(self):
        """Static halo fills of the padded (masked) views, once per plan."""
        out = []
        if self.world == 1:
            return out
        rp = self.rp
        if self.rank > 0:
            p = self._peer(self.rank - 1)
            out.append((p.rank, rp + p.nzl - rp, rp, 0))
        if self.rank + 1 < self.world:
            p = self._peer(self.rank + 1)
            out.append((p.rank, rp, rp, rp + self.nzl))
        if self.rank == 0:
            p = self._peer(self.world - 1)
            out.append((p.rank, rp + self.qa - p.z0, rp, 0))
        return out


def run_transfers_local(bufs, layouts, which):
    """Plane transfers between the buffers of ranks emulated in ONE process (tests): bufs[r] belongs to rank r."""
    for r, lay in enumerate(layouts):
        for src, s0, n, d0 in getattr(lay, which)():
            bufs[r][d0:d0 + n] = bufs[src][s0:s0 + n]


def run_transfers_dist(buf, layout, which, group=None):
    """The same plane transfers over torch.distributed point-to-point (NCCL on the box, gloo in the CPU tests)."""
    import torch.distributed as dist
    if layout.world == 1:
        return
    staged = buf.is_cuda and dist.get_backend(group) == "gloo"
    ops, hosts = [], []
    for r in range(layout.world):      # what the others pull from this rank
        if r == layout.rank:
            continue
        for src, s0, n, d0 in getattr(layout._peer(r), which)():
            if src == layout.rank:
                t = buf[s0:s0 + n].contiguous()
                ops.append(dist.P2POp(dist.isend, t.cpu() if staged else t, r, group=group))
    for src, s0, n, d0 in getattr(layout, which)():
        dst = buf[d0:d0 + n]
        if staged:
            host = torch.empty(dst.shape, dtype=dst.dtype)
            hosts.append((host, dst))
            ops.append(dist.P2POp(dist.irecv, host, src, group=group))
        else:
            ops.append(dist.P2POp(dist.irecv, dst, src, group=group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for host, dst in hosts:
        dst.copy_(host)


class SymmetricPull:
    """Padded estimate volumes of the v4 schedule in torch symmetric memory: every rank copies its halo planes straight
    out of the peers' buffers over NVLink after a device-side barrier (no NCCL call, CUDA-graph capturable)."""

    def __init__(self, layout, ny, nx, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = dist.group.WORLD if group is None else group
        self.layout = layout
        self.shape = (layout.height_max, int(ny), int(nx))
        self.store = symm.empty(self.shape, dtype=torch.float32, device=device)
        self.store.zero_()
        self.hdl = symm.rendezvous(self.store, group)
        self.buf = self.store[:layout.height]
        self.pulls = [(self.hdl.get_buffer(src, self.shape, torch.float32), s0, n, d0)
                      for src, s0, n, d0 in layout.est_transfers()]

    def barrier(self):
        self.hdl.barrier(channel=0)

    def pull(self):
        for peer, s0, n, d0 in self.pulls:
            self.buf[d0:d0 + n].copy_(peer[s0:s0 + n])
