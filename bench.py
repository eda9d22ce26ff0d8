#!/usr/bin/env python
"""Benchmark of the view-fusion hot path (BASELINE.json metric: fused Gvoxel/s, affine + weighted; RL iterations/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one synthetic batch: affine resampling of every view into the fused
frame + blending weights + normalised weighted sum (BASELINE.json configs[1]: 4 views 512^3 float32, blending-weighted
additive fusion).  Rank 0 prints ONE JSON line (contract in the task description):
  value        fused voxels / s with the views already resident in HBM (CUDA events, max over ranks)
  e2e          the same metric through the drop-in entry points of mvregfus_b200.multiview with HOST numpy buffers:
               transform_stack_dask_numpy per view + fuse_blockwise_arrays (the reference's fuse_blockwise socket,
               mv:3637-3916), host->device and device->host copies inside the timed region (the socket streams 128-plane
               rows: source bricks up, kernels, result rows down; h2d_bytes_per_step counts what was actually sent, summed
               over the ranks)
  e2e_device_api  pipelined device API (pinned views -> H2D -> fused kernel -> D2H), for comparison
  roofline     fused kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json); algorithmic bytes =
               sum_v N_v*b_in + N*2 per launch
  rl           BASELINE.json configs[3]: multi-view Richardson-Lucy iterations/s (z-slabs + halo exchange at N > 1,
               checked against the whole-volume run)
  cpu_baseline the oracle (numpy/scipy restatement of the reference CPU path) on a bounded sample, 128^3 chunks
N > 1 (torchrun): ONE fused volume, partitioned into z-slabs across the ranks (north_star); no collective on the fusion
data path; `value` = voxels of the whole volume / slowest rank -> "strong" scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VIEW_SHAPE = (512, 512, 512)
NVIEWS = 4
METRIC = "fused_gvoxel_per_s"
SEED = 1000
CHUNK = 128   # the reference's fusion chunk size (mvregfus/mv_graph.py:73)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# synthetic workload (SURVEY §8d recipe: smooth random field in an ellipsoid, observed by V views rotated about y
# with 1 % scale jitter and sub-voxel shifts).  The geometry (params, frames) is the same in both arms; the GPU arm
# renders the voxels on the device, the reference arm on the host (numpy/scipy only).
# ---------------------------------------------------------------------------------------------------
def workload_geometry():
    from mvregfus_b200 import mv_utils, synthetic
    params = synthetic.view_params(VIEW_SHAPE, NVIEWS, seed=SEED)
    props = [{"size": np.array(VIEW_SHAPE), "spacing": np.ones(3), "origin": np.zeros(3)} for _ in range(NVIEWS)]
    # fused frame = bounding box of all views (calc_stack_properties_from_views_and_params, mv:4916-4932), host float64
    lo, hi = [], []
    unit = np.array(list(np.ndindex(2, 2, 2)))
    for sp, p in zip(props, params):
        corners = unit * sp["size"] * sp["spacing"] + sp["origin"]
        Minv = mv_utils.params_to_matrix(mv_utils.invert_params(p))
        pts = corners @ Minv[:3, :3].T + Minv[:3, 3]
        lo.append(pts.min(0))
        hi.append(pts.max(0))
    lo, hi = np.min(lo, 0), np.max(hi, 0)
    spacing = np.ones(3)
    sp = {"size": np.ceil((hi - lo) / spacing).astype(np.uint16), "spacing": spacing, "origin": lo}
    return params, props, sp


def make_workload(device, seed=0, dtype="f32"):
    """Views rendered on the device (input generation is not part of any timed region)."""
    import torch
    import torch.nn.functional as F
    from mvregfus_b200 import fusion
    shape = VIEW_SHAPE
    g = torch.Generator(device=device)
    g.manual_seed(SEED + seed)
    fine = torch.rand((1, 1) + tuple(s // 16 for s in shape), generator=g, device=device)
    coarse = torch.rand((1, 1) + tuple(max(2, s // 128) for s in shape), generator=g, device=device)
    gt = F.interpolate(fine, size=shape, mode="trilinear", align_corners=True)
    gt = gt * F.interpolate(coarse, size=shape, mode="trilinear", align_corners=True)
    ax = [torch.linspace(-1, 1, s, device=device) for s in shape]
    r2 = (ax[0][:, None, None] / 0.9) ** 2 + (ax[1][None, :, None] / 0.9) ** 2 + (ax[2][None, None, :] / 0.9) ** 2
    gt = (gt[0, 0] * 4000.0 * (r2 <= 1.0)).contiguous()
    params, props, sp = workload_geometry()
    views = []
    for v in range(NVIEWS):
        A, c = params[v][:9].reshape(3, 3), params[v][9:]
        Ainv = np.linalg.inv(A)
        img = fusion.affine_resample(gt, Ainv, -Ainv @ c, shape, rule="scipy")  # view index -> ground-truth index
        if dtype == "u16":
            img = img.clamp_(0, 4095).round_().to(torch.int32).to(torch.uint16)
        views.append(img)
    del gt, fine, coarse, r2
    return views, params, props, sp


def host_views_for_box(params, props, sp, lo, box):
    """Host-rendered views (numpy/scipy only) holding real voxels wherever the fused sub-box [lo, lo+box) can reach them
    (bounding box of the sub-box in every view, +-4 voxels) and zeros elsewhere - the oracle's transform_chunk only
    reads the bounding box of its chunk (mv:1552-1571), so fusing the sub-box gives what the full views would give."""
    from scipy import ndimage
    from mvregfus_b200 import mv_utils, synthetic
    gt = synthetic.ground_truth(VIEW_SHAPE, seed=SEED)
    corners = (np.array(list(np.ndindex(2, 2, 2))) * box + lo).astype(np.float64)
    views = []
    for v in range(NVIEWS):
        M, off = mv_utils.index_space_affine(params[v], props[v]["spacing"], props[v]["origin"], sp["spacing"], sp["origin"])
        pts = corners @ M.T + off
        a = np.clip(np.floor(pts.min(0)).astype(int) - 4, 0, np.array(VIEW_SHAPE))
        b = np.clip(np.ceil(pts.max(0)).astype(int) + 5, 0, np.array(VIEW_SHAPE))
        A, c = params[v][:9].reshape(3, 3), params[v][9:]
        Ainv = np.linalg.inv(A)
        img = np.zeros(VIEW_SHAPE, dtype=np.float32)
        if np.all(b > a):   # view index y -> ground-truth index A^-1 (y - c), evaluated on the region [a, b) only
            img[a[0]:b[0], a[1]:b[1], a[2]:b[2]] = ndimage.affine_transform(
                gt, Ainv, Ainv @ (a - c), output_shape=tuple(b - a), order=1).astype(np.float32)
        views.append(img)
    return views


def oracle_pass(views_host, params, props, sp, box, threads):
    """One pass of the reference's CPU arithmetic (oracle) over a centred box^3 sub-box of the fused frame in the
    reference's own 128^3 chunks (transform_chunk -> get_weights_simple -> fuse_views_weights per chunk, as fuse_block
    does), chunks dispatched over `threads` host threads like dask's threaded scheduler (mv:3797).  Returns seconds."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import mvregfus_oracle as O
    dims = np.array([int(s) for s in sp["size"]])
    lo = (dims - box) // 2
    locs = [(z, y, x) for z in range(0, box, CHUNK) for y in range(0, box, CHUNK) for x in range(0, box, CHUNK)]
    maps = [O.index_space_affine(params[v], props[v]["spacing"], props[v]["origin"], sp["spacing"], sp["origin"])
            for v in range(len(views_host))]

    def one(loc):
        off = lo + np.array(loc)
        bprops = {"size": np.array([CHUNK] * 3), "spacing": sp["spacing"], "origin": sp["origin"] + off * sp["spacing"]}
        tv = [O.transform_chunk(views_host[v], maps[v][0], maps[v][1], off, [CHUNK] * 3) for v in range(len(views_host))]
        w = O.get_weights_simple(props, params, bprops)
        return O.fuse_views_weights(tv, params, bprops, weights=w)

    t0 = time.perf_counter()
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, locs))
    else:
        for loc in locs:
            one(loc)
    return time.perf_counter() - t0


def sample_text(box, threads):
    return ("centred %d^3 sub-box of the fused frame in %d^3 chunks (%d chunks), all %d views, %d host thread%s"
            % (box, CHUNK, (box // CHUNK) ** 3, NVIEWS, threads, "" if threads == 1 else "s"))


def workload_config(sp, world=1):
    return {"workload": "BASELINE.json configs[1]: %d views %dx%dx%d float32, affine resampling + blending-weighted "
                        "additive fusion, fused frame %s at spacing 1" % (NVIEWS, *VIEW_SHAPE, "x".join(str(int(s)) for s in sp["size"])),
            "views": NVIEWS, "view_shape": list(VIEW_SHAPE), "fused_shape": [int(s) for s in sp["size"]],
            "weights": "blending",
            "sharding": "whole volume on one GPU" if world == 1 else "one fused volume in %d z-slabs, one per GPU, no collective" % world,
            "l2": "inputs (2.1 GB) larger than L2 (126 MB); no explicit flush"}


def run_reference(args):
    """--impl reference: the reference's CPU arithmetic (oracle port; the reference itself cannot be imported on the box,
    SURVEY.md §8c) on all host cores, same config/metric, bounded sample per step.  numpy/scipy only: no GPU, no repo .so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    params, props, sp = workload_geometry()
    box = 256
    dims = np.array([int(s) for s in sp["size"]])
    views_host = host_views_for_box(params, props, sp, (dims - box) // 2, box)
    times = []
    for i in range(args.warmup + args.steps):
        dt = oracle_pass(views_host, params, props, sp, box, cores)
        if i >= args.warmup:
            times.append(dt)
    ms = float(np.mean(times) * 1e3)
    val = box ** 3 / (ms * 1e-3) / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Gvoxel/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(sp, args.gpus),
            "cpu_baseline": {"value": val, "unit": "Gvoxel/s", "cores": cores, "kind": "port", "sample": sample_text(box, cores)},
            "e2e": {"value": val, "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


class numa_local:
    """Context: run on the CPUs of the NUMA node the given GPU hangs off (so freshly allocated host memory is
    first-touched there), then restore the affinity.  Silently does nothing when the topology cannot be read."""

    def __init__(self, gpu_index):
        self.gpu, self.saved = gpu_index, None

    def __enter__(self):
        try:
            import torch
            pr = torch.cuda.get_device_properties(self.gpu)
            bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
                node = int(fh.read().strip())
            if node < 0:
                return self
            with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
                cpus = set()
                for part in fh.read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            allowed = os.sched_getaffinity(0)
            if cpus & allowed:
                self.saved = allowed
                os.sched_setaffinity(0, cpus & allowed)
        except Exception:
            self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass
        return False


# ---------------------------------------------------------------------------------------------------
# Richardson-Lucy, BASELINE.json configs[3]
# ---------------------------------------------------------------------------------------------------
RL_SHAPE = (512, 1024, 1024)
RL_SIGMA = (6.0, 2.0)  # sz, sxy  -> 19^3 PSF at spacing 1 (BASELINE.json configs[3])
RL_ITERS = 10


def rl_inputs(dev):
    """Transformed views, blending weights and PSFs of configs[3], rendered on the device."""
    import torch
    import torch.nn.functional as F
    from mvregfus_b200 import fusion, synthetic
    from mvregfus_b200 import multiview as mv
    shape = RL_SHAPE
    g = torch.Generator(device=dev)
    g.manual_seed(4004)
    fine = torch.rand((1, 1) + tuple(s // 16 for s in shape), generator=g, device=dev)
    gt = F.interpolate(fine, size=shape, mode="trilinear", align_corners=True)[0, 0]
    ax = [torch.linspace(-1, 1, s, device=dev) for s in shape]
    r2 = (ax[0][:, None, None] / 0.9) ** 2 + (ax[1][None, :, None] / 0.9) ** 2 + (ax[2][None, None, :] / 0.9) ** 2
    gt = (gt * 3000.0 * (r2 <= 1.0)).contiguous()
    del fine, r2
    params = synthetic.view_params(shape, NVIEWS, seed=4004)
    props = [{"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)} for _ in range(NVIEWS)]
    sp = {"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)}
    tviews = []
    for v in range(NVIEWS):
        A, c = params[v][:9].reshape(3, 3), params[v][9:]
        Ainv = np.linalg.inv(A)
        view = fusion.affine_resample(gt, Ainv, -Ainv @ c, shape, rule="scipy")          # the raw view
        view = view.clamp_(0, 4095).round_().to(torch.int32).to(torch.uint16)
        M, off = mv.index_space_affine(params[v], props[v]["spacing"], props[v]["origin"], sp["spacing"], sp["origin"])
        tviews.append(fusion.affine_resample(view, M, off, shape, rule="scipy"))             # transformed view
        del view
    del gt
    vs = fusion.build_viewset(None, props, params, sp, weights="blending")
    w = fusion.blend_weights(vs, shape, normalise=True)
    psfs = [mv.get_psf(params[v], sp, RL_SIGMA[0], RL_SIGMA[1]) for v in range(NVIEWS)]
    return tviews, w, psfs


def run_rl(dev, rank=0, world=1):
    """10 RL iterations on configs[3]; z-slabs with halo exchange at world > 1, and then the slab estimates are gathered
    and compared with a whole-volume run on rank 0 (`parity_vs_whole`).  Returns the `rl` JSON object."""
    import torch
    from mvregfus_b200 import _cabi, rl
    shape = RL_SHAPE
    tviews, w, psfs = rl_inputs(dev)
    if world == 1:
        plan = rl.RLPlan(tviews, [w[v] for v in range(NVIEWS)], psfs)
    else:  # z-slab of this rank; PSF-radius halos travel over NVLink inside plan.iterate()
        import torch.distributed as dist
        from mvregfus_b200 import slab
        # slabs balanced by work: rank 0 also blurs the planes its wrap-around refers to and the boundary ratios
        kz = int(np.asarray(psfs[0]).shape[0])
        bounds = slab.rl_slab_bounds(shape[0], world, kz, (int(_cabi.load().mvf_rl_padded_size(kz)) - 1) // 2)
        z0, z1 = bounds[rank]
        plan = rl.SlabRLPlan([t[z0:z1].clone() for t in tviews], [w[v][z0:z1].clone() for v in range(NVIEWS)], psfs,
                             shape[0], rank, world, bounds=bounds)
    halo_txt = ("pulled from the neighbours' symmetric-memory buffers, iteration replayed from a CUDA graph"
                if getattr(plan, "symm", None) is not None else "NCCL ring isend/irecv")
    plan.init_estimate()
    _cabi.launch_count(reset=True)
    plan.iterate()       # warm-up 1 runs eagerly (its launches are counted), warm-up 2 captures the CUDA graph of an iteration
    launches_per_iteration = _cabi.launch_count()
    plan.iterate()
    torch.cuda.synchronize()
    plan.init_estimate()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(RL_ITERS):
        plan.iterate()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    graphed = getattr(plan, "_graph_state", 0) == 2   # the timed iterations were replays of a captured CUDA graph
    parity = None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # parity of the sharded run: gather the slabs on rank 0, compare with the whole volume computed there
        est = plan.est.contiguous()
        if rank == 0:
            whole = torch.empty(shape, dtype=torch.float32, device=dev)
            whole[bounds[0][0]:bounds[0][1]] = est
            for r in range(1, world):
                dist.recv(whole[bounds[r][0]:bounds[r][1]], src=r)
            del plan
            torch.cuda.empty_cache()
            ref = rl.RLPlan(tviews, [w[v] for v in range(NVIEWS)], psfs)
            ref.init_estimate()
            for _ in range(RL_ITERS):   # the slabs were re-initialised after their warm-up
                ref.iterate()
            torch.cuda.synchronize()
            if torch.equal(whole, ref.est):
                parity = "bit-exact"
            else:
                err = float(((whole - ref.est).abs() / ref.est.abs().clamp_min(1.0)).max().item())
                parity = "max rel err %.3e" % err
                if err > 1e-3:
                    raise SystemExit("z-slab RL differs from the whole-volume run: %s" % parity)
        else:
            dist.send(est, dst=0)
    nvox = int(np.prod(shape))
    peak, how = measured_peak_gbs()
    alg = (12 + NVIEWS * (16 + 2)) * nvox
    per_it = ms / RL_ITERS
    return {"metric": "rl_iterations_per_s", "value": 1e3 / per_it, "unit": "it/s", "ms_per_iteration": per_it,
            "config": {"workload": "BASELINE.json configs[3]: %d views %dx%dx%d uint16 in the fused frame, Gaussian PSF "
                                   "sigma z=%g xy=%g (19^3, separable), %d iterations" % (NVIEWS, *shape, *RL_SIGMA, RL_ITERS)},
            "kernel_launches_per_iteration": launches_per_iteration,
            "cuda_graph": graphed,
            "n_gpus": world, "scaling": "strong", "slab_planes": [b - a for a, b in bounds] if world > 1 else [shape[0]],
            "sharding": "whole volume" if world == 1 else "z-slabs, PSF-radius halos over NVLink (%s)" % halo_txt,
            "parity_vs_whole": parity,
            "roofline": {"bound": "hbm", "achieved": alg / (per_it * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                         "frac": alg / (per_it * 1e-3) / 1e9 / (peak * world), "algorithmic_bytes_per_iteration": alg,
                         "peak_source": how}}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from mvregfus_b200 import _cabi, fusion, slab
    from mvregfus_b200 import multiview as mv
    from mvregfus_b200.image_array import ImageArray
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.load()
    views, params, props, sp = make_workload(dev, 0)   # every rank renders the same volume; rank r fuses z-slab r
    full = [int(s) for s in sp["size"]]
    nvox = int(np.prod(full))
    z0, z1 = slab.slab_bounds(full[0], world)[rank]
    dims, origin = [z1 - z0, full[1], full[2]], (z0, 0, 0)
    vs = fusion.build_viewset(views, props, params, sp, weights="blending")
    out = torch.empty(dims, dtype=torch.uint16, device=dev)

    def step():
        fusion.fuse_blend(vs, dims, origin=origin, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        _cabi.launch_count(reset=True)
        start.record()
        for a, b in evs:
            a.record()
            step()
            b.record()
        stop.record()
        barrier()
        launches = _cabi.launch_count()   # kernels this library launched inside the timed region
        total_ms = start.elapsed_time(stop)
        if total_ms < 400:  # keep the sampler alive long enough to see clocks under load
            t_end = time.time() + 0.5
            while time.time() < t_end:
                step()
            torch.cuda.synchronize()
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = nvox / (ms_per_step * 1e-3) / 1e9

    # ---- the production dtype on the same geometry: uint16 views (device resident, kernel only) ----------
    u16_ms = None
    if world == 1:
        views16 = [v.clamp(0, 4095).round().to(torch.int32).to(torch.uint16) for v in views]
        vs16 = fusion.build_viewset(views16, props, params, sp, weights="blending")
        for _ in range(3):
            fusion.fuse_blend(vs16, dims, origin=origin, out=out)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fusion.fuse_blend(vs16, dims, origin=origin, out=out)
        b.record()
        torch.cuda.synchronize()
        u16_ms = a.elapsed_time(b) / 10
        del views16, vs16

    # ---- end to end through the drop-in entry points, host numpy buffers in and out ----------------------
    # The caller's arrays live in page-locked host memory (fusion.pinned_empty), the documented way to feed the
    # drop-in path; transform_stack_dask_numpy is lazy like the reference's, fuse_blockwise_arrays then fuses
    # straight from the raw views on the device (block semantics of fuse_block included) and returns a host array.
    sp_slab = {"size": np.array(dims), "spacing": sp["spacing"], "origin": sp["origin"] + np.array(origin) * sp["spacing"]}
    with numa_local(local):   # first-touch the pinned buffers on the NUMA node of this rank's GPU (matters at N > 1)
        host_views = []
        for v in views:
            h = fusion.pinned_empty(v.shape, np.float32)
            h[...] = 0
            host_views.append(h)
    for h, v in zip(host_views, views):
        torch.from_numpy(h).copy_(v)
    torch.cuda.synchronize()
    host_views = [ImageArray(h, spacing=p["spacing"], origin=p["origin"]) for h, p in zip(host_views, props)]

    def e2e_step():
        tv = [mv.transform_stack_dask_numpy(h, params[i], stack_properties=sp_slab) for i, h in enumerate(host_views)]
        return mv.fuse_blockwise_arrays(tv, params, sp_slab, props)

    fused_host = e2e_step()
    fusion.TRANSFER["h2d_bytes"] = 0
    fused_host = e2e_step()
    h2d = int(fusion.TRANSFER["h2d_bytes"])   # what the entry points actually sent: whole views at N = 1, the slab's source bricks beyond
    d2h = fused_host.nbytes
    barrier()
    e_steps = max(3, min(args.steps, 6))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        fused_host = e2e_step()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = nvox / (float(t.item()) / e_steps) / 1e9
    if world > 1:   # whole-job bytes per step: every rank's uploads and its slab of the result
        tb = torch.tensor([h2d, d2h], device=dev, dtype=torch.float64)
        dist.all_reduce(tb)
        h2d, d2h = int(tb[0].item()), int(tb[1].item())

    # ---- for comparison: the pipelined device API (pinned views -> H2D -> fused kernel -> D2H) ------------
    pinned = [torch.from_numpy(np.asarray(h)) for h in host_views]
    out_host = [torch.empty(dims, dtype=torch.uint16, pin_memory=True) for _ in range(2)]
    dsets = [[torch.empty_like(v) for v in views] for _ in range(2)]
    outs = [torch.empty(dims, dtype=torch.uint16, device=dev) for _ in range(2)]
    vsets = [fusion.build_viewset(d, props, params, sp, weights="blending") for d in dsets]
    s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_k = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def pipelined(nsteps):
        for i in range(nsteps):
            b = i & 1
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_k[b])       # the kernel of step i-2 has consumed this buffer set
                for p, d in zip(pinned, dsets[b]):
                    d.copy_(p, non_blocking=True)
                ev_in[b].record(s_in)
            with torch.cuda.stream(s_k):
                s_k.wait_event(ev_in[b])
                if i >= 2:
                    s_k.wait_event(ev_out[b])      # the result buffer of step i-2 has left the device
                fusion.fuse_blend(vsets[b], dims, origin=origin, out=outs[b])
                ev_k[b].record(s_k)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_k[b])
                out_host[b].copy_(outs[b], non_blocking=True)
                ev_out[b].record(s_out)
        for s in (s_in, s_k, s_out):
            torch.cuda.current_stream().wait_stream(s)

    pipelined(2)
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    pipelined(6)
    e2.record()
    barrier()
    t = torch.tensor([s2.elapsed_time(e2)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_api_val = nvox / (float(t.item()) / 6 * 1e-3) / 1e9
    del dsets, outs, vsets, out_host

    line = None
    if rank == 0:
        peak, how = measured_peak_gbs()
        b_in = sum(v.numel() * v.element_size() for v in views)
        alg_bytes = b_in + nvox * 2               # whole volume: the slabs of all ranks together
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9 if world > 1 else alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r02", "fuse_axis_traffic.json")
        if world == 1 and os.path.exists(tpath):
            with open(tpath) as fh:
                tj = json.load(fh)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        line = {"metric": METRIC, "value": value, "unit": "Gvoxel/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(sp, world),
                "clocks": clk.summary(), "gpu_launches": launches,
                "e2e": {"value": e2e_val, "unit": "Gvoxel/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "mvregfus_b200.multiview.transform_stack_dask_numpy x%d + fuse_blockwise_arrays, host numpy "
                               "buffers in page-locked memory (whole views; a z-slab rank uploads the source bricks of its "
                               "slab, the reference's per-chunk crop mv:1552-1601)" % NVIEWS},
                "e2e_device_api": {"value": dev_api_val, "unit": "Gvoxel/s",
                                   "api": "fusion.fuse_blend on device tensors, copies of consecutive steps pipelined on 3 streams"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s",
                             "frac": achieved / (peak * world), "traffic": traffic, "traffic_source": traffic_src,
                             "peak_source": how,
                             "kernel": "fuse_axis_kernel<float> (+ ax_tables_kernel)", "kernel_ms": kernel_ms,
                             "algorithmic_bytes": alg_bytes}}
        if u16_ms is not None:
            b16 = b_in // 2 + nvox * 2
            line["roofline_uint16_views"] = {"kernel_ms": u16_ms, "achieved": b16 / (u16_ms * 1e-3) / 1e9, "unit": "GB/s",
                                             "frac": b16 / (u16_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": b16,
                                             "gvoxel_per_s": nvox / (u16_ms * 1e-3) / 1e9}
    host_np = [np.asarray(h) for h in host_views]
    del views, vs, out, pinned
    torch.cuda.empty_cache()
    if not args.skip_rl:
        rl_obj = run_rl(dev, rank, world)
        if rank == 0:
            line["rl"] = rl_obj
    if world >= 2 and not args.skip_extra:   # the multi-GPU configurations of BASELINE.json, one z-slab per rank
        from tools import extra_configs
        torch.cuda.empty_cache()
        xviews, xparams, xprops = extra_configs.make_views(dev)
        c2 = extra_configs.config2(dev, rank, world, xviews, xparams, xprops)
        if rank == 0:
            line["configs2"] = c2
        if world == 8:
            c4 = extra_configs.config4(dev, rank, world, 10, xviews, xparams, xprops)
            if rank == 0:
                line["configs4"] = c4
        del xviews
        torch.cuda.empty_cache()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            box = 256
            dt = oracle_pass(host_np, params, props, sp, box, 1)
            line["cpu_baseline"] = {"value": box ** 3 / dt / 1e9, "unit": "Gvoxel/s", "cores": 1, "kind": "port",
                                    "sample": sample_text(box, 1), "seconds": dt}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-rl", action="store_true")
    ap.add_argument("--skip-extra", action="store_true", help="no configs[2] / configs[4] extra keys at N >= 2")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
