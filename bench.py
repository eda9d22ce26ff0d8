#!/usr/bin/env python
"""Benchmark of the view-fusion hot path (BASELINE.json metric: fused Gvoxel/s, affine + weighted).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one synthetic batch: affine resampling of every view into the
fused frame + blending weights + normalised weighted sum (config[1] of BASELINE.json: 4 views 512^3 float32,
blending-weighted additive fusion).  Rank 0 prints ONE JSON line (contract in the task description):
  value      fused voxels / s with the views already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the host-array API: pinned host views -> H2D -> kernel -> D2H fused volume
  roofline   fused kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json), algorithmic bytes =
             sum_v N_v*b_in + N*2 per launch
  cpu_baseline  the oracle (numpy/scipy restatement of the reference CPU path) on a bounded sub-box
N > 1 (torchrun): every rank fuses its own time point of the same shape (the reference's unit of
parallel work is a file/time point, bin/mvregfus_run.py:229-231); no collective on the data path -> "weak".
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
# synthetic workload (SURVEY §8d recipe, generated on the device: smooth random field in an ellipsoid,
# observed by V views rotated about y with 1 % scale jitter and sub-voxel shifts)
# ---------------------------------------------------------------------------------------------------
def make_workload(device, seed, dtype="f32"):
    import torch
    import torch.nn.functional as F
    from mvregfus_b200 import fusion, synthetic
    from mvregfus_b200 import multiview as mv
    shape = VIEW_SHAPE
    g = torch.Generator(device=device)
    g.manual_seed(1000 + seed)
    fine = torch.rand((1, 1) + tuple(s // 16 for s in shape), generator=g, device=device)
    coarse = torch.rand((1, 1) + tuple(max(2, s // 128) for s in shape), generator=g, device=device)
    gt = F.interpolate(fine, size=shape, mode="trilinear", align_corners=True)
    gt = gt * F.interpolate(coarse, size=shape, mode="trilinear", align_corners=True)
    ax = [torch.linspace(-1, 1, s, device=device) for s in shape]
    r2 = (ax[0][:, None, None] / 0.9) ** 2 + (ax[1][None, :, None] / 0.9) ** 2 + (ax[2][None, None, :] / 0.9) ** 2
    gt = (gt[0, 0] * 4000.0 * (r2 <= 1.0)).contiguous()
    params = synthetic.view_params(shape, NVIEWS, seed=1000 + seed)
    props = [{"size": np.array(shape), "spacing": np.ones(3), "origin": np.zeros(3)} for _ in range(NVIEWS)]
    views = []
    for v in range(NVIEWS):
        A, c = params[v][:9].reshape(3, 3), params[v][9:]
        Ainv = np.linalg.inv(A)
        img = fusion.affine_resample(gt, Ainv, -Ainv @ c, shape, rule="scipy")  # view index -> ground-truth index
        if dtype == "u16":
            img = img.clamp_(0, 4095).round_().to(torch.int32).to(torch.uint16)
        views.append(img)
    del gt, fine, coarse, r2
    sp = mv.calc_stack_properties_from_views_and_params(props, params, spacing=[1.0, 1.0, 1.0])
    return views, params, props, sp


RL_SHAPE = (512, 1024, 1024)
RL_SIGMA = (6.0, 2.0)  # sz, sxy  -> 19^3 PSF at spacing 1 (BASELINE.json configs[3])
RL_ITERS = 10


def run_rl(dev, rank=0, world=1):
    """BASELINE.json configs[3]: 4 views 1024x1024x512 uint16, Gaussian PSF sigma 2x2x6, 10 RL iterations, on the
    transformed views and blending weights produced by the fusion kernels.  Returns the `rl` JSON object."""
    import torch
    import torch.nn.functional as F
    from mvregfus_b200 import _cabi, fusion, rl, synthetic
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
    if world == 1:
        plan = rl.RLPlan(tviews, [w[v] for v in range(NVIEWS)], psfs)
    else:  # z-slab of this rank; PSF-radius halos are exchanged over NCCL inside plan.iterate()
        import torch.distributed as dist
        from mvregfus_b200 import slab
        z0, z1 = slab.slab_bounds(shape[0], world)[rank]
        tv_slab = [t[z0:z1].clone() for t in tviews]
        w_slab = [w[v][z0:z1].clone() for v in range(NVIEWS)]
        del tviews, w
        torch.cuda.empty_cache()
        plan = rl.SlabRLPlan(tv_slab, w_slab, psfs, shape[0], rank, world)
    plan.init_estimate()
    plan.iterate()  # warm-up
    torch.cuda.synchronize()
    plan.init_estimate()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _cabi.launch_count(reset=True)
    a.record()
    for _ in range(RL_ITERS):
        plan.iterate()
    b.record()
    torch.cuda.synchronize()
    rl_launches = _cabi.launch_count()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    nvox = int(np.prod(shape))
    peak, how = measured_peak_gbs()
    alg = (12 + NVIEWS * (16 + 2)) * nvox
    per_it = ms / RL_ITERS
    return {"metric": "rl_iterations_per_s", "value": 1e3 / per_it, "unit": "it/s", "ms_per_iteration": per_it,
            "config": {"workload": "BASELINE.json configs[3]: %d views %dx%dx%d uint16 in the fused frame, Gaussian PSF "
                                   "sigma z=%g xy=%g (19^3, separable), %d iterations" % (NVIEWS, *shape, *RL_SIGMA, RL_ITERS)},
            "kernel_launches_per_iteration": rl_launches // RL_ITERS,
            "kernels": "per view and blur: rl_xy_kernel (x/y passes of every plane) + rl_xy_edge_kernel + rl_z_kernel "
                       "(z pass + fused ratio / correction / update epilogue)" if getattr(plan, "streamed", [False])[0]
                       else "rl_blur_kernel (fused x/y/z + epilogue)",
            "n_gpus": world,
            "sharding": "whole volume" if world == 1 else "z-slabs, ring halo exchange (%d planes) over NCCL before every blur, strong scaling" % ((19 - 1) // 2),
            "roofline": {"bound": "hbm", "achieved": alg / (per_it * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                         "frac": alg / (per_it * 1e-3) / 1e9 / (peak * world), "algorithmic_bytes_per_iteration": alg,
                         "peak_source": how}}


def cpu_baseline(views_host, params, props, sp, box=96, threads=1):
    """Oracle (= the reference's numpy/scipy arithmetic) on a centred sub-box of the same workload,
    chunked like the reference's transform_chunk / fuse_block.  Returns (Gvoxel/s, seconds, description)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import mvregfus_oracle as O
    dims = np.array([int(s) for s in sp["size"]])
    lo = (dims - box) // 2
    chunk = 32 if threads > 1 else box
    locs = [(z, y, x) for z in range(0, box, chunk) for y in range(0, box, chunk) for x in range(0, box, chunk)]
    maps = [O.index_space_affine(params[v], props[v]["spacing"], props[v]["origin"], sp["spacing"], sp["origin"])
            for v in range(len(views_host))]

    def one(loc):
        off = lo + np.array(loc)
        bprops = {"size": np.array([chunk] * 3), "spacing": sp["spacing"], "origin": sp["origin"] + off * sp["spacing"]}
        tv = [O.transform_chunk(views_host[v], maps[v][0], maps[v][1], off, [chunk] * 3) for v in range(len(views_host))]
        w = O.get_weights_simple(props, params, bprops)
        return O.fuse_views_weights(tv, params, bprops, weights=w)

    t0 = time.perf_counter()
    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, locs))
    else:
        for loc in locs:
            one(loc)
    dt = time.perf_counter() - t0
    return box ** 3 / dt / 1e9, dt, "centred %d^3 sub-box of the fused frame, %d^3 chunks, all %d views" % (box, chunk, len(views_host))


def run_reference(args):
    """--impl reference: the reference's CPU arithmetic (oracle port; the reference itself cannot be
    imported on the box) on the host cores, same config/metric, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from mvregfus_b200 import multiview as mv, synthetic
    cores = len(os.sched_getaffinity(0))
    # the same generator as the GPU arm, on a GPU if there is one (input generation is not timed)
    views, params, props, sp = make_workload(torch.device("cuda:0"), 0)
    views_host = [v.cpu().numpy() for v in views]
    times, box = [], 96
    for i in range(args.warmup + args.steps):
        gv, dt, sample = cpu_baseline(views_host, params, props, sp, box=box, threads=cores)
        if i >= args.warmup:
            times.append(dt)
    ms = float(np.mean(times) * 1e3)
    val = box ** 3 / (ms * 1e-3) / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Gvoxel/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(sp),
            "cpu_baseline": {"value": val, "unit": "Gvoxel/s", "cores": cores, "kind": "port", "sample": sample},
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


def workload_config(sp):
    return {"workload": "BASELINE.json configs[1]: %d views %dx%dx%d float32, affine resampling + blending-weighted "
                        "additive fusion, fused frame %s at spacing 1" % (NVIEWS, *VIEW_SHAPE, "x".join(str(int(s)) for s in sp["size"])),
            "views": NVIEWS, "view_shape": list(VIEW_SHAPE), "fused_shape": [int(s) for s in sp["size"]],
            "weights": "blending", "l2": "inputs (2.1 GB) larger than L2 (126 MB); no explicit flush"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from mvregfus_b200 import _cabi, fusion
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.load()
    views, params, props, sp = make_workload(dev, rank)
    dims = [int(s) for s in sp["size"]]
    nvox = int(np.prod(dims))
    vs = fusion.build_viewset(views, props, params, sp, weights="blending")
    out = torch.empty(dims, dtype=torch.uint16, device=dev)

    def step():
        fusion.fuse_blend(vs, dims, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    _cabi.launch_count(reset=True)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        start.record()
        for a, b in evs:
            a.record()
            step()
            b.record()
        stop.record()
        barrier()
        total_ms = start.elapsed_time(stop)
        if total_ms < 400:  # keep the sampler alive long enough to see clocks under load
            t_end = time.time() + 0.5
            while time.time() < t_end:
                step()
            torch.cuda.synchronize()
    launches = _cabi.launch_count()
    launches = args.steps  # launches inside the timed region (one fused kernel per step)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * nvox / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through host arrays: pinned views -> H2D -> fused kernel -> D2H -------------------
    # Every step copies its own inputs in and its result out; the copies of consecutive steps are pipelined on
    # three streams over two sets of device buffers (PCIe is full duplex: H2D of step i+1 runs under the kernel
    # and the D2H of step i).
    with numa_local(local):   # first-touch the pinned buffers on the NUMA node of this rank's GPU (matters at N > 1)
        pinned = [torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for v in views]
        for p in pinned:
            p.zero_()
    for p, v in zip(pinned, views):
        p.copy_(v)
    with numa_local(local):
        out_host = [torch.empty(dims, dtype=torch.uint16, pin_memory=True) for _ in range(2)]
        for o in out_host:
            o.zero_()
    dsets = [[torch.empty_like(v) for v in views] for _ in range(2)]
    outs = [torch.empty(dims, dtype=torch.uint16, device=dev) for _ in range(2)]
    vsets = [fusion.build_viewset(d, props, params, sp, weights="blending") for d in dsets]
    dviews, vs2 = dsets, vsets  # kept for the cleanup below
    h2d = sum(p.numel() * p.element_size() for p in pinned)
    d2h = out_host[0].numel() * out_host[0].element_size()
    s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_k = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(nsteps):
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
                fusion.fuse_blend(vsets[b], dims, out=outs[b])
                ev_k[b].record(s_k)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_k[b])
                out_host[b].copy_(outs[b], non_blocking=True)
                ev_out[b].record(s_out)
        torch.cuda.current_stream().wait_stream(s_in)
        torch.cuda.current_stream().wait_stream(s_k)
        torch.cuda.current_stream().wait_stream(s_out)

    e2e_run(2)
    barrier()
    e_steps = max(4, min(args.steps, 8))
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    e2e_run(e_steps)
    e2.record()
    barrier()
    t = torch.tensor([s2.elapsed_time(e2)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * nvox / (float(t.item()) / e_steps * 1e-3) / 1e9

    if rank == 0:
        peak, how = measured_peak_gbs()
        b_in = sum(v.numel() * v.element_size() for v in views)
        alg_bytes = b_in + nvox * 2
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "fuse_blend_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        line = {"metric": METRIC, "value": value, "unit": "Gvoxel/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(sp),
                "clocks": clk.summary(), "gpu_launches": launches,
                "e2e": {"value": e2e_val, "unit": "Gvoxel/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": how,
                             "kernel": "fuse_blend_kernel<float,4>", "kernel_ms": kernel_ms,
                             "algorithmic_bytes": alg_bytes}}
        if not args.skip_rl:
            del views, dviews, vs, vs2, out
            torch.cuda.empty_cache()
            line["rl"] = run_rl(dev, rank, world)
        if not args.no_cpu_baseline:
            views_host = [p.numpy() for p in pinned]
            gv, dt, sample = cpu_baseline(views_host, params, props, sp, box=192, threads=1)
            line["cpu_baseline"] = {"value": gv, "unit": "Gvoxel/s", "cores": 1, "kind": "port", "sample": sample,
                                    "seconds": dt}
        print(json.dumps(line))
    if rank != 0 and world > 1 and not args.skip_rl:
        del views, dviews, vs, vs2, out
        torch.cuda.empty_cache()
        run_rl(dev, rank, world)
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
