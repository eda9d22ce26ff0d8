"""Per-phase device times of one v4 RL iteration on z-slabs (configs[3]), every rank: torchrun --nproc-per-node N tools/rl_slab_phases.py
Phases: barrier + halo pull, x/y passes of the estimate (own planes | halo planes), then per view side / zz / xy+epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from mvregfus_b200 import rl, slab

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
tviews, w, psfs = bench.rl_inputs(dev)
bounds = slab.rl_slab_bounds(bench.RL_SHAPE[0], world, 19, 9) if os.environ.get("RL_BALANCED", "1") == "1" else slab.slab_bounds(bench.RL_SHAPE[0], world)
z0, z1 = bounds[rank]
plan = rl.SlabRLPlan([t[z0:z1].clone() for t in tviews], [w[v][z0:z1].clone() for v in range(len(tviews))], psfs, bench.RL_SHAPE[0], rank, world, bounds=bounds)
del tviews, w
plan.init_estimate()
for _ in range(3):
    plan.iterate()
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    plan.iterate()
b.record(); torch.cuda.synchronize()
total = a.elapsed_time(b) / 10

marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))

def phased():
    main = torch.cuda.current_stream()
    mark("start")
    plan.symm.barrier(); mark("barrier1")
    plan._ev_fork.record(main); plan.comm_stream.wait_event(plan._ev_fork)
    with torch.cuda.stream(plan.comm_stream):
        plan.symm.pull(); plan._ev_join.record(plan.comm_stream)
    plan.sums[1:].zero_()
    plan._zz_xy_estimate(lambda: (mark("xy_own"), main.wait_event(plan._ev_join), mark("join")))
    mark("xy_halo")
    for v in range(plan.n):
        plan._zz_view(v, (lambda: (mark("pre_barrier2"), plan.symm.barrier(), mark("barrier2"))) if v == plan.n - 1 else None)
        mark("view%d" % v)

rows = []
for it in range(5):
    marks.clear()
    dist.barrier(); torch.cuda.synchronize()
    phased()
    torch.cuda.synchronize()
    rows.append([marks[i][1].elapsed_time(marks[i + 1][1]) for i in range(len(marks) - 1)])
names = [m[0] for m in marks[1:]]
med = np.median(np.array(rows), 0)
out = "rank %d  graph-replayed iteration %.3f ms | eager phases: " % (rank, total) + "  ".join("%s %.3f" % (n, t) for n, t in zip(names, med)) + "  | sum %.3f" % med.sum()
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    print("\n".join(gathered))
dist.destroy_process_group()
