"""BASELINE.json configs[4] (the north-star configuration) end to end on N GPUs (torchrun): 8 views (512,1024,1024)
uint16 fused to (1024,2048,2048) at spacing 0.5 with blending weights, then 10 multi-view RL iterations.  The same
function feeds bench.py's `configs4` key at N = 8 (tools/extra_configs.py)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import extra_configs  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = extra_configs.config4(dev, rank, world, int(os.environ.get("C5_ITERS", "10")))
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
