"""Multi-GPU parity check, launched by hand under torchrun on the GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_check.py

Runs spectralelementmethod_b200.distributed_check.run (the same self-check bench.py runs
before timing a multi-GPU line) and prints the margins.  Not collected by pytest (needs
N GPUs)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from spectralelementmethod_b200 import distributed_check  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    res = distributed_check.run(rank, world, dev)
    if rank == 0:
        print("multigpu_check ok: " + json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
