"""The `cfg4` / `cfg5` objects of bench.py alone (configs 4 and 5 on N GPUs), without the config-2 part of the line:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      profiles/scripts/big_only.py cfg4 [cfg5]

Same code path as `bench.py --gpus 8` takes for these objects (bench.cfg4_object); prints one JSON line on rank 0.
"""
import json
import os
import sys
from types import SimpleNamespace

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)


def main():
    import torch
    import torch.distributed as dist
    import bench
    from siga_b200 import api, formats
    cfgs = sys.argv[1:] or ["cfg4"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def red(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    args = SimpleNamespace(cfg4_scale=1.0)
    ctx = api.Context(local)
    line = {"n_gpus": world}
    for cfg in cfgs:
        line[cfg] = bench.cfg4_object(args, ctx, api, formats, torch, dist, rank, world, local, barrier,
                                      lambda x: red(x, dist.ReduceOp.MAX), lambda x: red(x, dist.ReduceOp.SUM), cfg=cfg)
    ctx.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
