"""Host-link check for the multi-GPU end-to-end figure: pinned D2H / H2D bandwidth of one rank alone vs all ranks at once.
torchrun --nproc-per-node N scripts/pcie_concurrency.py"""
import json, os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 100_000_000
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
def bw(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9
res = {}
for name, fn in (("d2h", lambda: host.copy_(dev, non_blocking=True)), ("h2d", lambda: dev.copy_(host, non_blocking=True))):
    alone = 0.0
    for r in range(world):          # one rank at a time
        dist.barrier()
        if r == rank: alone = bw(fn)
        dist.barrier()
    dist.barrier()
    together = bw(fn)               # all ranks at once
    t = torch.tensor([alone, together], device="cuda", dtype=torch.float64)
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    res[name] = {"alone_GBps_per_rank": [round(float(x[0]), 1) for x in g], "together_GBps_per_rank": [round(float(x[1]), 1) for x in g],
                 "together_aggregate_GBps": round(sum(float(x[1]) for x in g), 1)}
if rank == 0: print(json.dumps({"n_gpus": world, "bytes": nbytes, **res}))
dist.destroy_process_group()
