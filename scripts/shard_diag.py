"""Per-rank view of the sharded update: pairs, contacts, own step time and stage times of every rank (bench.py prints rank 0 only).
torchrun --nproc-per-node N scripts/shard_diag.py"""
import ctypes as C, json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ncollide_b200 import _ffi
from ncollide_b200.parallel import ShardedWorld
from ncollide_b200.scenes import config_scene
from ncollide_b200.world import Context
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
ctx = Context(local)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
ctx.lib.ncb_set_stream(ctx.h, C.c_void_p(stream.cuda_stream))
scene = config_scene(3, 1_000_000 * world)
ctx.set_scene(scene); ctx.synchronize()
sw = ShardedWorld(ctx, scene, world, rank, dev)
cc = _ffi.UpdateCountsC()
for _ in range(4): counts = sw.step(cc)
dist.barrier(); torch.cuda.synchronize()
ms = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); counts = sw.step(cc); e1.record(stream); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
ctx.profile_enable(True)
acc = {}
for _ in range(3):
    sw.step(cc)
    for name, t, _l in ctx.profile_get(): acc[name] = acc.get(name, 0) + t / 3
ctx.profile_enable(False)
row = {"rank": rank, "mode": sw.mode, "pairs": counts["n_pairs"], "contacts": counts["n_contacts"], "epa_pairs": counts.get("n_epa_pairs"),
       "ms": round(sum(ms) / len(ms), 3), "stages": {k: round(v, 3) for k, v in acc.items()}}
rows = [None] * world
dist.all_gather_object(rows, row)
if rank == 0:
    for r in rows: print(json.dumps(r))
dist.destroy_process_group()
