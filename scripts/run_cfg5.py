"""BASELINE.json configs[4] ("config 5"): N = 16,000,000 convex shapes (1/2 cuboids, 1/2 hulls) sharded
over the GPUs of one node, at FULL size.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/run_cfg5.py [n_total] [out.json]
Times the sharded fresh-world update (device time, max over ranks) and checks size-independent parity properties:
  * every rank's pairs are oriented (larger, smaller), unique, and no pair is reported by two ranks (checked on a
    spatial sub-cube against the oracle's brute-force pair set: exact set equality);
  * the contacts of a sample of rank 0's pairs equal the oracle narrow phase (feature ids exact, values in tolerance).
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ncollide_b200 import _ffi  # noqa: E402
from ncollide_b200.parallel import ShardedWorld  # noqa: E402
from ncollide_b200.scenes import config_scene  # noqa: E402
from ncollide_b200.world import Context  # noqa: E402

RTOL, ATOL = 1e-4, 1e-5


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
    out_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/cfg5.json"
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.time()
    scene = config_scene(5, n_total)
    t_scene = time.time() - t0
    ctx = Context(local)
    stream = torch.cuda.Stream(device=local)
    ctx.lib.ncb_set_stream(ctx.h, C.c_void_p(stream.cuda_stream))
    with torch.cuda.stream(stream):
        ctx.set_scene(scene)
        sw = ShardedWorld(ctx, scene, world, rank, torch.device("cuda", local))
        cc = _ffi.UpdateCountsC()
        for _ in range(3):
            counts = sw.step(cc)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        steps = 3
        torch.cuda.synchronize()
        ev[0].record(stream)
        for _ in range(steps):
            counts = sw.step(cc)
        ev[1].record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([ev[0].elapsed_time(ev[1]) / steps], device="cuda")
        tot = torch.tensor([counts["n_pairs"], counts["n_contacts"], counts["n_contact_pairs"], counts["epa_overflow"], counts["ref_panics"]],
                           device="cuda", dtype=torch.int64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        res = ctx.world_fetch(counts)
    pairs = res.pairs
    ok_oriented = bool(np.all(pairs[:, 0] > pairs[:, 1]))
    keys = pairs[:, 0].astype(np.uint64) << np.uint64(32) | pairs[:, 1].astype(np.uint64)
    ok_unique = len(np.unique(keys)) == len(keys)

    # ---- sub-cube check: the union of all ranks' pairs restricted to the objects of a cube == oracle brute force
    side = float(scene.pos.max())
    lo, hi = side / 2 - 7.0, side / 2 + 7.0
    inside = np.all((scene.pos >= lo) & (scene.pos < hi), axis=1)
    mine = pairs[inside[pairs[:, 0]] & inside[pairs[:, 1]]]
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, mine)
    else:
        gathered = [mine]
    report = None
    if rank == 0:
        from oracle.pyoracle import Oracle

        orc = Oracle()
        sub = np.nonzero(inside)[0]
        union = np.concatenate(gathered)
        ukeys = union[:, 0].astype(np.uint64) << np.uint64(32) | union[:, 1].astype(np.uint64)
        no_cross_rank_dup = len(np.unique(ukeys)) == len(ukeys)
        fat = orc.compute_aabbs(scene_subset(scene, sub))
        want = orc.broad_phase(fat, None, mode=2)
        want = np.stack([sub[want[:, 0]], sub[want[:, 1]]], axis=1).astype(np.uint64)
        wkeys = np.sort(np.maximum(want[:, 0], want[:, 1]) << np.uint64(32) | np.minimum(want[:, 0], want[:, 1]))
        sub_equal = bool(np.array_equal(np.sort(ukeys), wkeys))
        # ---- narrow-phase sample of rank 0's own pairs
        rng = np.random.default_rng(5)
        pick = np.sort(rng.choice(len(pairs), size=min(60_000, len(pairs)), replace=False))
        oc, ooff, oalgo, ostats = orc.narrow_phase(scene, pairs[pick])
        cnt_ok = np.array_equal(np.diff(ooff).astype(np.int64), res.manifold_count[pick].astype(np.int64))
        bad = 0
        if cnt_ok:
            idx = np.concatenate([np.arange(res.manifold_start[p], res.manifold_start[p] + res.manifold_count[p]) for p in pick]) if len(pick) else np.zeros(0, int)
            got = res.contacts[idx.astype(np.int64)]
            for f in ("world1", "world2", "normal", "depth"):
                bad += int(np.sum(~np.isclose(got[f], oc[f], rtol=RTOL, atol=ATOL)))
            bad += int(np.sum(got["f1"] != oc["f1"])) + int(np.sum(got["f2"] != oc["f2"]))
        report = {
            "workload": f"config 5: {n_total} convex shapes (1/2 cuboids, 1/2 hulls <= 32 verts), fresh-world update, {world} GPU(s)",
            "n_gpus": world, "n_objects": n_total, "ms_per_update": float(ms.item()), "updates_per_s": 1e3 / float(ms.item()),
            "pairs": int(tot[0]), "contacts": int(tot[1]), "contact_pairs": int(tot[2]), "epa_overflow": int(tot[3]), "ref_panics": int(tot[4]),
            "contact_pairs_per_s": int(tot[2]) / (float(ms.item()) * 1e-3),
            "checks": {"pairs_oriented_rank0": ok_oriented, "pairs_unique_rank0": ok_unique, "subcube_objects": int(len(sub)),
                       "subcube_pairs": int(len(wkeys)), "subcube_pair_set_equals_oracle": sub_equal, "no_pair_reported_by_two_ranks": bool(no_cross_rank_dup),
                       "narrow_sample_pairs": int(len(pick)), "narrow_sample_manifold_sizes_equal": bool(cnt_ok), "narrow_sample_mismatches": int(bad)},
            "scene_build_s": t_scene,
        }
        os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
        with open(out_path, "w") as f:
            json.dump(report, f, indent=1)
        print(json.dumps(report))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def scene_subset(scene, idx):
    from ncollide_b200.scenes import WorldScene

    return WorldScene(pos=scene.pos[idx], rot=scene.rot[idx], shape_type=scene.shape_type[idx], shape_param=scene.shape_param[idx],
                      groups=None if scene.groups is None else scene.groups[idx], query_limit=scene.query_limit[idx], ang_pred=scene.ang_pred[idx],
                      hulls=scene.hulls, margin=scene.margin)


if __name__ == "__main__":
    main()
