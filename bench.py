#!/usr/bin/env python
"""bench.py — one fresh-world CollisionWorld::update per step on synthetic scenes (BASELINE.json configs[2]:
1M mixed balls / cuboids / convex hulls), plus batched TriMesh ray casting as a secondary figure.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n-objects M] [--no-rays] [--no-cpu]

N > 1 is launched by torchrun (one rank per GPU, NCCL): weak scaling — the world has N x 1M objects; every rank
computes the AABBs of its own 1M-object block, the AABBs are all-gathered (NCCL over NVLink), the LBVH is
replicated, and each rank runs the pair search for its slice of the Morton order and the narrow phase of its own
pairs.  `value` = (total objects / 1M) / step time = 1M-shape world updates per second over the whole job.

`--impl reference` times the CPU restatement of the reference (oracle/, "port": the Rust reference cannot be built
here — no rustc/cargo, nalgebra not vendored) on the host cores; it is single-threaded like the reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "world updates/sec at 1M shapes (contact pairs/sec and Mrays/s vs TriMesh reported alongside)"
UNIT = "updates/s (1M-shape worlds)"
N_PER_GPU = 1_000_000


def workload_config(n_per, world):
    """The `config` object: names the workload only, so both arms (native and --impl reference) print the same one."""
    return {
        "workload": f"configs[2]: {n_per} mixed balls/cuboids/convex hulls (<=32 verts) per GPU, fresh-world update "
                    "(AABBs -> broad-phase pair search -> contact manifolds)",
        "n_objects_total": n_per * world,
        "seed": 1003,
        "l2": "native arm: 256 MiB flush between timed iterations, working set > L2; reference arm: host memory",
        "parallelism": ("native arm: single GPU" if world == 1 else
                        f"native arm: {world} ranks, AABB block per rank, routed spatial ownership (Morton-bin owners + ghosts; records stored into "
                        "the owner's buffers over NVLink peer memory, or NCCL all-to-all: see the line's `sharding`), local LBVH per rank")
                       + "; reference arm: 1 host thread (the reference is single-threaded)",
    }


STAGE_KERNEL = {"cc_epa": "k_cc_epa", "cc_gjk": "k_cc_gjk", "cc_manifold": "k_cc_manifold", "pair_search": "k_pair_search",
                "narrow_other_join": "k_narrow"}


def source_hash():
    """Hash of the CUDA sources: a committed ncu traffic figure is only used when it was captured from this code."""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(ROOT, "ncollide_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def committed_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (profiles/r2_traffic.json), refused
    (None) when the capture was made from other kernel sources than the ones in the tree."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        d = json.load(open(p))
        if d.get("source_hash") != source_hash():
            return None, "committed capture is stale (kernel sources changed since): refused"
        return d["kernels"].get(key), f"profiles/r2_traffic.json ({d.get('captured')})"
    except Exception:
        return None, "no capture"


def live_traffic(regex, child_args, timeout=240):
    """One `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` pass over one launch of every kernel matching `regex`, in a child
    process that replays this bench's workload (after the timed region: nothing timed runs under the profiler).
    Returns total DRAM bytes over the matched kernels of ONE step, or None when ncu is not usable here."""
    import shutil

    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    log = os.path.join("/tmp", f"ncb_traffic_{os.getpid()}.csv")
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", f"regex:{regex}",
           "--csv", "--log-file", log, sys.executable, os.path.join(ROOT, "bench.py"), "--traffic-child"] + child_args
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0 or not os.path.exists(log):
            return None, f"ncu child failed (rc {r.returncode})"
        import csv

        rows = list(csv.reader(l for l in open(log) if l.startswith('"')))
        hdr = rows[0]
        ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        per_launch = {}
        for row in rows[1:]:
            if row[im] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = float(row[iv].replace(",", "")) * scale.get(row[iu], 1.0)
                per_launch.setdefault((row[iid], row[ik]), 0.0)
                per_launch[(row[iid], row[ik])] += v
        if not per_launch:
            return None, "no matching launch in the ncu log"
        # the child runs TRAFFIC_CHILD_STEPS identical steps: keep the launches of the last one
        names = [k[1] for k in per_launch]
        n_kernels = len(set(names))
        last = list(per_launch.items())[-n_kernels:]
        return float(sum(v for _, v in last)), f"live: ncu dram__bytes_read.sum + dram__bytes_write.sum over {[k[1][:40] for k, _ in last]}, one step"
    except Exception as ex:  # noqa: BLE001
        return None, f"ncu child: {repr(ex)[:120]}"
    finally:
        try:
            os.remove(log)
        except OSError:
            pass


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.lines = []
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


RAY_CPU_SAMPLE = 250_000


def rays_cpu_baseline(n_tris=1_000_000, n_rays=RAY_CPU_SAMPLE, kind="terrain"):
    """Timing analogue of the reference's build/ncollide3d/benches/query/ray.rs for a TriMesh: the oracle's reference-faithful BVT
    (median-split build, BinaryHeap best-first search, ray_trimesh.rs:22-50) on ONE host thread, full-size mesh, a sample of the rays."""
    from ncollide_b200.scenes import make_ray_scene
    from oracle.pyoracle import Oracle

    orc = Oracle()
    rs = make_ray_scene(kind, n_tris, n_rays, seed=1004)
    t0 = time.perf_counter()
    mesh = orc.trimesh(rs.verts, rs.tris)
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    toi, _face, _n = mesh.ray_cast(rs.origins, rs.dirs, mode=0)
    cast_s = time.perf_counter() - t0
    return {"value": n_rays / cast_s / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
            "sample": f"{n_rays} of {n_tris} rays vs the full {len(rs.tris)}-triangle {kind} TriMesh (BVT build {build_s:.2f} s not timed, like the "
                      "reference's bench); C++ restatement of the reference's BVT best-first ray cast, single thread, not the Rust binary",
            "hit_fraction": float((toi >= 0).mean()), "host_cores_available": os.cpu_count()}


def run_reference(args):
    """CPU arm: reference-faithful DBVT broad phase + narrow phase of the oracle, single thread (the reference is
    single-threaded: no thread / rayon / atomic use anywhere in its src/), on the native arm's config: every step is one full
    1M-object update (6-7 s), so the driver's --steps 20 --warmup 5 takes under three minutes.  At N > 1 the native arm's world has
    N x 1M objects; the CPU sample stays one 1M-object world per step (value is in 1M-shape world updates per second either way)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from ncollide_b200.scenes import config_scene
    from oracle.pyoracle import Oracle

    orc = Oracle()
    total_steps = args.steps + args.warmup
    n_per = args.n_objects or N_PER_GPU
    world = max(1, args.gpus)
    n = n_per  # the sample: one GPU's share of the workload
    scene = config_scene(3, n)
    times = []
    counts = None
    for i in range(total_steps):
        t, counts = orc.world_update_timed(scene)
        if i >= args.warmup:
            times.append(sum(t))
    ms = 1e3 * sum(times) / len(times)
    value = (n / 1e6) / (ms / 1e3)
    rays = None
    if not args.no_rays:
        n_tris = 1_000_000 if not args.n_objects else max(1000, args.n_objects)
        rc = rays_cpu_baseline(n_tris, min(RAY_CPU_SAMPLE, n_tris))
        rays = {"metric": "Mrays/s vs TriMesh", "value": rc["value"], "unit": "Mrays/s", "cpu_baseline": rc,
                "workload": f"{n_tris} rays per GPU vs {n_tris}-triangle terrain TriMesh, identity pose (first hit + TOI + normal)"}
    sample = (f"every step is one full {n}-object cfg3 update" if world == 1 else
              f"every step is one {n}-object cfg3 update = one GPU's share of the {n * world}-object world") + \
        "; C++ restatement of the reference (DBVT + per-pair generators), not the Rust binary"
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n_per, world),
        "same_config": True,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "counts": {"n_pairs": counts[0], "n_contacts": counts[1], "n_contact_pairs": counts[2]},
        "contact_pairs_per_sec": counts[2] / (ms / 1e3),
        "host_cores_available": os.cpu_count(),
        "rays": rays,
    }
    print(json.dumps(line))
    return 0


class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 3}


def run_native(args):
    import torch

    from ncollide_b200.scenes import config_scene, make_ray_scene
    from ncollide_b200.world import Context
    from ncollide_b200 import _ffi
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = Context(local_rank)
    stream = torch.cuda.Stream(dev)  # a real (non-default) stream, shared by torch, NCCL ordering and the library
    torch.cuda.set_stream(stream)
    ctx.lib.ncb_set_stream(ctx.h, C.c_void_p(stream.cuda_stream))

    n_per = args.n_objects or N_PER_GPU
    if args.rays_only:
        n_per = 2000
    n_total = n_per * world
    scene = config_scene(3, n_total)
    ctx.set_hulls(scene.hulls)

    # pinned host staging for the end-to-end arm
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    keep = []
    pin_scene = type(scene)(**{**scene.__dict__})
    for f in ("pos", "rot", "shape_type", "shape_param", "groups", "query_limit", "ang_pred"):
        t, a = pinned(getattr(scene, f))
        keep.append(t)
        setattr(pin_scene, f, a)
    ctx.set_objects(pin_scene)
    ctx.synchronize()

    from ncollide_b200.parallel import ShardedWorld

    sharded = ShardedWorld(ctx, scene, world, rank, dev)
    my_begin, my_end = sharded.obj_begin, sharded.obj_end
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    lib, h = ctx.lib, ctx.h
    counts_c = _ffi.UpdateCountsC()

    def step_device():
        """One update with device-resident inputs."""
        if world == 1:
            r = lib.ncb_world_update_device(h, C.c_float(scene.margin), C.c_uint32(0), C.c_uint32(0xFFFFFFFF), C.byref(counts_c))
            ctx.check(r, "ncb_world_update_device")
        else:
            return sharded.step(counts_c)
        return ctx._counts(counts_c)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_steps(fn, steps, warmup, with_flush=True):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        for _ in range(steps):
            if with_flush:
                flush.fill_(1)  # L2 flush (256 MiB > 126 MB L2) between timed iterations, outside the timed events
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out = fn()
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        ms = [a.elapsed_time(b) for a, b in evs]
        return ms, out

    # ---- timed region: device-resident inputs -------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_list, counts = timed_steps(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_mean = sum(ms_list) / len(ms_list)
    if dist is not None:
        t = torch.tensor([ms_mean], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
        agg = torch.tensor([counts["n_pairs"], counts["n_contacts"], counts["n_contact_pairs"]], device=dev, dtype=torch.int64)
        dist.all_reduce(agg)
        tot_pairs, tot_contacts, tot_contact_pairs = (int(x) for x in agg.tolist())
    else:
        ms_step = ms_mean
        tot_pairs, tot_contacts, tot_contact_pairs = counts["n_pairs"], counts["n_contacts"], counts["n_contact_pairs"]
    value = (n_total / 1e6) / (ms_step / 1e3)

    # ---- per-stage times (CUDA events on the launching stream) for the roofline ----------------------------
    ctx.profile_enable(True)
    stage_acc = {}
    launches_total = 0
    reps = 3
    for _ in range(reps):
        flush.fill_(1)
        step_device()
        prof = ctx.profile_get()
        launches_total = sum(p[2] for p in prof)
        for name, ms, _l in prof:
            stage_acc[name] = stage_acc.get(name, 0.0) + ms / reps
    ctx.profile_enable(False)

    peak, peak_kind = measured_peaks()
    stages = []
    for name, ms in stage_acc.items():
        stages.append({"stage": name, "ms": round(ms, 4)})
    dom_name = max(stage_acc, key=stage_acc.get) if stage_acc else None
    total_bytes = world_total_bytes(n_total // world if world > 1 else n_total, counts, scene)
    roofline = None
    if dom_name:
        dom_ms = stage_acc[dom_name]
        dom_bytes = stage_bytes(dom_name, n_total // world if world > 1 else n_total, counts, scene)
        ach = dom_bytes / (dom_ms / 1e3) / 1e9
        roofline = {
            "bound": "hbm", "kernel": dom_name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": None, "traffic_source": None,
            "peak_kind": f"{peak_kind} copy bandwidth", "kernel_ms": dom_ms, "algorithmic_bytes": dom_bytes,
            "share_of_step": dom_ms / max(sum(stage_acc.values()), 1e-9),
            "whole_step": {"algorithmic_bytes": total_bytes, "achieved": total_bytes / (ms_step / 1e3) / 1e9,
                           "frac": total_bytes / (ms_step / 1e3) / 1e9 / peak},
        }

    # ---- end to end: host buffers through ncb_world_update (N = 1) / staged calls (N > 1) -------------------
    bufs = ctx.alloc_result_buffers(int(counts["n_pairs"] * 1.1) + 1024, int(counts["n_contacts"] * 1.1) + 1024)
    pin_out = {}
    for k, a in bufs.items():
        t = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
        keep.append(t)
        pin_out[k] = t.numpy().view(a.dtype).reshape(a.shape)

    def step_e2e():
        if world == 1:
            # the objects persist in the world (world.rs:64-96); a step sets the poses and updates (world.rs:104-119)
            r = lib.ncb_world_update_poses(
                h, C.c_uint32(n_total), _ffi.ptr(pin_scene.pos), _ffi.ptr(pin_scene.rot), C.c_float(scene.margin), _ffi.ptr(pin_out["pairs"]),
                C.c_uint32(len(pin_out["pairs"])), _ffi.ptr(pin_out["algo"]), _ffi.ptr(pin_out["start"]), _ffi.ptr(pin_out["count"]),
                _ffi.ptr(pin_out["contacts"]), C.c_uint32(len(pin_out["contacts"])), C.byref(counts_c))
            ctx.check(r, "ncb_world_update_poses")
        else:
            # every rank uploads the poses of its own block, steps, and reads its own results back
            sharded.upload_own_poses(pin_scene.pos, pin_scene.rot)
            ctx.check(lib.ncb_world_fetch_early(h, _ffi.ptr(pin_out["pairs"]), C.c_uint32(len(pin_out["pairs"])), _ffi.ptr(pin_out["algo"]),
                                                _ffi.ptr(pin_out["contacts"]), C.c_uint32(len(pin_out["contacts"]))), "fetch_early")
            sharded.step(counts_c, with_poses=True)  # routed mode: the poses travel with the records (no pose all-gather)
            ctx.check(lib.ncb_world_fetch(h, _ffi.ptr(pin_out["pairs"]), C.c_uint32(len(pin_out["pairs"])), _ffi.ptr(pin_out["algo"]),
                                          _ffi.ptr(pin_out["start"]), _ffi.ptr(pin_out["count"]), _ffi.ptr(pin_out["contacts"]),
                                          C.c_uint32(len(pin_out["contacts"]))), "fetch")
        return ctx._counts(counts_c)

    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = (n_total / 1e6) / (e2e_ms / 1e3)
    n_up = n_total
    # per step: the poses (pos 12 B + rot 16 B per object); shapes / groups / query limits persist on the device like the objects of a
    # CollisionWorld persist between updates
    h2d = n_up * 28 if world == 1 else n_per * 28
    d2h = counts["n_pairs"] * (8 + 1 + 4 + 1) + counts["n_contacts"] * 52 + 256

    # ---- secondary figure: batched TriMesh ray casting (configs[3] / SURVEY §8d cfg 4a + 4b) ---------------
    rays = None
    if not args.no_rays:
        n_tris, n_rays = (1_000_000, 1_000_000) if (not args.n_objects or args.rays_only) else (max(1000, args.n_objects), max(1000, args.n_objects))
        fmax = float(np.finfo(np.float32).max)
        variants = {}
        for kind, posed in (("terrain", False), ("terrain", True), ("soup", False), ("soup", True)):
            rs = make_ray_scene(kind, n_tris, n_rays * world, seed=1004, random_pose=posed)
            mesh = ctx.trimesh(rs.verts, rs.tris)
            lo, hi = rank * n_rays, (rank + 1) * n_rays
            origins, dirs = rs.origins[lo:hi], rs.dirs[lo:hi]
            pose_arg = None
            if posed:  # the rays move with the mesh, so the posed cast answers the same question as the identity one
                from ncollide_b200.scenes import transform_rays

                origins, dirs = transform_rays(rs.pose, origins, dirs)
                pose_arg = np.ascontiguousarray(rs.pose, dtype=np.float32)
            d_o = torch.from_numpy(origins).to(dev)
            d_d = torch.from_numpy(dirs).to(dev)
            d_toi = torch.empty(n_rays, dtype=torch.float32, device=dev)
            d_face = torch.empty(n_rays, dtype=torch.int32, device=dev)
            d_n = torch.empty((n_rays, 3), dtype=torch.float32, device=dev)

            def ray_step():
                ctx.check(lib.ncb_trimesh_ray_cast_device(mesh.h, _ffi.ptr(pose_arg), C.c_uint32(n_rays), C.c_void_p(d_o.data_ptr()),
                                                          C.c_void_p(d_d.data_ptr()), C.c_float(fmax), C.c_void_p(d_toi.data_ptr()),
                                                          C.c_void_p(d_face.data_ptr()), C.c_void_p(d_n.data_ptr())), "ray_cast_device")

            rms, _ = timed_steps(ray_step, max(args.steps, 5), 3)
            rms_mean = sum(rms) / len(rms)
            if dist is not None:
                t = torch.tensor([rms_mean], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                rms_mean = float(t.item())
            T, V = len(rs.tris), len(rs.verts)
            ray_bytes = n_rays * (24 + 8 + 12) + V * 12 + T * 12 + (2 * T - 1) * 32
            # end to end with host buffers (pinned, like the world update's)
            o_t, o_h = pinned(origins)
            d_t, d_h = pinned(dirs)
            keep.extend([o_t, d_t])
            out_pin = {}
            for nm, shp, dt in (("toi", (n_rays,), np.float32), ("face", (n_rays,), np.uint32), ("normal", (n_rays, 3), np.float32)):
                tt = torch.empty(int(np.prod(shp)) * 4, dtype=torch.uint8).pin_memory()
                keep.append(tt)
                out_pin[nm] = tt.numpy().view(dt).reshape(shp)
            mesh.toi_and_normal_with_ray(pose_arg, o_h, d_h, out=out_pin)
            barrier()
            t0 = time.perf_counter()
            for _ in range(5):
                mesh.toi_and_normal_with_ray(pose_arg, o_h, d_h, out=out_pin)
            barrier()
            ray_e2e_ms = (time.perf_counter() - t0) * 1e3 / 5
            assert np.array_equal(out_pin["toi"], d_toi.cpu().numpy()), "host-buffer ray cast differs from the device-resident one"
            if dist is not None:
                t = torch.tensor([ray_e2e_ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ray_e2e_ms = float(t.item())
            name = f"{kind}_{'random_pose' if posed else 'identity'}"
            variants[name] = {
                "value": n_rays * world / (rms_mean / 1e3) / 1e6, "unit": "Mrays/s", "ms_per_batch": rms_mean,
                "workload": f"{n_rays} rays per GPU vs {T}-triangle {kind} TriMesh, {'one random mesh pose' if posed else 'identity pose'} (first hit + TOI + normal)",
                "hit_fraction": float((d_toi >= 0).float().mean().item()),
                "roofline": {"bound": "hbm", "achieved": ray_bytes / (rms_mean / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": ray_bytes / (rms_mean / 1e3) / 1e9 / peak, "algorithmic_bytes": ray_bytes},
                "e2e": {"value": n_rays * world / (ray_e2e_ms / 1e3) / 1e6, "unit": "Mrays/s", "ms_per_batch": ray_e2e_ms,
                        "h2d_bytes_per_step": n_rays * 24, "d2h_bytes_per_step": n_rays * 20,
                        "call": "ncb_trimesh_ray_cast_uv: pinned host rays in, toi / face / normal out to pinned host buffers, chunked "
                                "upload | cast | download pipeline"},
            }
            mesh.close()
            del d_o, d_d, d_toi, d_face, d_n
        head = variants["terrain_identity"]
        rays = {"metric": "Mrays/s vs TriMesh", **head, "variants": {k: v for k, v in variants.items() if k != "terrain_identity"}}
        tr, tr_src = committed_traffic("k_ray_cast")
        rays["roofline"]["traffic"], rays["roofline"]["traffic_source"] = (tr, tr_src) if n_rays == 1_000_000 else (None, None)
        if rank == 0 and world == 1 and not args.no_cpu:
            rays["cpu_baseline"] = rays_cpu_baseline(n_tris, min(RAY_CPU_SAMPLE, n_rays))
        if args.rays_only:
            if rank == 0:
                try:
                    rays["dim2_polyline_rays"] = bench_polyline_rays(ctx)
                except Exception as ex:  # noqa: BLE001
                    rays["dim2_polyline_rays"] = {"error": repr(ex)[:200]}
                print(json.dumps(rays))
            ctx.close()
            return 0

    # ---- CPU baseline beside it (rank 0, N = 1): the oracle port on a bounded sample -----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.pyoracle import Oracle

        orc = Oracle()
        n_cpu = min(n_total, 1_000_000)
        cs = scene if n_cpu == n_total else config_scene(3, n_cpu)
        t, c = orc.world_update_timed(cs)
        cpu_ms = sum(t) * 1e3
        cpu = {"value": (n_cpu / 1e6) / (cpu_ms / 1e3), "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"one full {n_cpu}-object cfg3 step (aabb {t[0]:.2f}s + DBVT broad phase {t[1]:.2f}s + narrow phase {t[2]:.2f}s); "
                         "C++ restatement of the reference, single thread like the reference, not the Rust binary",
               "host_cores_available": os.cpu_count(), "pairs": c[0], "contacts": c[1]}

    # ---- DRAM traffic of the dominant stage's kernels: one ncu pass in a child process (nothing timed runs under the profiler) ----
    if rank == 0 and world == 1 and roofline and n_per == N_PER_GPU:
        key = STAGE_KERNEL.get(dom_name)
        tr, src = (None, "not measured")
        if key and not args.no_traffic:
            tr, src = live_traffic(key, [])
        if tr is None and key:
            tr2, src2 = committed_traffic(key)
            tr, src = tr2, (src2 if tr2 is not None else f"{src}; {src2}")
        roofline["traffic"], roofline["traffic_source"] = tr, src

    # ---- widened rows, N = 1 only, reported beside the headline; never allowed to break the line ----
    widened = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            widened = bench_widened(ctx, scene)
        except Exception as ex:  # noqa: BLE001
            widened = {"error": repr(ex)[:200]}
        try:
            widened["proximity_sensors"] = bench_proximity(ctx, scene)
        except Exception as ex:  # noqa: BLE001
            widened["proximity_sensors"] = {"error": repr(ex)[:200]}
        try:
            widened["dim2_contact"] = bench_dim2(ctx)
        except Exception as ex:  # noqa: BLE001
            widened["dim2_contact"] = {"error": repr(ex)[:200]}
        try:
            widened["dim2_polyline_rays"] = bench_polyline_rays(ctx)
        except Exception as ex:  # noqa: BLE001
            widened["dim2_polyline_rays"] = {"error": repr(ex)[:200]}

    # ---- secondary worlds (after everything else: they replace the world on the device) -----------------------------------
    #   strong scaling: ONE fixed 8 M-object cfg3 world at every N (the driver's N = 1, 2, 4, 8 runs give the curve);
    #   config 5 at full size: 16 M convex shapes, only when 8 GPUs are present (BASELINE.json configs[4]).
    secondary = None
    if not args.no_secondary and n_per == N_PER_GPU:
        secondary = {}
        jobs = [("strong_scaling_8M", 3, 8_000_000)] + ([("cfg5_16M", 5, 16_000_000)] if world == 8 else [])
        for name, cfg, n_sec in jobs:
            try:
                t_sec = time.perf_counter()
                sc = config_scene(cfg, n_sec)
                ctx.set_hulls(sc.hulls)
                ctx.set_objects(sc)
                ctx.synchronize()
                sw = ShardedWorld(ctx, sc, world, rank, dev)
                cc = _ffi.UpdateCountsC()

                def sec_step():
                    if world == 1:
                        ctx.check(lib.ncb_world_update_device(h, C.c_float(sc.margin), C.c_uint32(0), C.c_uint32(0xFFFFFFFF), C.byref(cc)), "update")
                        return ctx._counts(cc)
                    return sw.step(cc)

                sms, scounts = timed_steps(sec_step, 5, 3)
                sms_mean = sum(sms) / len(sms)
                tot = [scounts["n_pairs"], scounts["n_contacts"], scounts["n_contact_pairs"], scounts["epa_overflow"]]
                if dist is not None:
                    t = torch.tensor([sms_mean], device=dev, dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    sms_mean = float(t.item())
                    agg = torch.tensor(tot, device=dev, dtype=torch.int64)
                    dist.all_reduce(agg)
                    tot = [int(x) for x in agg.tolist()]
                secondary[name] = {
                    "workload": f"{sc.name}: {n_sec} objects in ONE world over {world} GPU(s), fresh-world update, device-resident inputs",
                    "scaling": "strong" if name.startswith("strong") else "config 5 (full size)",
                    "ms_per_update": sms_mean, "updates_per_s": 1e3 / sms_mean, "steps": 5, "warmup": 3,
                    "pairs": tot[0], "contacts": tot[1], "contact_pairs": tot[2], "contact_pairs_per_sec": tot[2] / (sms_mean / 1e3),
                    "epa_overflow": tot[3], "sharding": sw.mode if world > 1 else "single GPU",
                    "wall_s_incl_scene_build": round(time.perf_counter() - t_sec, 1),
                }
            except Exception as ex:  # noqa: BLE001  (never allowed to break the headline line)
                secondary[name] = {"error": repr(ex)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n_per, world),
            "sharding": sharded.mode if world > 1 else "single GPU",
            "pairs": tot_pairs, "contacts": tot_contacts, "contact_pairs": tot_contact_pairs,
            "contact_pairs_per_sec": tot_contact_pairs / (ms_step / 1e3),
            "broad_phase_pairs_per_sec": tot_pairs / (ms_step / 1e3),
            "stages_ms": stages,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "call": "ncb_world_update_poses: pinned host poses in (set_position on every object), update, every pair / manifold / contact "
                            "out to pinned host buffers" if world == 1 else f"per rank: ncb_set_positions_range (own block) + {sharded.mode} sharded update "
                            "(p2p / routed: the poses travel inside the routed records, over NVLink peer memory / NCCL all-to-all; spatial / "
                            "slices: NCCL pose all-gather) + ncb_world_fetch of the rank's own pairs and contacts"},
            "gpu_launches": int(launches_total * args.steps),
            "gpu_launches_per_step": int(launches_total),
            "clocks": clocks,
            "counts": {k: v for k, v in counts.items() if k != "n_algo"} | {"n_algo": counts["n_algo"]},
            "rays": rays,
            "widened": widened,
            "secondary": secondary,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


TRAFFIC_CHILD_STEPS = 3


def run_traffic_child(args):
    """The workload's device-resident steps and nothing else (the parent runs this under ncu)."""
    import ctypes as C

    from ncollide_b200 import _ffi
    from ncollide_b200.scenes import config_scene
    from ncollide_b200.world import Context

    ctx = Context(0)
    scene = config_scene(3, args.n_objects or N_PER_GPU)
    ctx.set_hulls(scene.hulls)
    ctx.set_objects(scene)
    counts_c = _ffi.UpdateCountsC()
    for _ in range(TRAFFIC_CHILD_STEPS):
        ctx.check(ctx.lib.ncb_world_update_device(ctx.h, C.c_float(scene.margin), C.c_uint32(0), C.c_uint32(0xFFFFFFFF), C.byref(counts_c)), "update")
    ctx.close()
    return 0


def stage_bytes(name, n, counts, scene):
    P, Cn = counts["n_pairs"], counts["n_contacts"]
    H = int((scene.shape_type == 2).sum()) * n / max(scene.n, 1)
    vbar = float(np.diff(scene.hulls.vert_off).mean()) if scene.hulls.n_hulls else 0.0
    na = counts["n_algo"]
    types = scene.shape_type
    fc, fh = [(types == t).mean() for t in (1, 2)]
    # pairs per key estimated from the dispatched algorithm counts and the cuboid / hull mix of the scene
    w_bcub, w_bh = fc / max(fc + fh, 1e-9), fh / max(fc + fh, 1e-9)
    w_cc, w_ch, w_hh = fc * fc, 2 * fc * fh, fh * fh
    wsum = max(w_cc + w_ch + w_hh, 1e-9)
    key_pairs = {
        "narrow_ball_ball": na["ball_ball"], "narrow_plane": na["plane_ball"] + na["plane_convex"],
        "narrow_ball_cuboid": na["ball_convex"] * w_bcub, "narrow_ball_hull": na["ball_convex"] * w_bh,
        "narrow_cuboid_cuboid": na["convex_convex"] * w_cc / wsum, "narrow_cuboid_hull": na["convex_convex"] * w_ch / wsum,
        "narrow_hull_hull": na["convex_convex"] * w_hh / wsum,
    }
    hull_ops = {"narrow_ball_ball": 0, "narrow_plane": 0, "narrow_ball_cuboid": 0, "narrow_ball_hull": 1, "narrow_cuboid_cuboid": 0,
                "narrow_cuboid_hull": 1, "narrow_hull_hull": 2}
    cc = na["convex_convex"]
    cc_hull_ops = (key_pairs["narrow_cuboid_hull"] + 2 * key_pairs["narrow_hull_hull"]) / max(cc, 1)  # hull operands per cc pair
    cc_contacts = Cn * cc / max(P, 1)
    if name == "cc_gjk":      # read the pair + both operands (+ hull vertices), decide
        return cc * (2 * 44 + 4) + cc * cc_hull_ops * 12 * vbar
    if name == "cc_epa":      # re-read operands of the penetrating subset, hand the witness points on
        ne = counts.get("n_epa_pairs", 0)
        return ne * (2 * 44 + 40) + ne * cc_hull_ops * 12 * vbar
    if name == "cc_manifold":  # operands of the pairs that reach clipping + the contacts written
        nm = counts.get("n_manifold_jobs", 0)
        return nm * (2 * 44 + 40) + cc_contacts * 48 + cc * 5
    if name == "aabb":
        return n * (28 + 16 + 24) + H * 12 * vbar
    if name == "morton_sort":
        return n * (24 + 8)
    if name == "lbvh_build":
        return n * (24 + 64)
    if name == "pair_search":
        return n * (24 + 12) + P * 8
    if name == "pair_sort":
        return P * 9 * 2
    if name in key_pairs:
        kp = key_pairs[name]
        share = kp / max(P, 1)
        return kp * (2 * 44 + 4) + kp * hull_ops[name] * 12 * vbar + Cn * share * 48
    return 0.0


def world_total_bytes(n, counts, scene):
    # compulsory bytes of the whole step: object records in, AABBs, pairs, operands per pair, contacts out
    names = ["aabb", "pair_search", "narrow_ball_ball", "narrow_plane", "narrow_ball_cuboid", "narrow_ball_hull", "narrow_cuboid_cuboid",
             "narrow_cuboid_hull", "narrow_hull_hull"]
    return float(sum(stage_bytes(k, n, counts, scene) for k in names))



def bench_widened(ctx, scene):
    """Stepping world (CollisionWorld::update over several steps, 10 % of the poses set per step) and world ray queries
    (first_interference_with_ray) on the same scene, through the C ABI with host buffers; wall clock."""
    from ncollide_b200.world import SteppingWorld

    rng = np.random.default_rng(11)
    n = scene.n
    w = SteppingWorld(ctx, scene)
    t0 = time.perf_counter()
    w.update(fetch=False)
    first_ms = (time.perf_counter() - t0) * 1e3
    pos, rot = scene.pos.copy(), scene.rot.copy()
    times = []
    info = None
    for _ in range(8):
        idx = np.sort(rng.choice(n, size=max(1, n // 10), replace=False)).astype(np.uint32)
        pos[idx] = (pos[idx] + rng.normal(0, 0.004, size=(len(idx), 3))).astype(np.float32)
        t0 = time.perf_counter()
        w.set_positions(idx, pos[idx], rot[idx])
        info = w.update(fetch=False)
        times.append((time.perf_counter() - t0) * 1e3)
    side = float(scene.pos.max())
    n_rays = min(n, 1_000_000)
    ro = rng.uniform(0, side, size=(n_rays, 3)).astype(np.float32)
    rd = rng.normal(size=(n_rays, 3)).astype(np.float32)
    rt = []
    rows = 0
    for _ in range(3):
        t0 = time.perf_counter()
        hit = w.ray_cast(ro, rd, 20.0, first_only=True)
        rt.append((time.perf_counter() - t0) * 1e3)
        rows = len(hit[0])
    w.close()
    return {
        "stepping_world": {"workload": f"{n} objects, {n // 10} poses set per update (ncb_sim_set_positions + ncb_sim_step)",
                           "ms_per_update": statistics.median(times[2:]), "first_update_ms": first_ms, "pairs": info["counts"]["n_pairs"],
                           "contacts": info["counts"]["n_contacts"], "pairs_regenerated": info["counts"]["n_manifold_jobs"]},
        "world_ray_queries": {"workload": f"{n_rays} rays, first_interference_with_ray within 20 units, {n}-object world (ncb_sim_ray_cast)",
                              "ms": min(rt[1:]), "Mrays_per_s": n_rays / (min(rt[1:]) / 1e3) / 1e6, "hits": rows},
    }


def bench_dim2(ctx, n_pairs=1_000_000, cpu_sample=100_000):
    """Widened row N4 (2-D build, first slice): ncollide2d query::contact for a batch of random 2-D cuboid / polygon / ball pairs
    through ncb2d_contact (host buffers: shapes + poses in, contacts out; wall clock), next to the oracle on one host core."""
    from ncollide_b200 import dim2

    rng = np.random.default_rng(21)
    sh = dim2.Shapes2D()
    t1 = rng.choice([0, 1, 2], size=n_pairs)
    t2 = rng.choice([0, 1, 2], size=n_pairs)
    # a small library of shapes, indexed at random (building a million Python shape objects would dominate)
    lib_t, lib_p = [], []
    for t in (0, 1, 2):
        for _ in range(64):
            if t == 0:
                sh.ball(rng.uniform(0.2, 0.6))
            elif t == 1:
                sh.cuboid(rng.uniform(0.2, 0.6), rng.uniform(0.2, 0.6))
            else:
                k = int(rng.integers(3, 13))
                ang = np.sort(rng.uniform(0, 2 * np.pi, size=k)) + np.arange(k) * 1e-3
                sh.polygon(np.stack([0.5 * np.cos(ang), 0.35 * np.sin(ang)], axis=1))
    typ, par, pts, nrm = sh.arrays()
    pick1, pick2 = t1 * 64 + rng.integers(0, 64, size=n_pairs), t2 * 64 + rng.integers(0, 64, size=n_pairs)
    c1 = rng.uniform(-50, 50, size=(n_pairs, 2))
    m1 = dim2.isometry2(c1, rng.uniform(-np.pi, np.pi, size=n_pairs))
    m2 = dim2.isometry2(c1 + rng.uniform(-1.2, 1.2, size=(n_pairs, 2)), rng.uniform(-np.pi, np.pi, size=n_pairs))
    args = (typ[pick1], par[pick1], m1, typ[pick2], par[pick2], m2, pts)
    dim2.contact(ctx, *args, prediction=0.02, poly_normals=nrm)
    tms = []
    for _ in range(3):  # pageable host buffers: the median of three calls (a single call has been seen at 8x under host memory pressure)
        t0 = time.perf_counter()
        found, out, info = dim2.contact(ctx, *args, prediction=0.02, poly_normals=nrm)
        tms.append((time.perf_counter() - t0) * 1e3)
    ms = float(np.median(tms))
    res = {"workload": f"{n_pairs} random 2-D pairs (balls, cuboids, convex polygons of 3-12 vertices), query::contact with prediction 0.02",
           "ms": ms, "ms_calls": [round(t, 2) for t in tms], "Mpairs_per_s": n_pairs / ms / 1e3, "contacts_found": int(found.sum()), **info}
    # the 2-D world update: 1 M objects (balls, cuboids, polygons), about 3 fat-box neighbours each
    n_w = n_pairs
    side = float(np.sqrt(n_w * 0.8 / 2.5))
    w = dim2.World2D.from_library(sh, rng.integers(0, 192, size=n_w), rng.uniform(0, side, size=(n_w, 2)), rng.uniform(-np.pi, np.pi, size=n_w))
    import torch

    def pin(shape, dtype):
        t = torch.empty(int(np.prod(shape)) * np.dtype(dtype).itemsize, dtype=torch.uint8).pin_memory()
        return t, t.numpy().view(dtype).reshape(shape)

    keep2 = []
    for f in ("pos", "rot", "type", "param", "query_limit", "ang_pred"):  # page-locked inputs, like the 3-D end-to-end arm
        t, a = pin(getattr(w, f).shape, getattr(w, f).dtype)
        a[...] = getattr(w, f)
        keep2.append(t)
        setattr(w, f, a)
    first = dim2.world_update(ctx, w)
    cap_p, cap_c = len(first["pairs"]) + 4096, len(first["contacts"]) + 4096
    pb = {}
    for k, shp, dt in (("pairs", (cap_p, 2), np.uint32), ("start", (cap_p,), np.uint32), ("count", (cap_p,), np.uint8),
                       ("contacts", (cap_c, 7), np.float32), ("features", (cap_c, 2), np.uint32)):
        t, pb[k] = pin(shp, dt)
        keep2.append(t)
    dim2.world_update(ctx, w, bufs=pb)
    t0 = time.perf_counter()
    for _ in range(3):
        wr = dim2.world_update(ctx, w, bufs=pb)
    wms = (time.perf_counter() - t0) * 1e3 / 3
    res["world_update"] = {"workload": f"{n_w} 2-D objects, fresh-world update through ncb2d_world_update (page-locked host buffers in and out)", "ms": wms,
                           "pairs": int(len(wr["pairs"])), "contacts": int(len(wr["contacts"])), **wr["diag"]}
    try:
        from oracle.pyoracle import Oracle

        orc = Oracle()
        n_cw = 100_000
        cw = dim2.World2D.from_library(sh, rng.integers(0, 192, size=n_cw), rng.uniform(0, side * np.sqrt(n_cw / n_w), size=(n_cw, 2)),
                                       rng.uniform(-np.pi, np.pi, size=n_cw))
        t0 = time.perf_counter()
        op = orc.world_update2d(cw)
        res["world_update"]["cpu_baseline"] = {"ms": (time.perf_counter() - t0) * 1e3, "objects": n_cw, "pairs": int(len(op[0])), "cores": 1,
                                               "kind": "port", "note": "sweep broad phase of the oracle, not the DBVT"}
        sl = slice(0, cpu_sample)
        t0 = time.perf_counter()
        of, oo, _ = orc.contact2d(args[0][sl], args[1][sl], args[2][sl], args[3][sl], args[4][sl], args[5][sl], pts, prediction=0.02, poly_normals=nrm)
        cms = (time.perf_counter() - t0) * 1e3
        res["cpu_baseline"] = {"Mpairs_per_s": cpu_sample / cms / 1e3, "cores": 1, "kind": "port", "sample": f"the first {cpu_sample} pairs",
                               "answers_equal": bool(np.array_equal(of.astype(bool), found[sl]))}
    except Exception as ex:  # noqa: BLE001
        res["cpu_baseline"] = {"error": repr(ex)[:120]}
    return res


def bench_polyline_rays(ctx, n_edges=1_000_000, n_rays=1_000_000, cpu_sample=100_000):
    """Widened row N4 (2-D build): RayCast for Polyline, 1 M rays against a 1 M-edge height profile.  Device-resident rays timed with CUDA
    events on the context's stream (it is torch's current stream here), and the host-buffer call (page-locked buffers, wall clock)."""
    import ctypes as C

    import torch

    from ncollide_b200 import dim2
    from ncollide_b200.scenes import make_polyline_scene

    pts, edges, o, d = make_polyline_scene("terrain", n_edges, n_rays, 31)
    pl = dim2.Polyline(ctx, pts, edges)

    def pin(a):
        t = torch.from_numpy(a.copy()).pin_memory()
        return t, t.numpy()

    keep = []
    to, po = pin(o)
    td, pd = pin(d)
    out = {}
    for k, shp, dt in (("toi", (n_rays,), np.float32), ("feature", (n_rays,), np.uint32), ("normal", (n_rays, 2), np.float32)):
        t, out[k] = pin(np.zeros(shp, dtype=dt))
        keep.append(t)
    pl.toi_and_normal_with_ray(None, po, pd, out=out)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        toi, feat, nrm = pl.toi_and_normal_with_ray(None, po, pd, out=out)
        ts.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = float(np.median(ts))
    dev = torch.device("cuda", torch.cuda.current_device())
    d_o, d_d = to.to(dev), td.to(dev)
    d_toi, d_feat = torch.empty(n_rays, dtype=torch.float32, device=dev), torch.empty(n_rays, dtype=torch.int32, device=dev)
    d_n = torch.empty((n_rays, 2), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream()  # main() made it the context's stream

    def cast():
        ctx.check(ctx.lib.ncb2d_polyline_ray_cast_device(pl.h, None, C.c_uint32(n_rays), C.c_void_p(d_o.data_ptr()), C.c_void_p(d_d.data_ptr()),
                                                         C.c_float(np.finfo(np.float32).max), None, C.c_void_p(d_toi.data_ptr()),
                                                         C.c_void_p(d_feat.data_ptr()), C.c_void_p(d_n.data_ptr())), "ncb2d_polyline_ray_cast_device")

    cast()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        cast()
    e1.record(stream)
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / 10
    same = bool(np.array_equal(d_toi.cpu().numpy().view(np.uint32), toi.view(np.uint32)))
    res = {"workload": f"{n_rays} rays against a {n_edges}-edge Polyline (noisy height profile), toi_and_normal_with_ray",
           "device_ms": dev_ms, "device_Mrays_per_s": n_rays / dev_ms / 1e3, "e2e_ms": e2e_ms, "e2e_Mrays_per_s": n_rays / e2e_ms / 1e3,
           "hits": int((toi >= 0).sum()), "device_equals_host_call": same, "traversal_overflows": int(ctx.traversal_overflows())}
    try:
        from oracle.pyoracle import Oracle

        op = Oracle().polyline(pts, edges)
        t0 = time.perf_counter()
        ot, of, on = op.ray_cast(o[:cpu_sample], d[:cpu_sample], mode=0)
        cms = (time.perf_counter() - t0) * 1e3
        res["cpu_baseline"] = {"Mrays_per_s": cpu_sample / cms / 1e3, "cores": 1, "kind": "port", "sample": f"the first {cpu_sample} rays",
                               "toi_equal": bool(np.array_equal(ot.view(np.uint32), toi[:cpu_sample].view(np.uint32)))}
    except Exception as ex:  # noqa: BLE001
        res["cpu_baseline"] = {"error": repr(ex)[:120]}
    pl.close()
    return res


def bench_proximity(ctx, scene, fraction=0.2, reps=6):
    """Widened row N4 (proximity-only interactions): the same scene with a seeded `fraction` of the objects turned into
    GeometricQueryType::Proximity sensors.  Fresh-world update on the device (wall clock around ncb_world_update_device, which
    ends with the counter read-back) next to the same update without sensors, and the batched detector entry (ncb_proximity)
    on the update's own sensor pairs with host buffers."""
    import copy

    from ncollide_b200.scenes import with_sensors

    s = with_sensors(copy.copy(scene), fraction, 5)
    out = {"workload": f"{s.n} objects, {int(s.query_kind.sum())} of them Proximity sensors (margin = their query limit), fresh-world update"}

    def timed():
        ts, c = [], None
        for _ in range(reps):
            t0 = time.perf_counter()
            c = ctx.world_update_device(s.margin)
            ts.append((time.perf_counter() - t0) * 1e3)
        return statistics.median(ts[2:]), c

    ctx.set_scene(scene)
    out["ms_per_update_without_sensors"], c0 = timed()
    ctx.set_scene(s)
    out["ms_per_update"], c = timed()
    res = ctx.world_fetch(c)
    sel = res.pair_algo == 6
    out.update({"pairs": c["n_pairs"], "proximity_pairs": c["n_algo"]["proximity"], "statuses": c["n_proximity"], "contacts": c["n_contacts"],
                "contacts_without_sensors": c0["n_contacts"]})
    pp = np.ascontiguousarray(res.pairs[sel])
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        st = ctx.proximity(pp)
        ts.append((time.perf_counter() - t0) * 1e3)
    assert np.array_equal(st, res.proximity[sel]), "ncb_proximity disagrees with the world update on the same pairs"
    out["batch_entry"] = {"pairs": int(len(pp)), "ms": min(ts[1:]), "Mpairs_per_s": len(pp) / (min(ts[1:]) / 1e3) / 1e6,
                          "note": "ncb_proximity, host buffers (8 B in + 1 B out per pair), unsorted pairs"}
    ctx.set_scene(scene)  # back to a sensor-free world
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--n-objects", type=int, default=0, help="objects per GPU (default 1,000,000)")
    ap.add_argument("--no-rays", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the stepping-world / world-query figures")
    ap.add_argument("--no-secondary", action="store_true", help="skip the strong-scaling (8 M objects) and config-5 (16 M, 8 GPUs) worlds")
    ap.add_argument("--rays-only", action="store_true", help="debug: only the ray-casting sub-benchmark, prints its dict")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu child that measures the dominant kernel's DRAM traffic")
    ap.add_argument("--traffic-child", action="store_true", help="internal: replay the workload's device steps (run under ncu by the parent)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.traffic_child:
        return run_traffic_child(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
