"""Multi-GPU plumbing for the world update: one process per GPU, objects block-partitioned for
the AABB computation.  Default ("routed"): every rank sends each object of its block to the rank that owns its Morton bin and,
as a ghost, to the ranks whose region its fat box meets (two small all-reduces, two all-to-alls, one all-reduce of the regions);
nothing is all-gathered and no rank scans all N boxes.  "spatial" (the previous design): fat AABBs all-gathered, every rank selects
its owned + ghost objects out of all N.  "slices": replicated LBVH, query slices of the Morton order.  In every mode a rank builds
its own LBVH over what it holds and runs the pair search + narrow phase of its share of the pairs; every pair is reported by exactly
one rank.  torch.distributed is plumbing only."""
from __future__ import annotations

import ctypes as C


def shard_range(n, world, rank):
    """Contiguous block [begin, end) of `n` items owned by `rank` (the first n % world ranks get one extra)."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_gather_rows(full, begin, end, world, group=None):
    """In-place all-gather of row blocks: rank r owns full[begin_r:end_r]; afterwards every rank holds all rows.
    Blocks may have different sizes (n % world != 0)."""
    import torch
    import torch.distributed as dist

    n = full.shape[0]
    sizes = [shard_range(n, world, r) for r in range(world)]
    if all(e - b == sizes[0][1] - sizes[0][0] for b, e in sizes):
        # NCCL's in-place form (the send block IS its slot of `full`) needs no copy; gloo (CPU tests) gets a private send block
        dist.all_gather_into_tensor(full, full[begin:end] if full.is_cuda else full[begin:end].clone(), group=group)
        return full
    # ragged blocks: pad every block to the largest one (collectives need equal counts)
    mx = max(e - b for b, e in sizes)
    send = torch.zeros((mx,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    send[: end - begin] = full[begin:end]
    recv = torch.empty((world * mx,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    for r, (b, e) in enumerate(sizes):
        full[b:e] = recv[r * mx : r * mx + (e - b)]
    return full


NCB_ROUTE_REPEAT = 2


def perform_collective(op, group=None):
    """Executes one collective request of ShardedWorld.routed_plan with torch.distributed (NCCL on GPUs)."""
    import torch.distributed as dist

    kind = op[0]
    if kind == "all_reduce_max":
        dist.all_reduce(op[1], op=dist.ReduceOp.MAX, group=group)
    elif kind == "all_reduce_sum":
        dist.all_reduce(op[1], op=dist.ReduceOp.SUM, group=group)
    elif kind == "all_to_all":
        dist.all_to_all_single(op[1], op[2], group=group)  # (recv, send), world equal parts
    else:
        raise ValueError(kind)


def run_plans_lockstep(plans):
    """Test / replay helper: drives the routed plans of ALL ranks inside one process (one context per rank on the same device),
    performing each collective among the ranks' tensors with plain tensor operations.  Returns the plans' return values."""
    import torch

    world = len(plans)
    results = [None] * world
    ops = []
    for r, g in enumerate(plans):
        try:
            ops.append(next(g))
        except StopIteration as st:  # a plan without collectives
            results[r] = st.value
            ops.append(None)
    while any(o is not None for o in ops):
        kind = ops[0][0]
        assert all(o is not None and o[0] == kind for o in ops), "ranks disagree on the collective sequence"
        if kind == "p2p_round":  # peer-memory exchange: nothing to perform, the ranks only have to advance stage by stage
            pass
        elif kind in ("all_reduce_max", "all_reduce_sum"):
            stack = torch.stack([o[1] for o in ops])
            red = stack.max(dim=0).values if kind == "all_reduce_max" else stack.sum(dim=0)
            for o in ops:
                o[1].copy_(red)
        else:
            sends = [o[2].reshape(world, -1).clone() for o in ops]
            for q, o in enumerate(ops):
                recv = o[1].reshape(world, -1)
                for src in range(world):
                    recv[src].copy_(sends[src][q])
        nxt = []
        for r, g in enumerate(plans):
            try:
                nxt.append(next(g))
            except StopIteration as st:
                results[r] = st.value
                nxt.append(None)
        ops = nxt
    return results


class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 3}


class ShardedWorld:
    """One rank's view of a world sharded over `world` GPUs.  Shapes are resident on every rank (narrow-phase operands are read by
    global handle); poses are either replicated too (device-resident arm) or travel with the routed records.

    The context's stream is set to torch's current stream here, so the library's kernels and torch.distributed's collectives are
    ordered on ONE stream (both sides enqueue on it); a caller that changes torch's current stream afterwards must call
    ``ctx.lib.ncb_set_stream`` again."""

    def __init__(self, ctx, scene, world, rank, device, mode=None, spatial=None):
        import os

        import torch

        self.ctx, self.scene, self.world, self.rank, self.device = ctx, scene, world, rank, device
        # "p2p" (default): routed ownership, records stored straight into the owner's buffers over NVLink peer memory, no NCCL in the
        #     step (falls back to "routed" on every rank when the peer mapping cannot be set up);
        # "routed": the same ownership with NCCL all-to-all / all-reduce between the stages (NCB_SHARD=routed);
        # "spatial": all-gather + select (NCB_SHARD=spatial); "slices": replicated LBVH (NCB_SHARD=slices)
        if mode is None:
            mode = os.environ.get("NCB_SHARD", "p2p")
            if spatial is not None:
                mode = "spatial" if spatial else "slices"
        assert mode in ("p2p", "routed", "spatial", "slices")
        self.mode = mode
        self.p2p_connected = False
        self.spatial = mode == "spatial"
        self.n = scene.n
        self._route_views = {}
        self.obj_begin, self.obj_end = shard_range(self.n, world, rank)
        self.q_begin, self.q_end = shard_range(self.n, world, rank)
        if torch.device(device).type == "cuda":
            # torch's default stream has handle 0, which ncb_set_stream reads as "the context's own stream": name the legacy
            # default stream explicitly (cudaStreamLegacy == 0x1) so that both sides really share it
            ctx.lib.ncb_set_stream(ctx.h, C.c_void_p(torch.cuda.current_stream(device).cuda_stream or 1))

    # -- peer-memory exchange set-up -------------------------------------------------------------------------
    def p2p_alloc(self):
        """Allocates + exports this rank's receive buffers: (192 bytes of IPC handles, 3 raw device pointers)."""
        import numpy as np

        handles = np.zeros(3 * 64, dtype=np.uint8)
        ptrs = np.zeros(3, dtype=np.uint64)
        self.ctx.check(self.ctx.lib.ncb_route_p2p_alloc(self.ctx.h, C.c_int(self.rank), C.c_int(self.world), C.c_uint32(self.n),
                                                        C.c_void_p(handles.ctypes.data), C.c_void_p(ptrs.ctypes.data)), "ncb_route_p2p_alloc")
        return handles, ptrs

    def connect_p2p(self, group=None):
        """Multi-process set-up (once): all-gathers the IPC handles with torch.distributed and maps the peers' buffers.  Every rank
        learns whether ALL ranks succeeded; otherwise all fall back to the NCCL-routed mode together."""
        import numpy as np
        import torch
        import torch.distributed as dist

        # peers may still have this rank's previous buffers mapped: every rank unmaps first, and only then may anyone reallocate
        self.ctx.lib.ncb_route_p2p_close(self.ctx.h)
        dist.barrier(group=group)
        ok = 1
        try:
            handles, _ = self.p2p_alloc()
        except Exception:  # noqa: BLE001
            handles, ok = np.zeros(3 * 64, dtype=np.uint8), 0
        mine = torch.from_numpy(handles).to(self.device)
        everyone = torch.empty(self.world * 3 * 64, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        if ok:
            all_h = np.ascontiguousarray(everyone.cpu().numpy())
            if self.ctx.lib.ncb_route_p2p_connect(self.ctx.h, C.c_void_p(all_h.ctypes.data), None) != 0:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.p2p_connected = bool(flag.item())
        if not self.p2p_connected:
            self.ctx.lib.ncb_route_p2p_close(self.ctx.h)
            self.mode = "routed"
        return self.p2p_connected

    @staticmethod
    def connect_p2p_local(ranks):
        """All ranks inside ONE process (replay tests): raw device pointers instead of IPC handles."""
        import numpy as np

        ptrs = np.concatenate([sw.p2p_alloc()[1] for sw in ranks]).astype(np.uint64)
        for sw in ranks:
            sw.ctx.check(sw.ctx.lib.ncb_route_p2p_connect(sw.ctx.h, None, C.c_void_p(ptrs.ctypes.data)), "ncb_route_p2p_connect")
            sw.p2p_connected = True
            sw.mode = "p2p"

    def aabb_tensors(self):
        import torch

        lib, h = self.ctx.lib, self.ctx.h
        return [torch.as_tensor(_CudaArray(lib.ncb_device_ptr(h, w), (self.n, 4), "<f4"), device=self.device) for w in (0, 1)]

    def pose_tensors(self):
        import torch

        lib, h = self.ctx.lib, self.ctx.h
        return [
            torch.as_tensor(_CudaArray(lib.ncb_device_ptr(h, 4), (self.n, 3), "<f4"), device=self.device),
            torch.as_tensor(_CudaArray(lib.ncb_device_ptr(h, 5), (self.n, 4), "<f4"), device=self.device),
        ]

    def route_tensor(self, which, typestr="<f4"):
        import torch

        nbytes = C.c_uint64(0)
        p = self.ctx.lib.ncb_route_buffer(self.ctx.h, C.c_int(which), C.c_int(self.world), C.byref(nbytes))
        if not p:
            raise RuntimeError(f"ncb_route_buffer({which}) is not allocated yet")
        key = (which, int(p), nbytes.value, typestr)  # the views are reused from step to step (buffers only move when they grow)
        t = self._route_views.get(which)
        if t is None or t[0] != key:
            t = (key, torch.as_tensor(_CudaArray(p, (nbytes.value // 4,), typestr), device=self.device))
            self._route_views[which] = t
        return t[1]

    def upload_own_poses(self, pos, rot, gather=None):
        """End-to-end input path: this rank uploads the poses of ITS block from host memory.  In the all-gather modes the blocks are
        then all-gathered so that every rank holds every pose; in routed mode the poses travel with the records of the next step
        (``step(..., with_poses=True)``), so nothing else moves here."""
        b, e = self.obj_begin, self.obj_end
        from ._ffi import as_f32, ptr

        pb, rb = as_f32(pos[b:e]), as_f32(rot[b:e])
        self.ctx.check(self.ctx.lib.ncb_set_positions_range(self.ctx.h, C.c_uint32(b), C.c_uint32(e - b), ptr(pb), ptr(rb)), "set_positions_range")
        if gather is None:
            gather = self.mode not in ("routed", "p2p")
        if self.world > 1 and gather:
            for t in self.pose_tensors():
                all_gather_rows(t, b, e, self.world)

    def routed_plan(self, counts_c, with_poses=False):
        """Generator over the collectives of one routed update: yields ("all_reduce_max" | "all_reduce_sum", tensor) or
        ("all_to_all", recv, send); the driver performs the collective among all ranks and resumes.  Returns the counts."""
        lib, h = self.ctx.lib, self.ctx.h
        args = (C.c_float(self.scene.margin), C.c_int(self.rank), C.c_int(self.world), C.c_uint32(self.obj_begin), C.c_uint32(self.obj_end),
                C.c_int(1 if with_poses else 0))
        if self.mode == "p2p" and self.p2p_connected:  # the exchange happens inside the stages; the driver only keeps the ranks in step
            for stage in range(4):
                self.ctx.check(lib.ncb_world_update_routed(h, stage, *args, None), f"routed stage {stage}")
                yield ("p2p_round",)
            self.ctx.check(lib.ncb_world_update_routed(h, 4, *args, C.byref(counts_c)), "routed stage 4")
            return self.ctx._counts(counts_c)
        self.ctx.check(lib.ncb_world_update_routed(h, 0, *args, None), "routed stage 0")
        yield ("all_reduce_max", self.route_tensor(0))
        self.ctx.check(lib.ncb_world_update_routed(h, 1, *args, None), "routed stage 1")
        yield ("all_reduce_sum", self.route_tensor(1, "<i4"))
        for _attempt in range(4):
            self.ctx.check(lib.ncb_world_update_routed(h, 2, *args, None), "routed stage 2")
            yield ("all_to_all", self.route_tensor(3), self.route_tensor(2))
            yield ("all_reduce_max", self.route_tensor(4))
            self.ctx.check(lib.ncb_world_update_routed(h, 3, *args, None), "routed stage 3")
            yield ("all_to_all", self.route_tensor(6), self.route_tensor(5))
            r = lib.ncb_world_update_routed(h, 4, *args, C.byref(counts_c))
            if r != NCB_ROUTE_REPEAT:
                self.ctx.check(r, "routed stage 4")
                return self.ctx._counts(counts_c)
        raise RuntimeError("routed update: bucket capacities did not settle")

    def step(self, counts_c, with_poses=False):
        lib, h, m = self.ctx.lib, self.ctx.h, C.c_float(self.scene.margin)
        if self.mode == "p2p" and self.world > 1:
            if not self.p2p_connected:
                self.connect_p2p()  # first step: one handle exchange; may switch every rank to "routed"
            if self.p2p_connected:  # the whole step in one call: no collective, no host work between the stages
                self.ctx.check(lib.ncb_world_update_routed(h, -1, m, C.c_int(self.rank), C.c_int(self.world), C.c_uint32(self.obj_begin),
                                                           C.c_uint32(self.obj_end), C.c_int(1 if with_poses else 0), C.byref(counts_c)), "routed (p2p)")
                return self.ctx._counts(counts_c)
        if self.mode == "routed" and self.world > 1:
            plan = self.routed_plan(counts_c, with_poses)
            try:
                while True:
                    perform_collective(next(plan))
            except StopIteration as st:
                return st.value
        self.ctx.check(lib.ncb_world_update_stage(h, 0, m, C.c_uint32(self.obj_begin), C.c_uint32(self.obj_end), None), "stage 0")
        if self.world > 1:
            for t in self.aabb_tensors():
                all_gather_rows(t, self.obj_begin, self.obj_end, self.world)
        if self.mode == "spatial" and self.world > 1:
            self.ctx.check(lib.ncb_world_update_sharded(h, m, C.c_int(self.rank), C.c_int(self.world), C.byref(counts_c)), "sharded stage 1")
        else:
            self.ctx.check(lib.ncb_world_update_stage(h, 1, m, C.c_uint32(self.q_begin), C.c_uint32(self.q_end), C.byref(counts_c)), "stage 1")
        return self.ctx._counts(counts_c)
