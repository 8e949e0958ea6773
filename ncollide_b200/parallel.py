"""Multi-GPU plumbing for the world update: one process per GPU, objects
block-partitioned for the AABB computation, fat AABBs all-gathered (NCCL on GPU; gloo in the CPU tests), then every rank
selects the objects it owns spatially (equal-count Morton ranges) plus the ghosts around them, builds its own LBVH over
those and runs the pair search + narrow phase of its share of the pairs.  torch.distributed is plumbing only."""
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
        dist.all_gather_into_tensor(full, full[begin:end].clone(), group=group)
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


class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 3}


class ShardedWorld:
    """One rank's view of a world sharded over `world` GPUs.  All objects (poses, shapes) are resident on every rank;
    per step each rank computes the AABBs of its block, all-gathers them, builds the replicated LBVH and processes
    its slice of the Morton order."""

    def __init__(self, ctx, scene, world, rank, device, spatial=None):
        import os

        self.ctx, self.scene, self.world, self.rank, self.device = ctx, scene, world, rank, device
        # spatial = True: ownership by Morton range + ghosts, local LBVH (ncb_world_update_sharded);
        # spatial = False: replicated LBVH, query slices of the Morton order (the first design; NCB_SHARD=slices)
        self.spatial = (os.environ.get("NCB_SHARD", "spatial") != "slices") if spatial is None else spatial
        self.n = scene.n
        self.obj_begin, self.obj_end = shard_range(self.n, world, rank)
        self.q_begin, self.q_end = shard_range(self.n, world, rank)

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

    def upload_own_poses(self, pos, rot):
        """End-to-end input path: this rank uploads the poses of ITS block from host memory, then the blocks are
        all-gathered so that every rank holds every pose (narrow-phase operands are read from the replicated arrays)."""
        b, e = self.obj_begin, self.obj_end
        from ._ffi import ptr

        self.ctx.check(self.ctx.lib.ncb_set_positions_range(self.ctx.h, C.c_uint32(b), C.c_uint32(e - b), ptr(pos[b:e]), ptr(rot[b:e])), "set_positions_range")
        if self.world > 1:
            for t in self.pose_tensors():
                all_gather_rows(t, b, e, self.world)

    def step(self, counts_c):
        lib, h, m = self.ctx.lib, self.ctx.h, C.c_float(self.scene.margin)
        self.ctx.check(lib.ncb_world_update_stage(h, 0, m, C.c_uint32(self.obj_begin), C.c_uint32(self.obj_end), None), "stage 0")
        if self.world > 1:
            for t in self.aabb_tensors():
                all_gather_rows(t, self.obj_begin, self.obj_end, self.world)
        if self.spatial and self.world > 1:
            self.ctx.check(lib.ncb_world_update_sharded(h, m, C.c_int(self.rank), C.c_int(self.world), C.byref(counts_c)), "sharded stage 1")
        else:
            self.ctx.check(lib.ncb_world_update_stage(h, 1, m, C.c_uint32(self.q_begin), C.c_uint32(self.q_end), C.byref(counts_c)), "stage 1")
        return self.ctx._counts(counts_c)
