"""Host-side mirror of the reference's pipeline interface for the hot path, over the C ABI (include/ncb200.h).

  * ``CollisionWorld``   pipeline/world.rs:28-119   (``new(margin)``, ``add``, ``update``, ``contact_pairs``)
  * ``BroadPhase``       pipeline/broad_phase/broad_phase.rs:41-99 (``create_proxy``, ``remove``, ``proxy``,
                         ``deferred_set_bounding_volume``, ``update(handler)`` with interference_started / _stopped,
                         persistent over updates like DBVTBroadPhase)
  * ``NarrowPhase``      contact_generator/contact_manifold_generator.rs:10-36 (batched ``generate_contacts``)
  * ``TriMesh``          shape/trimesh.rs:100-197 + query/ray/ray_trimesh.rs:22-50 (batched ``toi_and_normal_with_ray``)

Same names, argument meaning and error behaviour as the reference where a Python batch API allows it; misuse that
panics in the reference raises here.  Everything computes on the GPU through libncb200.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _ffi
from ._ffi import CONTACT_DTYPE, NcbError, as_f32, as_u32, ptr
from .scenes import DEFAULT_GROUPS, WorldScene
from .shapes import HullLibrary

ALGO_NAMES = ["none", "ball_ball", "plane_ball", "plane_convex", "ball_convex", "convex_convex", "proximity", "capsule_capsule", "capsule_shape"]
ALGO_PROXIMITY = 6


class Proximity:
    """query/proximity/proximity.rs:4-12 (NONE: the pair has no proximity detector / is a contact pair)."""

    Intersecting, WithinMargin, Disjoint, NONE = 0, 1, 2, 255
    NAMES = {0: "Intersecting", 1: "WithinMargin", 2: "Disjoint", 255: "None"}


@dataclass
class UpdateResult:
    pairs: np.ndarray  # [P,2] u32 (object1 = larger handle, object2)
    pair_algo: np.ndarray  # [P] u8 NCB_ALGO_*
    manifold_start: np.ndarray  # [P] u32
    manifold_count: np.ndarray  # [P] u8
    contacts: np.ndarray  # [C] CONTACT_DTYPE
    counts: dict
    proximity: np.ndarray | None = None  # [P] u8 Proximity status (255 for contact pairs); None when the world has no sensor

    def contacts_of(self, p):
        s = int(self.manifold_start[p])
        return self.contacts[s : s + int(self.manifold_count[p])]


class Context:
    """One per GPU (ncb_create)."""

    def __init__(self, device=0):
        self.lib = _ffi.load_library()
        h = C.c_void_p()
        r = self.lib.ncb_create(C.c_int(device), C.byref(h))
        if r != 0:
            raise NcbError(f"ncb_create failed ({r}): {self.lib.ncb_last_error(None).decode()}")
        self.h = h
        self._keep = []

    def check(self, r, what):
        if r < 0:
            raise NcbError(f"{what} failed ({r}): {self.lib.ncb_last_error(self.h).decode()}")
        return r

    def close(self):
        if getattr(self, "h", None):
            self.lib.ncb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- uploads -----------------------------------------------------------------------------------
    def set_hulls(self, lib: HullLibrary):
        hc, keep = _ffi.pack_hull_library(lib)
        self.check(self.lib.ncb_set_hulls(self.h, C.byref(hc)), "ncb_set_hulls")

    def set_objects(self, scene: WorldScene):
        oc, keep = _ffi.pack_objects(scene)
        self.check(self.lib.ncb_set_objects(self.h, C.byref(oc)), "ncb_set_objects")
        self.n = oc.n

    def set_query_types(self, kinds):
        """GeometricQueryType per object: 0 Contacts, 1 Proximity (ncb_set_query_types); None = all Contacts."""
        k = None if kinds is None else np.ascontiguousarray(kinds, dtype=np.uint8)
        self.check(self.lib.ncb_set_query_types(self.h, C.c_uint32(self.n), ptr(k)), "ncb_set_query_types")
        self.has_sensors = k is not None and bool(k.any())

    def set_scene(self, scene: WorldScene):
        self.set_hulls(scene.hulls)
        self.set_objects(scene)
        self.set_query_types(getattr(scene, "query_kind", None))

    def set_positions(self, pos, rot):
        pos, rot = as_f32(pos), as_f32(rot)
        self.check(self.lib.ncb_set_positions(self.h, C.c_uint32(len(pos)), ptr(pos), ptr(rot)), "ncb_set_positions")

    def synchronize(self):
        self.check(self.lib.ncb_synchronize(self.h), "ncb_synchronize")

    def set_kinematics(self, on=True):
        """Ask fresh-world updates / generate_contacts for the ContactKinematic of every contact (``fetch_kinematics``)."""
        self.check(self.lib.ncb_set_kinematics(self.h, C.c_int(1 if on else 0)), "ncb_set_kinematics")

    def fetch_kinematics(self, n_contacts):
        """ContactKinematic records (local1, local2, NeighborhoodGeometry per side, dilations) aligned with the contacts of the last
        update / generate_contacts call."""
        out = np.zeros(n_contacts, dtype=_ffi.KINEMATIC_DTYPE)
        self.check(self.lib.ncb_world_fetch_kinematics(self.h, ptr(out), C.c_uint32(n_contacts)), "ncb_world_fetch_kinematics")
        return out

    def traversal_overflows(self):
        """Query / ray BVH walks that ran out of their fixed stack since the context was created (must be 0)."""
        out = C.c_uint32(0)
        self.check(self.lib.ncb_traversal_overflows(self.h, C.byref(out)), "ncb_traversal_overflows")
        return int(out.value)

    # -- stage entry points ---------------------------------------------------------------------------
    def compute_aabbs(self, margin, mode=2):
        """mode 0: shape AABB, 1: + query_limit (compute_aabb), 2: + margin (what the broad phase stores)."""
        out = np.zeros((self.n, 6), dtype=np.float32)
        self.check(self.lib.ncb_compute_aabbs(self.h, C.c_float(margin), C.c_int(mode), ptr(out)), "ncb_compute_aabbs")
        return out

    def broad_phase(self, aabbs, groups=None):
        aabbs = as_f32(aabbs).reshape(-1, 6)
        n = len(aabbs)
        g = as_u32(groups) if groups is not None else None
        cap = max(8 * n, 1024)
        while True:
            out = np.zeros((cap, 2), dtype=np.uint32)
            npairs = C.c_uint32()
            r = self.check(
                self.lib.ncb_broad_phase(self.h, C.c_uint32(n), ptr(aabbs), ptr(g), ptr(out), C.c_uint32(cap), C.byref(npairs)),
                "ncb_broad_phase",
            )
            if r == 0:
                return out[: npairs.value].copy()
            cap = int(npairs.value) + 16

    def generate_contacts(self, pairs):
        pairs = as_u32(pairs).reshape(-1, 2)
        P = len(pairs)
        cap = max(4 * P, 64)
        start = np.zeros(P, dtype=np.uint32)
        count = np.zeros(P, dtype=np.uint8)
        algo = np.zeros(P, dtype=np.uint8)
        while True:
            out = np.zeros(cap, dtype=CONTACT_DTYPE)
            nc = C.c_uint32()
            r = self.check(
                self.lib.ncb_generate_contacts(self.h, C.c_uint32(P), ptr(pairs), ptr(out), C.c_uint32(cap), C.byref(nc), ptr(start), ptr(count), ptr(algo)),
                "ncb_generate_contacts",
            )
            if r == 0:
                return UpdateResult(pairs, algo, start, count, out[: nc.value].copy(), {"n_contacts": nc.value})
            cap = int(nc.value) + 16

    def proximity(self, pairs, margins=None):
        """ProximityDetector::update with fresh detectors for (object1, object2) pairs (ncb_proximity) -> u8 statuses."""
        pairs = as_u32(pairs).reshape(-1, 2)
        out = np.zeros(len(pairs), dtype=np.uint8)
        m = None if margins is None else as_f32(margins)
        if m is not None and len(m) != len(pairs):
            raise ValueError("one margin per pair")
        self.check(self.lib.ncb_proximity(self.h, C.c_uint32(len(pairs)), ptr(pairs), ptr(m), ptr(out)), "ncb_proximity")
        return out

    def world_fetch_proximity(self, n_pairs):
        out = np.zeros(n_pairs, dtype=np.uint8)
        self.check(self.lib.ncb_world_fetch_proximity(self.h, ptr(out), C.c_uint32(n_pairs)), "ncb_world_fetch_proximity")
        return out

    # -- fused update -----------------------------------------------------------------------------------
    @staticmethod
    def _counts(c):
        return {
            "n_pairs": c.n_pairs,
            "n_contacts": c.n_contacts,
            "n_contact_pairs": c.n_contact_pairs,
            "n_algo": {**{ALGO_NAMES[i]: c.n_algo[i] for i in range(6)}, "proximity": c.n_proximity_pairs,
                       "capsule_capsule": c.n_capsule_pairs[0], "capsule_shape": c.n_capsule_pairs[1]},
            "n_proximity": {"intersecting": c.n_proximity[0], "within_margin": c.n_proximity[1], "disjoint": c.n_proximity[2]},
            "epa_overflow": c.epa_overflow,
            "ref_panics": c.ref_panics,
            "n_epa_pairs": c.n_epa_pairs,
            "n_manifold_jobs": c.n_manifold_jobs,
            "stack_overflow": c.stack_overflow,
            "n_epa_restarts": c.n_epa_restarts,
        }

    def world_update_device(self, margin, q_begin=0, q_end=0xFFFFFFFF):
        c = _ffi.UpdateCountsC()
        self.check(
            self.lib.ncb_world_update_device(self.h, C.c_float(margin), C.c_uint32(q_begin), C.c_uint32(q_end), C.byref(c)),
            "ncb_world_update_device",
        )
        return self._counts(c)

    def world_update_sharded(self, margin, rank, world):
        """AABBs of all objects (stage 0) + ``ncb_world_update_sharded``: the share of rank `rank` of `world`."""
        self.check(self.lib.ncb_world_update_stage(self.h, 0, C.c_float(margin), C.c_uint32(0), C.c_uint32(0xFFFFFFFF), None), "stage 0")
        c = _ffi.UpdateCountsC()
        self.check(self.lib.ncb_world_update_sharded(self.h, C.c_float(margin), C.c_int(rank), C.c_int(world), C.byref(c)), "ncb_world_update_sharded")
        return self._counts(c)

    def world_fetch(self, counts):
        P, Cn = counts["n_pairs"], counts["n_contacts"]
        pairs = np.zeros((P, 2), dtype=np.uint32)
        algo = np.zeros(P, dtype=np.uint8)
        start = np.zeros(P, dtype=np.uint32)
        count = np.zeros(P, dtype=np.uint8)
        contacts = np.zeros(Cn, dtype=CONTACT_DTYPE)
        self.check(
            self.lib.ncb_world_fetch(self.h, ptr(pairs), C.c_uint32(P), ptr(algo), ptr(start), ptr(count), ptr(contacts), C.c_uint32(Cn)),
            "ncb_world_fetch",
        )
        prox = self.world_fetch_proximity(P) if getattr(self, "has_sensors", False) else None
        return UpdateResult(pairs, algo, start, count, contacts, counts, prox)

    def world_update(self, scene: WorldScene, bufs=None):
        """The end-to-end host-buffer call (ncb_world_update): uploads the objects, runs the step, copies results back."""
        oc, keep = _ffi.pack_objects(scene)
        n = oc.n
        qk = getattr(scene, "query_kind", None)
        if qk is not None or getattr(self, "has_sensors", False):
            # query types belong to the object set: install the set first so that they can be attached to it
            self.check(self.lib.ncb_set_objects(self.h, C.byref(oc)), "ncb_set_objects")
            self.n = n
            self.set_query_types(qk)
        self.n = n
        if bufs is None:
            bufs = self.alloc_result_buffers(max(8 * n, 1024), max(8 * n, 1024))
        while True:
            c = _ffi.UpdateCountsC()
            r = self.check(
                self.lib.ncb_world_update(
                    self.h, C.byref(oc), C.c_float(scene.margin), ptr(bufs["pairs"]), C.c_uint32(len(bufs["pairs"])), ptr(bufs["algo"]),
                    ptr(bufs["start"]), ptr(bufs["count"]), ptr(bufs["contacts"]), C.c_uint32(len(bufs["contacts"])), C.byref(c),
                ),
                "ncb_world_update",
            )
            if r == 0:
                P, Cn = c.n_pairs, c.n_contacts
                prox = self.world_fetch_proximity(P) if getattr(self, "has_sensors", False) else None
                return UpdateResult(bufs["pairs"][:P], bufs["algo"][:P], bufs["start"][:P], bufs["count"][:P], bufs["contacts"][:Cn], self._counts(c), prox)
            bufs = self.alloc_result_buffers(c.n_pairs + 16, c.n_contacts + 16)

    def world_update_poses(self, pos, rot, margin, bufs=None):
        """``set_position`` on every object + ``CollisionWorld::update`` (world.rs:104-119) for a world whose objects are already on
        the device (``set_objects``): uploads the poses only, runs the step, copies the results back (ncb_world_update_poses)."""
        pos, rot = as_f32(pos).reshape(-1, 3), as_f32(rot).reshape(-1, 4)
        n = len(pos)
        if bufs is None:
            bufs = self.alloc_result_buffers(max(8 * n, 1024), max(8 * n, 1024))
        while True:
            c = _ffi.UpdateCountsC()
            r = self.check(
                self.lib.ncb_world_update_poses(
                    self.h, C.c_uint32(n), ptr(pos), ptr(rot), C.c_float(margin), ptr(bufs["pairs"]), C.c_uint32(len(bufs["pairs"])),
                    ptr(bufs["algo"]), ptr(bufs["start"]), ptr(bufs["count"]), ptr(bufs["contacts"]), C.c_uint32(len(bufs["contacts"])), C.byref(c),
                ),
                "ncb_world_update_poses",
            )
            if r == 0:
                P, Cn = c.n_pairs, c.n_contacts
                prox = self.world_fetch_proximity(P) if getattr(self, "has_sensors", False) else None
                return UpdateResult(bufs["pairs"][:P], bufs["algo"][:P], bufs["start"][:P], bufs["count"][:P], bufs["contacts"][:Cn], self._counts(c), prox)
            bufs = self.alloc_result_buffers(c.n_pairs + 16, c.n_contacts + 16)

    @staticmethod
    def alloc_result_buffers(cap_pairs, cap_contacts):
        return {
            "pairs": np.zeros((cap_pairs, 2), dtype=np.uint32),
            "algo": np.zeros(cap_pairs, dtype=np.uint8),
            "start": np.zeros(cap_pairs, dtype=np.uint32),
            "count": np.zeros(cap_pairs, dtype=np.uint8),
            "contacts": np.zeros(cap_contacts, dtype=CONTACT_DTYPE),
        }

    def profile_enable(self, on=True):
        self.lib.ncb_profile_enable(self.h, C.c_int(1 if on else 0))

    def profile_get(self):
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        launches = (C.c_uint32 * 32)()
        n = self.lib.ncb_profile_get(self.h, names, ms, launches)
        return [(names[i].decode(), float(ms[i]), int(launches[i])) for i in range(max(n, 0))]

    def trimesh(self, verts, tris):
        return TriMesh(self, verts, tris)


class GeometricQueryType:
    """pipeline/object/query_type.rs:8-37: Contacts(linear, angular) or Proximity(margin) (a sensor)."""

    @staticmethod
    def Contacts(linear, angular):
        return ("contacts", float(linear), float(angular))

    @staticmethod
    def Proximity(margin):
        return ("proximity", float(margin), 0.0)


class CollisionWorld:
    """pipeline/world.rs: ``CollisionWorld::new(margin)``, ``add``, ``update``, ``contact_pairs``.

    Objects are appended to host SoA arrays by ``add`` (handle = insertion index, as the slab gives on a fresh
    world); ``update`` runs one fresh-world step on the device and returns / stores the result.
    """

    def __init__(self, margin, device=0, ctx=None):
        self.margin = float(margin)
        self.ctx = ctx or Context(device)
        self._pos, self._rot, self._type, self._param, self._groups, self._ql, self._ang = [], [], [], [], [], [], []
        self._kind = []
        self._hulls = []
        self._hull_ids = {}
        self.result = None
        self._dirty = True

    def add(self, position, shape, groups=None, query_type=None, data=None):
        """position = (translation xyz, quaternion ijkw); returns the object handle."""
        if query_type is None or query_type[0] not in ("contacts", "proximity"):
            raise ValueError("query_type must be GeometricQueryType.Contacts(..) or GeometricQueryType.Proximity(..)")
        if query_type[1] < 0:
            raise ValueError("The proximity margin / contact prediction must be positive or null.")  # the reference asserts it
        t, q = position
        self._pos.append(np.asarray(t, dtype=np.float32))
        self._rot.append(np.asarray(q, dtype=np.float32))
        self._type.append(shape.type_id)
        if shape.type_id == 2:
            key = id(shape)
            if key not in self._hull_ids:
                self._hull_ids[key] = len(self._hulls)
                self._hulls.append(shape)
            self._param.append(np.array([self._hull_ids[key], 0, 0, 0], dtype=np.float32))
        else:
            self._param.append(shape.param())
        self._groups.append(np.asarray(groups if groups is not None else DEFAULT_GROUPS, dtype=np.uint32))
        self._ql.append(query_type[1])
        self._ang.append(query_type[2])
        self._kind.append(1 if query_type[0] == "proximity" else 0)
        self._dirty = True
        return len(self._pos) - 1

    def scene(self):
        n = len(self._pos)
        return WorldScene(
            pos=np.array(self._pos, dtype=np.float32).reshape(n, 3),
            rot=np.array(self._rot, dtype=np.float32).reshape(n, 4),
            shape_type=np.array(self._type, dtype=np.uint32),
            shape_param=np.array(self._param, dtype=np.float32).reshape(n, 4),
            groups=np.array(self._groups, dtype=np.uint32).reshape(n, 3),
            query_limit=np.array(self._ql, dtype=np.float32),
            ang_pred=np.array(self._ang, dtype=np.float32),
            hulls=HullLibrary(self._hulls),
            margin=self.margin,
            query_kind=np.array(self._kind, dtype=np.uint8) if any(self._kind) else None,
        )

    def update(self):
        s = self.scene()
        if self._dirty:
            self.ctx.set_hulls(s.hulls)
            self._dirty = False
        self.result = self.ctx.world_update(s)
        return self.result

    def contact_pairs(self, effective_only=True):
        """Iterator of (handle1, handle2, algorithm, contacts) — world.rs:452-470."""
        r = self.result
        if r is None:
            return
        for p in range(len(r.pairs)):
            if r.pair_algo[p] == 0 or r.pair_algo[p] == ALGO_PROXIMITY:
                continue
            c = r.contacts_of(p)
            # InteractionGraph::is_interaction_effective (interaction_graph.rs:390-398): deepest contact with depth >= 0
            if effective_only and (len(c) == 0 or not (c["depth"].max() >= 0)):
                continue
            yield int(r.pairs[p, 0]), int(r.pairs[p, 1]), ALGO_NAMES[r.pair_algo[p]], c


    def proximity_pairs(self, effective_only=True):
        """Iterator of (handle1, handle2, status) over the proximity interactions — world.rs:433-445 (`proximity_pairs`);
        effective_only keeps the Intersecting pairs only (interaction_graph.rs:399)."""
        r = self.result
        if r is None or r.proximity is None:
            return
        for p in np.nonzero(r.pair_algo == ALGO_PROXIMITY)[0]:
            st = int(r.proximity[p])
            if effective_only and st != Proximity.Intersecting:
                continue
            yield int(r.pairs[p, 0]), int(r.pairs[p, 1]), st

    def proximity_events(self):
        """ProximityEvents of the (fresh-world) update: (collider1, collider2, prev_status = Disjoint, new_status) for every
        proximity pair whose status is not Disjoint (narrow_phase.rs:108-121; a new interaction starts as Disjoint)."""
        return [(a, b, Proximity.Disjoint, st) for a, b, st in self.proximity_pairs(effective_only=False) if st != Proximity.Disjoint]


class BroadPhaseInterferenceHandler:
    """pipeline/broad_phase/broad_phase.rs:28-38"""

    def is_interference_allowed(self, a, b):
        return True

    def interference_started(self, a, b):
        pass

    def interference_stopped(self, a, b):
        pass


class BroadPhase:
    """The ``BroadPhase`` trait (pipeline/broad_phase/broad_phase.rs:68-99) over the device persistent broad phase
    (``ncb_bp_*``, csrc/bp_persistent.cu), a drop-in for ``DBVTBroadPhase::new(margin)``:
    ``create_proxy`` / ``remove`` / ``deferred_set_bounding_volume`` / ``update(handler)`` / ``proxy`` /
    ``num_interferences`` with the reference's handle recycling, loosening rule and handler argument order.

    Collision groups (``groups=(membership, whitelist, blacklist)`` per proxy) are filtered on the device; an
    arbitrary ``handler.is_interference_allowed`` is applied on the host to the started events (its answer must be
    stable for a pair, as it is for the reference's collision-world handler)."""

    def __init__(self, margin, ctx=None, device=0):
        self.margin = np.float32(margin)
        self.ctx = ctx or Context(device)
        self._lib = self.ctx.lib
        h = C.c_void_p()
        self.ctx.check(self._lib.ncb_bp_create(self.ctx.h, C.c_float(float(margin)), C.byref(h)), "ncb_bp_create")
        self._h = h
        self._data = {}
        self._groups = np.zeros((0, 3), dtype=np.uint32)
        self._any_groups = False
        self._vetoed = set()
        self._pending_set = ([], [])

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ncb_bp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- proxies ---------------------------------------------------------------------------------------------
    def create_proxies(self, bvs, datas=None, groups=None):
        self._flush_sets()
        bvs = as_f32(bvs).reshape(-1, 6)
        n = len(bvs)
        out = np.zeros(n, dtype=np.uint32)
        self.ctx.check(self._lib.ncb_bp_create_proxies(self._h, C.c_uint32(n), ptr(bvs), ptr(out)), "ncb_bp_create_proxies")
        top = int(out.max()) + 1 if n else 0
        if top > len(self._groups):
            g = np.zeros((max(top, 2 * len(self._groups)), 3), dtype=np.uint32)
            g[:, 0] = g[:, 1] = 0x3FFFFFFF  # CollisionGroups::new(): member of all, whitelist all (collision_groups.rs:39-46)
            g[: len(self._groups)] = self._groups
            self._groups = g
        self._groups[out, 0] = self._groups[out, 1] = 0x3FFFFFFF
        self._groups[out, 2] = 0
        if groups is not None:
            self._groups[out] = as_u32(groups).reshape(n, 3)
            self._any_groups = True
        for k, hnd in enumerate(out.tolist()):
            self._data[hnd] = datas[k] if datas is not None else hnd
        return out

    def create_proxy(self, bv, data=None, groups=None):
        g = None if groups is None else np.asarray(groups, dtype=np.uint32).reshape(1, 3)
        return int(self.create_proxies(np.asarray(bv, dtype=np.float32).reshape(1, 6), [data], g)[0])

    def remove(self, handles, removal_handler=None):
        self._flush_sets()
        handles = as_u32(handles).reshape(-1)
        n = C.c_uint32()
        self.ctx.check(self._lib.ncb_bp_remove(self._h, C.c_uint32(len(handles)), ptr(handles), C.byref(n)), "ncb_bp_remove")
        ev = np.zeros((n.value, 2), dtype=np.uint32)
        if n.value:
            self.ctx.check(self._lib.ncb_bp_events(self._h, None, ptr(ev)), "ncb_bp_events")
        for a, b in ev.tolist():
            if (a, b) in self._vetoed:
                self._vetoed.discard((a, b))
            elif removal_handler is not None:
                removal_handler(self._data[a], self._data[b])
        for hnd in handles.tolist():
            self._data.pop(hnd, None)
        return ev

    def proxy(self, handle):
        self._flush_sets()
        mm = np.zeros(6, dtype=np.float32)
        r = self._lib.ncb_bp_proxy(self._h, C.c_uint32(int(handle)), ptr(mm))
        self.ctx.check(min(r, 0), "ncb_bp_proxy")
        return (mm, self._data[int(handle)]) if r == 1 else None

    def deferred_set_bounding_volume(self, handle, bv):
        # queued on the host and sent as one batch (order kept) before the next call that needs it
        self._pending_set[0].append(int(handle))
        self._pending_set[1].append(np.asarray(bv, dtype=np.float32).reshape(6))

    def deferred_set_bounding_volumes(self, handles, bvs):
        self._flush_sets()
        handles, bvs = as_u32(handles).reshape(-1), as_f32(bvs).reshape(-1, 6)
        self.ctx.check(self._lib.ncb_bp_set_bounding_volumes(self._h, C.c_uint32(len(handles)), ptr(handles), ptr(bvs)),
                       "ncb_bp_set_bounding_volumes")

    def deferred_recompute_all_proximities_with(self, handles):
        self._flush_sets()
        handles = as_u32(np.atleast_1d(handles)).reshape(-1)
        self.ctx.check(self._lib.ncb_bp_recompute_with(self._h, C.c_uint32(len(handles)), ptr(handles)), "ncb_bp_recompute_with")

    def deferred_recompute_all_proximities(self):
        self._flush_sets()
        self.ctx.check(self._lib.ncb_bp_recompute_all(self._h), "ncb_bp_recompute_all")

    def set_groups(self, handle, groups):
        """Changes the collision groups of a live proxy and re-queues it (glue/update.rs:83-86)."""
        self._groups[int(handle)] = np.asarray(groups, dtype=np.uint32).reshape(3)
        self._any_groups = True
        self.deferred_recompute_all_proximities_with([handle])

    def _flush_sets(self):
        hs, bs = self._pending_set
        if hs:
            self._pending_set = ([], [])
            self.deferred_set_bounding_volumes(np.array(hs, dtype=np.uint32), np.stack(bs))

    def num_interferences(self):
        n = C.c_uint32()
        self.ctx.check(self._lib.ncb_bp_num_interferences(self._h, C.byref(n)), "ncb_bp_num_interferences")
        return n.value - len(self._vetoed)

    def pairs(self):
        n = self.num_interferences() + len(self._vetoed)
        out = np.zeros((n, 2), dtype=np.uint32)
        cnt = C.c_uint32()
        self.ctx.check(self._lib.ncb_bp_pairs(self._h, ptr(out), C.c_uint32(n), C.byref(cnt)), "ncb_bp_pairs")
        if self._vetoed:
            keep = [tuple(p) not in self._vetoed for p in out.tolist()]
            out = out[np.array(keep, dtype=bool)]
        return out

    def _query(self, kind, q, width):
        self._flush_sets()
        q = as_f32(q).reshape(-1, width)
        cap = max(64 * len(q), 4096)
        while True:
            out = np.zeros((cap, 2), dtype=np.uint32)
            n = C.c_uint32()
            r = self.ctx.check(self._lib.ncb_bp_query(self._h, C.c_int(kind), C.c_uint32(len(q)), ptr(q), ptr(out), C.c_uint32(cap), C.byref(n)),
                               "ncb_bp_query")
            if r == 0:
                return out[: n.value]
            cap = n.value

    def interferences_with_bounding_volumes(self, bvs):
        """Batched ``interferences_with_bounding_volume``: [K,2] rows (query index, handle), sorted."""
        return self._query(0, bvs, 6)

    def interferences_with_rays(self, origins, dirs, max_toi):
        o, d = as_f32(origins).reshape(-1, 3), as_f32(dirs).reshape(-1, 3)
        t = np.broadcast_to(np.asarray(max_toi, dtype=np.float32).reshape(-1, 1), (len(o), 1))
        return self._query(1, np.concatenate([o, d, t], axis=1), 7)

    def interferences_with_points(self, points):
        return self._query(2, points, 3)

    def interferences_with_bounding_volume(self, bv):
        return [self._data[h] for h in self.interferences_with_bounding_volumes(np.asarray(bv).reshape(1, 6))[:, 1].tolist()]

    def interferences_with_ray(self, origin, direction, max_toi):
        return [self._data[h] for h in self.interferences_with_rays(origin, direction, max_toi)[:, 1].tolist()]

    def interferences_with_point(self, point):
        return [self._data[h] for h in self.interferences_with_points(np.asarray(point).reshape(1, 3))[:, 1].tolist()]

    def update_events(self):
        """One ``update``; returns (started[k,2], stopped[m,2]) handle arrays without going through a handler."""
        self._flush_sets()
        ns, nst = C.c_uint32(), C.c_uint32()
        g = self._groups if self._any_groups else None
        self.ctx.check(self._lib.ncb_bp_update(self._h, ptr(g), C.c_uint32(0 if g is None else len(g)), C.byref(ns), C.byref(nst)),
                       "ncb_bp_update")
        started = np.zeros((ns.value, 2), dtype=np.uint32)
        stopped = np.zeros((nst.value, 2), dtype=np.uint32)
        if ns.value or nst.value:
            self.ctx.check(self._lib.ncb_bp_events(self._h, ptr(started) if ns.value else None, ptr(stopped) if nst.value else None),
                           "ncb_bp_events")
        return started, stopped

    def update(self, handler):
        started, stopped = self.update_events()
        for a, b in started.tolist():
            if handler.is_interference_allowed(self._data[a], self._data[b]):
                handler.interference_started(self._data[a], self._data[b])
            else:
                self._vetoed.add((min(a, b), max(a, b)))
        for a, b in stopped.tolist():
            if (a, b) in self._vetoed:
                self._vetoed.discard((a, b))
            else:
                handler.interference_stopped(self._data[a], self._data[b])


class SteppingWorld:
    """``CollisionWorld`` stepped over time (``set_position`` on some objects, then ``update``) with the reference's temporal
    coherence: persistent broad phase, pairs in callback orientation, GJK warm start, manifold cache with stable contact
    ids, contact events (``ncb_sim_*``, csrc/sim.cu).  The object set is the scene given at construction."""

    def __init__(self, ctx: Context, scene: WorldScene):
        self.ctx = ctx
        ctx.set_scene(scene)
        ctx.n = scene.n
        h = C.c_void_p()
        ctx.check(ctx.lib.ncb_sim_create(ctx.h, C.c_float(scene.margin), C.byref(h)), "ncb_sim_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.ctx.lib.ncb_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_positions(self, handles, pos, rot):
        hs = None if handles is None else as_u32(handles).reshape(-1)
        p, r = as_f32(pos).reshape(-1, 3), as_f32(rot).reshape(-1, 4)
        self.ctx.check(self.ctx.lib.ncb_sim_set_positions(self._h, C.c_uint32(len(p)), ptr(hs), ptr(p), ptr(r)), "ncb_sim_set_positions")

    def set_collision_groups(self, handles, groups):
        """``CollisionObject::set_collision_groups`` on live objects (groups[k] = membership, whitelist, blacklist): the next update
        redispatches them in the broad phase and updates their pairs (``ncb_sim_set_collision_groups``)."""
        hs, g = as_u32(handles).reshape(-1), as_u32(groups).reshape(-1, 3)
        if len(hs) != len(g):
            raise ValueError("one (membership, whitelist, blacklist) row per handle")
        self.ctx.check(self.ctx.lib.ncb_sim_set_collision_groups(self._h, C.c_uint32(len(hs)), ptr(hs), ptr(g)), "ncb_sim_set_collision_groups")

    def remove(self, handles):
        """``CollisionWorld::remove``: the objects and their pairs disappear (no events); handles are recycled by ``add``."""
        hs = as_u32(handles).reshape(-1)
        self.ctx.check(self.ctx.lib.ncb_sim_remove(self._h, C.c_uint32(len(hs)), ptr(hs)), "ncb_sim_remove")

    def add(self, scene: WorldScene):
        """``CollisionWorld::add`` for every object of `scene` (same hull library as the world); returns their handles."""
        oc, keep = _ffi.pack_objects(scene)
        out = np.zeros(scene.n, dtype=np.uint32)
        qk = getattr(scene, "query_kind", None)
        qk = None if qk is None else np.ascontiguousarray(qk, dtype=np.uint8)
        self.ctx.check(self.ctx.lib.ncb_sim_add_with_query_types(self._h, C.byref(oc), ptr(qk), ptr(out)), "ncb_sim_add_with_query_types")
        self.ctx.n = max(self.ctx.n, int(out.max()) + 1) if len(out) else self.ctx.n
        return out

    def update(self, fetch=True):
        """One ``CollisionWorld::update``.  Returns dict(pairs, algo, off, contacts, ids, events, counts)."""
        c = _ffi.UpdateCountsC()
        self.ctx.check(self.ctx.lib.ncb_sim_step(self._h, C.byref(c)), "ncb_sim_step")
        counts = self.ctx._counts(c)
        if not fetch:
            return {"counts": counts}
        P, Cn, E = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self.ctx.check(self.ctx.lib.ncb_sim_sizes(self._h, C.byref(P), C.byref(Cn), C.byref(E)), "ncb_sim_sizes")
        P, Cn, E = P.value, Cn.value, E.value
        pairs = np.zeros((P, 2), dtype=np.uint32)
        algo = np.zeros(P, dtype=np.uint8)
        start = np.zeros(P, dtype=np.uint32)
        count = np.zeros(P, dtype=np.uint8)
        contacts = np.zeros(Cn, dtype=CONTACT_DTYPE)
        ids = np.zeros(Cn, dtype=np.uint32)
        events = np.zeros((E, 3), dtype=np.uint32)
        self.ctx.check(self.ctx.lib.ncb_sim_fetch(self._h, ptr(pairs), ptr(algo), ptr(start), ptr(count), ptr(contacts), ptr(ids), ptr(events)),
                       "ncb_sim_fetch")
        off = np.concatenate([start, [Cn]]).astype(np.uint32) if P else np.zeros(1, dtype=np.uint32)
        # proximity side (sensors): status per pair + ProximityEvents (collider1, collider2, prev, new)
        prox = np.full(P, 255, dtype=np.uint8)
        ne = C.c_uint32()
        self.ctx.check(self.ctx.lib.ncb_sim_fetch_proximity(self._h, ptr(prox) if P else None, None, C.c_uint32(0), C.byref(ne)), "ncb_sim_fetch_proximity")
        pev = np.zeros((ne.value, 4), dtype=np.uint32)
        if ne.value:
            self.ctx.check(self.ctx.lib.ncb_sim_fetch_proximity(self._h, None, ptr(pev), C.c_uint32(ne.value), C.byref(ne)), "ncb_sim_fetch_proximity")
        return {"pairs": pairs, "algo": algo, "off": off, "count": count, "contacts": contacts, "ids": ids, "events": events, "counts": counts,
                "prox": prox, "prox_events": pev}

    step = update

    def query(self, kind, q, groups=None):
        """kind 0: ``interferences_with_aabb`` (q[n,6]); kind 2: ``interferences_with_point`` (q[n,3]) (glue/query.rs:79-181).
        Returns rows (query, handle), sorted."""
        q = as_f32(q).reshape(-1, 6 if kind == 0 else 3)
        g = None if groups is None else as_u32(groups).reshape(3)
        cap = max(16 * len(q), 4096)
        while True:
            idx = np.zeros((cap, 2), dtype=np.uint32)
            n = C.c_uint32()
            r = self.ctx.check(self.ctx.lib.ncb_sim_query(self._h, C.c_int(kind), C.c_uint32(len(q)), ptr(q), ptr(g), ptr(idx), C.c_uint32(cap),
                                                          C.byref(n)), "ncb_sim_query")
            if r == 0:
                return idx[: n.value]
            cap = n.value

    def ray_cast(self, origins, dirs, max_toi, groups=None, first_only=False):
        """``interferences_with_ray`` / ``first_interference_with_ray`` for a batch of rays (glue/query.rs:13-77,183-224).
        Returns (idx[K,2] (ray, handle), toi[K], normal[K,3], feature[K]), rows sorted by (ray, handle)."""
        o, d = as_f32(origins).reshape(-1, 3), as_f32(dirs).reshape(-1, 3)
        t = np.broadcast_to(np.asarray(max_toi, dtype=np.float32).reshape(-1, 1), (len(o), 1))
        rays = np.ascontiguousarray(np.concatenate([o, d, t], axis=1), dtype=np.float32)
        g = None if groups is None else as_u32(groups).reshape(3)
        cap = len(o) if first_only else max(4 * len(o), 4096)
        while True:
            idx = np.empty((cap, 2), dtype=np.uint32)
            val = np.empty((cap, 4), dtype=np.float32)
            feat = np.empty(cap, dtype=np.uint32)
            n = C.c_uint32()
            r = self.ctx.check(self.ctx.lib.ncb_sim_ray_cast(self._h, C.c_uint32(len(o)), ptr(rays), ptr(g), C.c_int(int(first_only)), ptr(idx), ptr(val),
                                                             ptr(feat), C.c_uint32(cap), C.byref(n)), "ncb_sim_ray_cast")
            if r == 0:
                k = n.value
                return idx[:k], val[:k, 0].copy(), val[:k, 1:4].copy(), feat[:k]
            cap = n.value


class TriMesh:
    """``TriMesh::new(points, indices, None)`` + batched ``RayCast::toi_and_normal_with_ray``."""

    def __init__(self, ctx: Context, verts, tris):
        self.ctx = ctx
        self.verts = as_f32(verts).reshape(-1, 3)
        self.tris = as_u32(tris).reshape(-1, 3)
        h = C.c_void_p()
        ctx.check(
            ctx.lib.ncb_trimesh_create(ctx.h, C.c_uint32(len(self.verts)), ptr(self.verts), C.c_uint32(len(self.tris)), ptr(self.tris), C.byref(h)),
            "ncb_trimesh_create",
        )
        self.h = h
        self.n_tris = len(self.tris)

    def set_uvs(self, uvs):
        """``TriMesh::new(points, indices, Some(uvs))``: per-vertex texture coordinates (None clears them)."""
        u = as_f32(uvs).reshape(-1, 2) if uvs is not None else None
        if u is not None and len(u) != len(self.verts):
            raise ValueError("one uv per vertex")
        self.ctx.check(self.ctx.lib.ncb_trimesh_set_uvs(self.h, ptr(u)), "ncb_trimesh_set_uvs")
        self.has_uvs = u is not None

    def toi_and_normal_with_ray(self, pose, origins, dirs, max_toi=None, want_normals=True, out=None):
        """pose: None or 7 floats (t xyz, q ijkw).  max_toi: None, one value, or one per ray.  `out`: optional dict of preallocated
        (e.g. pinned) result arrays "toi" / "face" / "normal".  Returns (toi [-1 = None], face [i or i+T], normals)."""
        toi, face, normal, _ = self._cast(pose, origins, dirs, max_toi, want_normals, False, out)
        return toi, face, normal

    def toi_and_normal_and_uv_with_ray(self, pose, origins, dirs, max_toi=None, out=None):
        """``RayCast::toi_and_normal_and_uv_with_ray`` (ray_trimesh.rs:52-94): also the interpolated uv of the hit point."""
        return self._cast(pose, origins, dirs, max_toi, True, True, out)

    def _cast(self, pose, origins, dirs, max_toi, want_normals, want_uv, out):
        o, d = as_f32(origins).reshape(-1, 3), as_f32(dirs).reshape(-1, 3)
        n = len(o)
        out = out or {}
        toi = out.get("toi") if out.get("toi") is not None else np.zeros(n, dtype=np.float32)
        face = out.get("face") if out.get("face") is not None else np.zeros(n, dtype=np.uint32)
        normal = (out.get("normal") if out.get("normal") is not None else np.zeros((n, 3), dtype=np.float32)) if want_normals else None
        uv = (out.get("uv") if out.get("uv") is not None else np.zeros((n, 2), dtype=np.float32)) if want_uv else None
        p = as_f32(pose) if pose is not None else None
        per_ray = None
        if max_toi is None:
            max_toi = np.finfo(np.float32).max
        elif np.ndim(max_toi) > 0:
            per_ray = as_f32(max_toi).reshape(-1)
            if len(per_ray) != n:
                raise ValueError("one max_toi per ray")
            max_toi = 0.0
        self.ctx.check(
            self.ctx.lib.ncb_trimesh_ray_cast_uv(self.h, ptr(p), C.c_uint32(n), ptr(o), ptr(d), C.c_float(max_toi), ptr(per_ray), ptr(toi), ptr(face),
                                                 ptr(normal), ptr(uv)),
            "ncb_trimesh_ray_cast_uv",
        )
        return toi, face, normal, uv

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.ncb_trimesh_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
