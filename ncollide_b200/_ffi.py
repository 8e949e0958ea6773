"""ctypes binding of libncb200.so (include/ncb200.h).  No fallback: if the CUDA library is missing or no GPU is
usable, calls fail loudly."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libncb200.so")

EXPORTED_SYMBOLS = [
    "ncb_version", "ncb_create", "ncb_destroy", "ncb_last_error", "ncb_set_stream", "ncb_get_stream", "ncb_synchronize", "ncb_traversal_overflows", "ncb_set_kinematics", "ncb_world_fetch_kinematics",
    "ncb_set_hulls", "ncb_set_objects", "ncb_set_positions", "ncb_set_positions_range", "ncb_compute_aabbs", "ncb_broad_phase", "ncb_generate_contacts",
    "ncb_world_update_device", "ncb_world_fetch", "ncb_world_update", "ncb_world_update_poses", "ncb_device_ptr", "ncb_world_update_stage", "ncb_world_update_sharded", "ncb_world_update_routed", "ncb_route_buffer", "ncb_route_p2p_alloc", "ncb_route_p2p_connect", "ncb_route_p2p_close", "ncb_world_fetch_early",
    "ncb_profile_enable", "ncb_profile_get", "ncb_trimesh_create", "ncb_trimesh_destroy", "ncb_trimesh_ray_cast",
    "ncb_trimesh_ray_cast_device", "ncb_trimesh_ray_cast_uv", "ncb_trimesh_set_uvs", "ncb2d_contact", "ncb2d_world_update", "ncb2d_proximity", "ncb2d_polyline_create", "ncb2d_polyline_destroy", "ncb2d_polyline_ray_cast",
    "ncb2d_polyline_ray_cast_device", "ncb2d_ray_cast", "ncb2d_world_fetch_proximity", "ncb2d_world_ray_cast", "ncb2d_world_query",
    "ncb_bp_create", "ncb_bp_destroy", "ncb_bp_create_proxies", "ncb_bp_set_bounding_volumes", "ncb_bp_remove", "ncb_bp_update",
    "ncb_sim_create", "ncb_sim_destroy", "ncb_sim_set_positions", "ncb_sim_set_collision_groups", "ncb_sim_step", "ncb_sim_sizes", "ncb_sim_fetch", "ncb_sim_remove", "ncb_sim_add", "ncb_sim_ray_cast", "ncb_sim_query",
    "ncb_set_query_types", "ncb_proximity", "ncb_world_fetch_proximity", "ncb_sim_fetch_proximity", "ncb_sim_add_with_query_types",
    "ncb_bp_events", "ncb_bp_query", "ncb_bp_recompute_with", "ncb_bp_recompute_all", "ncb_bp_num_interferences", "ncb_bp_pairs", "ncb_bp_proxy",
]

HULL_FIELDS = (
    "vert_off face_off edge_off fadj_off vadj_off points vert_first_adj vert_num_adj face_first face_num face_normal "
    "vertices_adj_to_face edges_adj_to_face edge_vertices edge_faces edge_dir faces_adj_to_vertex edges_adj_to_vertex"
).split()
_FLOAT_HULL_FIELDS = {"points", "face_normal", "edge_dir"}


class HullLibraryC(C.Structure):
    _fields_ = [("n_hulls", C.c_uint32)] + [(n, C.c_void_p) for n in HULL_FIELDS]


class ObjectsC(C.Structure):
    _fields_ = [
        ("n", C.c_uint32),
        ("pos", C.c_void_p),
        ("rot", C.c_void_p),
        ("shape_type", C.c_void_p),
        ("shape_param", C.c_void_p),
        ("groups", C.c_void_p),
        ("query_limit", C.c_void_p),
        ("ang_pred", C.c_void_p),
    ]


class UpdateCountsC(C.Structure):
    _fields_ = [
        ("n_pairs", C.c_uint32),
        ("n_contacts", C.c_uint32),
        ("n_contact_pairs", C.c_uint32),
        ("n_algo", C.c_uint32 * 6),
        ("epa_overflow", C.c_uint32),
        ("ref_panics", C.c_uint32),
        ("n_epa_pairs", C.c_uint32),
        ("n_manifold_jobs", C.c_uint32),
        ("n_proximity_pairs", C.c_uint32),
        ("n_proximity", C.c_uint32 * 3),
        ("n_capsule_pairs", C.c_uint32 * 2),
        ("stack_overflow", C.c_uint32),
        ("n_epa_restarts", C.c_uint32),
    ]


CONTACT_DTYPE = np.dtype(
    [("world1", np.float32, 3), ("world2", np.float32, 3), ("normal", np.float32, 3), ("depth", np.float32),
     ("f1", np.uint32), ("f2", np.uint32), ("pair", np.uint32)]
)
assert CONTACT_DTYPE.itemsize == 52
KINEMATIC_DTYPE = np.dtype(
    [("local1", np.float32, 3), ("local2", np.float32, 3), ("dir1", np.float32, 3), ("dir2", np.float32, 3),
     ("dil1", np.float32), ("dil2", np.float32), ("g1", np.uint32), ("g2", np.uint32)]
)
assert KINEMATIC_DTYPE.itemsize == 64

_lib = None


class NcbError(RuntimeError):
    pass


def load_library():
    """Loads libncb200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NcbError(
            f"{LIB_PATH} is missing: build it with `python -m ncollide_b200.build` (nvcc, sm_100a). "
            "ncollide_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.ncb_version.restype = C.c_char_p
    lib.ncb_last_error.restype = C.c_char_p
    lib.ncb_last_error.argtypes = [C.c_void_p]
    lib.ncb_get_stream.restype = C.c_void_p
    lib.ncb_device_ptr.restype = C.c_void_p
    lib.ncb_device_ptr.argtypes = [C.c_void_p, C.c_int]
    lib.ncb_route_buffer.restype = C.c_void_p
    lib.ncb_route_buffer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    lib.ncb_destroy.argtypes = [C.c_void_p]
    lib.ncb_trimesh_destroy.argtypes = [C.c_void_p]
    lib.ncb_bp_destroy.argtypes = [C.c_void_p]
    lib.ncb_bp_destroy.restype = None
    lib.ncb_sim_destroy.argtypes = [C.c_void_p]
    lib.ncb_sim_destroy.restype = None
    _lib = lib
    return lib


def ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def as_f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


def as_u32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a if shape is None else a.reshape(shape)


def pack_hull_library(lib):
    """lib: shapes.HullLibrary -> (HullLibraryC, keepalive list)"""
    keep = []
    h = HullLibraryC()
    h.n_hulls = lib.n_hulls
    for f in HULL_FIELDS:
        a = np.ascontiguousarray(getattr(lib, f), dtype=np.float32 if f in _FLOAT_HULL_FIELDS else np.uint32)
        keep.append(a)
        setattr(h, f, a.ctypes.data)
    return h, keep


def pack_objects(scene):
    keep = [
        as_f32(scene.pos), as_f32(scene.rot), as_u32(scene.shape_type), as_f32(scene.shape_param),
        as_u32(scene.groups) if scene.groups is not None else None, as_f32(scene.query_limit), as_f32(scene.ang_pred),
    ]
    o = ObjectsC()
    o.n = len(keep[0])
    o.pos, o.rot, o.shape_type, o.shape_param = (k.ctypes.data for k in keep[:4])
    o.groups = keep[4].ctypes.data if keep[4] is not None else None
    o.query_limit, o.ang_pred = keep[5].ctypes.data, keep[6].ctypes.data
    return o, keep
