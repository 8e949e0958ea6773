"""Differential fuzzing of the ray primitives of csrc/ray.cu compiled for the host (tests/host_shim/ray_host.cpp) against the oracle's
brute-force mode: TriMesh (slab_toi + ray_triangle, the kernel's hit rule and epilogue) and ncollide2d Polyline (slab_toi2 + segment_ray2),
random meshes / polylines, poses and per-ray limits, for a wall-clock budget.  CPU only.
python scripts/fuzz_rays_host_shim.py [seconds] [seed0]  ->  one JSON summary line."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from ncollide_b200.scenes import make_polyline_scene, make_ray_scene, transform_rays  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402
from test_device_source_on_host import _build_shim  # noqa: E402

F = np.float32


def vp(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def bits(a):
    return np.ascontiguousarray(a, dtype=F).view(np.uint32)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 95_000
    shim, orc = _build_shim("libray_host.so", "ray_host.cpp"), Oracle()
    t0, seed = time.time(), seed0
    tot = dict(rounds=0, trimesh_rays=0, trimesh_hits=0, trimesh_differences=0, polyline_rays=0, polyline_hits=0, polyline_differences=0)
    bad = []
    fmax = float(np.finfo(F).max)
    while time.time() - t0 < budget:
        rng = np.random.default_rng(seed)
        kind = ("terrain", "soup")[seed % 2]
        posed = bool(seed % 3 == 0)
        rs = make_ray_scene(kind, int(rng.integers(200, 3000)), 1500, seed=seed, random_pose=posed)
        pose = np.ascontiguousarray(rs.pose, dtype=F) if posed else None
        o, d = (rs.origins, rs.dirs) if not posed else transform_rays(rs.pose, rs.origins, rs.dirs)
        max_toi = fmax if seed % 2 else float(rng.uniform(3.0, 15.0))
        n = len(o)
        toi, face, nrm = np.zeros(n, dtype=F), np.zeros(n, dtype=np.uint32), np.zeros((n, 3), dtype=F)
        shim.shim_trimesh_ray_cast(C.c_uint32(len(rs.tris)), vp(rs.verts), vp(rs.tris), vp(pose), C.c_uint64(n), vp(o), vp(d), C.c_float(max_toi),
                                   vp(toi), vp(face), vp(nrm))
        ot, of, on = orc.trimesh(rs.verts, rs.tris).ray_cast(o, d, max_toi=max_toi, pose=pose, mode=1)
        hit = ot >= 0
        ok = np.array_equal(toi >= 0, hit) and np.array_equal(face[hit], of[hit]) and np.array_equal(bits(toi[hit]), bits(ot[hit])) and \
            np.array_equal(bits(nrm[hit]), bits(on[hit]))
        tot["trimesh_rays"] += n
        tot["trimesh_hits"] += int(hit.sum())
        if not ok:
            tot["trimesh_differences"] += 1
            bad.append(("trimesh", seed))
        pkind = ("terrain", "soup")[(seed // 2) % 2]
        pts, edges, po, pd = make_polyline_scene(pkind, int(rng.integers(50, 3000)), 1500, seed)
        if edges is None:
            edges = np.stack([np.arange(len(pts) - 1), np.arange(1, len(pts))], axis=1).astype(np.uint32)
        pose2 = None
        if seed % 3 == 1:
            ang = rng.uniform(-np.pi, np.pi)
            pose2 = np.array([rng.uniform(-3, 3), rng.uniform(-3, 3), np.cos(ang), np.sin(ang)], dtype=F)
            rot = np.array([[pose2[2], -pose2[3]], [pose2[3], pose2[2]]], dtype=np.float64)
            po = np.ascontiguousarray(po.astype(np.float64) @ rot.T + pose2[:2].astype(np.float64), dtype=F)
            pd = np.ascontiguousarray(pd.astype(np.float64) @ rot.T, dtype=F)
        limits = rng.uniform(1.0, 14.0, size=len(po)).astype(F) if seed % 2 == 0 else None
        n = len(po)
        toi, feat, nrm = np.zeros(n, dtype=F), np.zeros(n, dtype=np.uint32), np.zeros((n, 2), dtype=F)
        shim.shim2_polyline_ray_cast(C.c_uint32(len(edges)), vp(pts), vp(edges), vp(pose2), C.c_uint64(n), vp(po), vp(pd), C.c_float(fmax), vp(limits),
                                     vp(toi), vp(feat), vp(nrm))
        ot, of, on = orc.polyline(pts, edges).ray_cast(po, pd, max_toi=limits, pose=pose2, mode=1)
        ok = np.array_equal(bits(toi), bits(ot)) and np.array_equal(feat, of) and np.array_equal(bits(nrm), bits(on))
        tot["polyline_rays"] += n
        tot["polyline_hits"] += int((ot >= 0).sum())
        if not ok:
            tot["polyline_differences"] += 1
            bad.append(("polyline", seed))
        tot["rounds"] += 1
        seed += 1
    tot.update(mismatches=bad[:20], seconds=round(time.time() - t0, 1), seed0=seed0)
    print(json.dumps(tot))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
