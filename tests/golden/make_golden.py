"""Generates the golden fixtures in this directory from the CPU oracle.

The reference is Rust and cannot be built or imported in this image (no rustc/cargo, nalgebra not vendored), so these
vectors are NOT outputs of the reference binary: they are outputs of the oracle (oracle/), which is itself pinned on
the reference's own known-answer tests (tests/test_oracle_kat.py).  They guard the oracle and the device against drift.

    python -m tests.golden.make_golden
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ncollide_b200.scenes import WorldScene, config_scene, make_ray_scene, make_world_scene  # noqa: E402
from ncollide_b200.shapes import ConvexHull, HullLibrary  # noqa: E402

HULL_FIELDS = HullLibrary.FIELDS


def scene_to_dict(s):
    d = {k: getattr(s, k) for k in ("pos", "rot", "shape_type", "shape_param", "groups", "query_limit", "ang_pred")}
    if getattr(s, "query_kind", None) is not None:
        d["query_kind"] = s.query_kind
    d["margin"] = np.float32(s.margin)
    d["hull_n"] = np.uint32(s.hulls.n_hulls)
    for f in HULL_FIELDS:
        d["hull_" + f] = getattr(s.hulls, f)
    return d


class _Lib:
    FIELDS = HULL_FIELDS


def scene_from_npz(z):
    lib = _Lib()
    lib.n_hulls = int(z["hull_n"])
    for f in HULL_FIELDS:
        setattr(lib, f, z["hull_" + f])
    lib.max_verts = 0
    return WorldScene(
        pos=z["pos"], rot=z["rot"], shape_type=z["shape_type"], shape_param=z["shape_param"], groups=z["groups"],
        query_limit=z["query_limit"], ang_pred=z["ang_pred"], hulls=lib, margin=float(z["margin"]),
        query_kind=z["query_kind"] if "query_kind" in z else None,
    )


def make_proximity(o):
    """prox_mixed_plane_400.npz: a 400-object world with a plane and 35 % Proximity sensors — the fresh-world
    narrow phase (algorithm, manifold sizes, contacts, proximity status per pair), the batched detectors with random explicit
    margins, and 5 stepping-world updates (statuses + ProximityEvents per step)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ncollide_b200.scenes import with_sensors
    from sim_scenario import drive

    s = with_sensors(make_world_scene(400, 1020, (1, 1, 1), side=4.4, n_hulls=10, plane=True), 0.35, 1021, margin=0.1)
    d = scene_to_dict(s)
    fat = o.compute_aabbs(s)
    pairs = o.broad_phase(fat, s.groups, 0)
    c, off, algo, prox = o.narrow_phase_kinds(s, pairs)
    d.update(fat_aabbs=fat, pairs=pairs, manifold_off=off, algo=algo, prox=prox)
    for n in ("world1", "world2", "normal", "depth", "f1", "f2"):
        d["c_" + n] = c[n]
    rng = np.random.default_rng(1022)
    bp = rng.integers(0, s.n, size=(3000, 2)).astype(np.uint32)
    bp = bp[bp[:, 0] != bp[:, 1]]
    bm = rng.uniform(0, 1.5, size=len(bp)).astype(np.float32)
    d.update(batch_pairs=bp, batch_margins=bm, batch_prox=o.proximity(s, bp, bm))
    log = drive(o.sim(s), s, steps=5, seed=1020)
    for t, r in enumerate(log):
        for k in ("pairs", "algo", "off", "prox", "prox_events", "events"):
            d[f"s{t}_{k}"] = r[k]
    np.savez_compressed(os.path.join(HERE, "prox_mixed_plane_400.npz"), **d)
    print("prox", s.n, "objects", len(pairs), "pairs", int((algo == 6).sum()), "proximity pairs", np.bincount(prox[algo == 6], minlength=3).tolist(),
          "statuses;", [len(r["prox_events"]) for r in log], "proximity events per step")


def make_capsules(o):
    """capsule_mixed_plane_300.npz (SURVEY §8f N3 groundwork): balls / cuboids / capsules and a plane — fat AABBs, pairs, the two capsule
    generators' manifolds.  Oracle output like everything else here; the capsule part of the oracle is checked against closed-form
    geometry only (tests/test_oracle_capsule.py), the reference has no capsule test."""
    rng = np.random.default_rng(1030)
    n = 300
    s = make_world_scene(n, 1030, (1, 1, 0), side=4.2, plane=True)
    cap = np.arange(n) % 3 == 2
    s.shape_type[:n][cap] = 4
    s.shape_param[:n][cap, 0] = rng.uniform(0.2, 0.5, size=int(cap.sum())).astype(np.float32)
    s.shape_param[:n][cap, 1] = rng.uniform(0.15, 0.3, size=int(cap.sum())).astype(np.float32)
    s.shape_param[:n][cap, 2] = 0
    d = scene_to_dict(s)
    fat = o.compute_aabbs(s)
    pairs = o.broad_phase(fat, s.groups, 0)
    c, off, algo, _ = o.narrow_phase(s, pairs)
    d.update(fat_aabbs=fat, pairs=pairs, manifold_off=off, algo=algo)
    for k in ("world1", "world2", "normal", "depth", "f1", "f2"):
        d["c_" + k] = c[k]
    np.savez_compressed(os.path.join(HERE, "capsule_mixed_plane_300.npz"), **d)
    print("capsules", s.n, "objects", len(pairs), "pairs", int((algo >= 7).sum()), "capsule pairs", len(c), "contacts")


def make_dim2(o):
    """ncollide2d: a 600-object world of all five shape kinds with two planes and sensors (pairs, manifolds, features, proximity
    statuses), 400 world rays (all / first), 400 shape ray casts and a 300-edge polyline with 400 rays."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_dim2 import random_world
    from test_rays2d import random_shape_rays

    from ncollide_b200.scenes import make_polyline_scene

    w = random_world(600, 2201, (0, 1, 2, 4), angular=0.05, planes=2, with_groups=True)
    w.set_sensors(np.random.default_rng(2202).random(w.n) < 0.2)
    pairs, off, c, feats, panics, fat = o.world_update2d(w)
    d = dict(pos=w.pos, rot=w.rot, type=w.type, param=w.param, points=w.points, normals=w.normals, query_limit=w.query_limit, ang_pred=w.ang_pred,
             groups=w.groups, query_kind=w.query_kind, margin=np.float32(w.margin), fat=fat, pairs=pairs, off=off, contacts=c, feats=feats,
             prox=o.last_proximity2d, panics=np.uint32(panics))
    rng = np.random.default_rng(2203)
    lo, hi = w.pos.min(axis=0), w.pos.max(axis=0)
    dirs = rng.normal(size=(400, 2))
    rays = np.concatenate([rng.uniform(lo, hi, size=(400, 2)), dirs / np.linalg.norm(dirs, axis=1, keepdims=True), rng.uniform(1.0, 8.0, size=(400, 1))],
                          axis=1).astype(np.float32)
    idx, val, ft = o.world_ray_cast2d(w, rays)
    idx1, val1, ft1 = o.world_ray_cast2d(w, rays, first_only=True)
    d.update(q_rays=rays, q_idx=idx, q_val=val, q_feat=ft, q_first_idx=idx1, q_first_val=val1, q_first_feat=ft1)
    typ, par, pose, srays, spts = random_shape_rays(400, 2204, kinds=(0, 1, 2, 3, 4))
    f, out, sf = o.ray_cast2d(typ, par, pose, srays, spts)
    d.update(s_type=typ, s_param=par, s_pose=pose, s_rays=srays, s_points=spts, s_found=f, s_out=out, s_feat=sf)
    pts, edges, po, pd = make_polyline_scene("terrain", 300, 400, 2205)
    toi, pf, pn = o.polyline(pts, edges).ray_cast(po, pd, mode=0)
    d.update(p_points=pts, p_origins=po, p_dirs=pd, p_toi=toi, p_feat=pf, p_normal=pn)
    np.savez_compressed(os.path.join(HERE, "dim2_world_600.npz"), **d)
    print("dim2", w.n, "objects", len(pairs), "pairs", len(c), "contacts", int((o.last_proximity2d != 255).sum()), "sensor pairs", len(idx), "ray rows")


def main():
    from oracle.pyoracle import Oracle

    o = Oracle()
    if len(sys.argv) > 1 and sys.argv[1] == "dim2":  # only the 2-D fixture (the others stay as they are)
        make_dim2(o)
        return
    scenes = {
        "world_cfg1_balls_300": config_scene(1, 300),
        "world_cfg2_mixed_plane_400": config_scene(2, 400),
        "world_cfg3_mixed_hulls_500": make_world_scene(500, 1003, (1, 1, 1), side=5.5, n_hulls=12, angular=0.02),
    }
    for name, s in scenes.items():
        d = scene_to_dict(s)
        fat = o.compute_aabbs(s)
        pairs = o.broad_phase(fat, s.groups, 0)
        c, off, algo, _ = o.narrow_phase(s, pairs)
        d.update(fat_aabbs=fat, pairs=pairs, manifold_off=off, algo=algo)
        for n in ("world1", "world2", "normal", "depth", "f1", "f2"):
            d["c_" + n] = c[n]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, s.n, "objects", len(pairs), "pairs", len(c), "contacts")
    # stepping world: 5 updates of a 400-object world driven by tests/sim_scenario.drive, then world queries
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from sim_scenario import drive

    s = make_world_scene(400, 1010, (1, 1, 1), side=4.6, n_hulls=10, plane=True)
    sim = o.sim(s)
    log = drive(sim, s, steps=5, seed=1010)
    d = scene_to_dict(s)
    for t, r in enumerate(log):
        for k in ("pairs", "algo", "off", "ids", "events"):
            d[f"s{t}_{k}"] = r[k]
        for n in ("world1", "world2", "normal", "depth", "f1", "f2"):
            d[f"s{t}_c_{n}"] = r["contacts"][n]
    rng = np.random.default_rng(1011)
    ro = rng.uniform(-1, 5.6, size=(300, 3)).astype(np.float32)
    rd = rng.normal(size=(300, 3)).astype(np.float32)
    idx, toi, normal, feat = sim.ray_cast(ro, rd, 30.0)
    idx1, toi1, normal1, feat1 = sim.ray_cast(ro, rd, 30.0, first_only=True)
    pts = rng.uniform(0, 4.6, size=(300, 3)).astype(np.float32)
    d.update(q_ro=ro, q_rd=rd, q_idx=idx, q_toi=toi, q_normal=normal, q_feat=feat, q_first_idx=idx1, q_first_toi=toi1, q_pts=pts,
             q_point_rows=sim.query(2, pts))
    np.savez_compressed(os.path.join(HERE, "sim_mixed_plane_400.npz"), **d)
    print("sim", [len(r["pairs"]) for r in log], "pairs per step,", len(idx), "ray hits")
    make_proximity(o)
    make_capsules(o)
    make_dim2(o)
    for kind in ("terrain", "soup"):
        rs = make_ray_scene(kind, 2000, 600, seed=1004)
        om = o.trimesh(rs.verts, rs.tris)
        toi, face, normal = om.ray_cast(rs.origins, rs.dirs, mode=0)
        np.savez_compressed(os.path.join(HERE, f"rays_{kind}_2000.npz"), verts=rs.verts, tris=rs.tris, origins=rs.origins, dirs=rs.dirs,
                            toi=toi, face=face, normal=normal)
        print(kind, (toi >= 0).sum(), "hits")


if __name__ == "__main__":
    main()
