"""Differential fuzzing of the 2-D DEVICE SOURCE (csrc/dim2.cu compiled for the host through tests/host_shim) against the oracle, for a
wall-clock budget: query::contact, query::proximity, shape ray casts, contains_point on random (shape, pose) pairs of all five kinds, and
fresh 2-D worlds (fat boxes, manifolds, features, sensors).  CPU only.
python scripts/fuzz_dim2_host_shim.py [seconds] [seed0]  ->  one JSON summary line."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle.pyoracle import Oracle  # noqa: E402
from test_device_source_on_host import _build_shim  # noqa: E402
from test_dim2 import random_pairs, random_world  # noqa: E402
from test_rays2d import random_shape_rays  # noqa: E402

F = np.float32


def vp(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def bits(a):
    return np.ascontiguousarray(a, dtype=F).view(np.uint32)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 90_000
    shim, orc = _build_shim("libdim2_host.so", "dim2_host.cpp"), Oracle()
    shim.shim2_narrow_sensors.restype = C.c_uint64
    t0, seed = time.time(), seed0
    tot = dict(rounds=0, contact_pairs=0, contacts_found=0, contact_words_not_bit_exact=0, proximity_pairs=0, proximity_differences=0, rays=0,
               ray_differences=0, points=0, point_differences=0, worlds=0, world_pairs=0, world_contacts=0, world_words_not_bit_exact=0)
    bad = []
    while time.time() - t0 < budget:
        rng = np.random.default_rng(seed)
        kinds = [(0, 1, 2, 3, 4), (1, 2, 4), (0, 4), (2,), (0, 1, 2)][seed % 5]
        pred = float(rng.choice([0.0, 0.02, 0.1, 0.3]))
        t1, p1, m1, t2, p2, m2, pts, nrm = random_pairs(20000, seed, kinds, spread=float(rng.choice([0.6, 1.2])))
        n = len(t1)
        found, out, flags = np.zeros(n, dtype=np.uint8), np.zeros((n, 7), dtype=F), np.zeros(2, dtype=np.uint32)
        shim.shim2_contact(C.c_uint64(n), vp(t1), vp(p1), vp(m1), vp(t2), vp(p2), vp(m2), vp(pts), vp(nrm), C.c_float(pred), vp(found), vp(out), vp(flags))
        of, oo, opan = orc.contact2d(t1, p1, m1, t2, p2, m2, pts, pred, poly_normals=nrm)
        hit = found.astype(bool)
        diff = int((bits(out[hit]) != bits(oo[hit])).sum()) if np.array_equal(found, of) else -1
        tot["contact_pairs"] += n
        tot["contacts_found"] += int(hit.sum())
        tot["contact_words_not_bit_exact"] += max(diff, 0)
        if diff != 0 or flags[0] != opan:
            bad.append(("contact", seed, diff))
        mg = rng.uniform(0.0, 0.4, size=n).astype(F)
        st = np.full(n, 9, dtype=np.uint8)
        shim.shim2_proximity(C.c_uint64(n), vp(t1), vp(p1), vp(m1), vp(t2), vp(p2), vp(m2), vp(pts), vp(mg), vp(st))
        d = int((st != orc.proximity2d(t1, p1, m1, t2, p2, m2, pts, mg)).sum())
        tot["proximity_pairs"] += n
        tot["proximity_differences"] += d
        if d:
            bad.append(("proximity", seed, d))
        typ, par, pose, rays, rp = random_shape_rays(20000, seed, kinds=(0, 1, 2, 3, 4))
        f, o, ft = np.zeros(len(typ), dtype=np.uint8), np.zeros((len(typ), 3), dtype=F), np.zeros(len(typ), dtype=np.uint32)
        shim.shim2_ray_cast(C.c_uint64(len(typ)), vp(typ), vp(par), vp(pose), vp(rp), vp(rays), vp(f), vp(o), vp(ft))
        rf, ro, rft = orc.ray_cast2d(typ, par, pose, rays, rp)
        h = f.astype(bool)
        d = 0 if (np.array_equal(f, rf) and np.array_equal(ft, rft) and np.array_equal(bits(o[h]), bits(ro[h]))) else 1
        tot["rays"] += len(typ)
        tot["ray_differences"] += d
        if d:
            bad.append(("rays", seed, d))
        q = (pose[:, :2] + rng.normal(size=(len(typ), 2)) * 0.4).astype(F)
        ins = np.zeros(len(typ), dtype=np.uint8)
        shim.shim2_contains_point(C.c_uint64(len(typ)), vp(typ), vp(par), vp(pose), vp(rp), vp(q), vp(ins))
        d = int((ins.astype(bool) != orc.contains_point2d(typ, par, pose, q, rp)).sum())
        tot["points"] += len(typ)
        tot["point_differences"] += d
        if d:
            bad.append(("points", seed, d))
        w = random_world(int(rng.integers(200, 2500)), seed, [(0, 1, 2, 4), (1, 2), (0, 1, 2)][seed % 3], angular=float(rng.choice([0.0, 0.05, 0.3])),
                         planes=int(rng.integers(0, 3)), density=float(rng.choice([1.5, 2.5, 4.0])))
        if seed % 2:
            w.set_sensors(rng.random(w.n) < 0.2)
        pairs, off, oc, ofe, pan, fat = orc.world_update2d(w)
        oprox = orc.last_proximity2d
        boxes = np.zeros((w.n, 6), dtype=F)
        shim.shim2_aabbs(C.c_uint32(w.n), vp(w.pos), vp(w.rot), vp(w.type), vp(w.param), vp(w.query_limit), vp(w.points), vp(w.normals),
                         C.c_float(w.margin), vp(boxes))
        P = len(pairs)
        pr = np.ascontiguousarray(pairs, dtype=np.uint32)
        doff, dc, df = np.zeros(P + 1, dtype=np.uint32), np.zeros((4 * P + 16, 7), dtype=F), np.zeros((4 * P + 16, 2), dtype=np.uint32)
        fl, prox = np.zeros(3, dtype=np.uint32), np.full(P, 255, dtype=np.uint8)
        nc = shim.shim2_narrow_sensors(C.c_uint32(w.n), vp(w.pos), vp(w.rot), vp(w.type), vp(w.param), vp(w.query_limit), vp(w.ang_pred), vp(w.points),
                                       vp(w.normals), C.c_uint64(P), vp(pr), vp(doff), vp(dc), vp(df), C.c_uint64(len(dc)), vp(fl), vp(w.query_kind), vp(prox))
        ok = (np.array_equal(bits(boxes), bits(fat)) and np.array_equal(doff, off) and nc == len(oc) and np.array_equal(df[:nc], ofe)
              and np.array_equal(prox, oprox) and fl[0] == pan)
        wd = int((bits(dc[:nc]) != bits(oc)).sum()) if ok else -1
        tot["worlds"] += 1
        tot["world_pairs"] += P
        tot["world_contacts"] += int(nc)
        tot["world_words_not_bit_exact"] += max(wd, 0)
        if wd != 0:
            bad.append(("world", seed, wd))
        tot["rounds"] += 1
        seed += 1
    tot.update(mismatches=bad[:20], seconds=round(time.time() - t0, 1), seed0=seed0)
    print(json.dumps(tot))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
