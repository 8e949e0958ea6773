"""Differential fuzzing of the device against the oracle: many small random worlds with varied sizes, densities, margins,
prediction distances, angular predictions, collision groups and degenerate placements, for a wall-clock budget.
python scripts/fuzz_parity.py [seconds] [seed0]  ->  one JSON summary line; a mismatch dumps the scene to gpurun_out/."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from ncollide_b200.scenes import make_world_scene  # noqa: E402
from ncollide_b200.world import Context  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

RTOL, ATOL = 1e-4, 1e-5
F = np.float32


def canon(p):
    p = np.sort(np.asarray(p, dtype=np.uint32).reshape(-1, 2), axis=1)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


def random_scene(rng, seed):
    n = int(rng.integers(2, 600))
    kinds = [(1, 1, 1), (0, 1, 1), (1, 0, 1), (0, 0, 1), (0, 1, 0), (1, 1, 0)][rng.integers(0, 6)]
    dens = rng.choice([0.6, 1.0, 2.0, 4.0])
    side = max(0.5, (n ** (1 / 3)) * 1.1 / dens)
    s = make_world_scene(n, seed, kinds, side=side, n_hulls=int(rng.integers(1, 24)), plane=bool(rng.random() < 0.3),
                         linear=float(rng.choice([0.0, 0.002, 0.02, 0.2])), angular=float(rng.choice([0.0, 0.0, 0.01, 0.1, 0.5])),
                         margin=float(rng.choice([0.0, 0.01, 0.02, 0.1])), name=f"fuzz{seed}")
    mode = rng.integers(0, 6)
    if mode == 0:  # snap positions to a coarse grid: exact coincidences / touching faces
        s.pos[:] = (np.round(s.pos * 2) / 2).astype(F)
    elif mode == 1:  # axis-aligned rotations
        s.rot[:] = (0, 0, 0, 1)
    elif mode == 2:  # scale shapes
        k = F(rng.choice([0.05, 0.3, 3.0]))
        s.shape_param[:, :3] = (s.shape_param[:, :3] * F(k)).astype(F)
        hull = s.shape_type == 2
        s.shape_param[hull, 0] = np.round(s.shape_param[hull, 0] / F(k))  # hull ids are not lengths
    elif mode == 3:  # far from the origin
        s.pos[:] = (s.pos + rng.uniform(-3000, 3000, size=3).astype(F)).astype(F)
    if rng.random() < 0.3:  # random collision groups
        m = len(s.groups)
        s.groups[:, 0] = (1 << rng.integers(0, 4, size=m)).astype(np.uint32)
        s.groups[:, 1] = rng.integers(1, 16, size=m).astype(np.uint32)
        s.groups[:, 2] = np.where(rng.random(m) < 0.2, 1 << rng.integers(0, 4, size=m), 0).astype(np.uint32)
    if os.environ.get("FUZZ_SENSORS", "1") != "0" and rng.random() < 0.4:  # some objects are Proximity sensors
        s.query_kind = (rng.random(s.n) < rng.choice([0.1, 0.5, 1.0])).astype(np.uint8)
        if not s.query_kind.any():
            s.query_kind = None
        else:
            s.query_limit = np.where(s.query_kind != 0, F(rng.choice([0.0, 0.05, 0.3])), s.query_limit).astype(F)
    return s


def compare(res, s, orc):
    want = orc.broad_phase(orc.compute_aabbs(s), s.groups, mode=1)
    if not np.array_equal(canon(res.pairs), canon(want)):
        return "pair set"
    if getattr(s, "query_kind", None) is not None:
        oc, ooff, oalgo, oprox = orc.narrow_phase_kinds(s, res.pairs)
        if res.proximity is None or not np.array_equal(res.proximity, oprox):
            return "proximity statuses"
    else:
        oc, ooff, oalgo, _ = orc.narrow_phase(s, res.pairs)
    if not np.array_equal(res.pair_algo, oalgo):
        return "algo"
    if not np.array_equal(res.manifold_count, np.diff(ooff)):
        return "manifold sizes"
    idx = np.concatenate([np.arange(st, st + c) for st, c in zip(res.manifold_start, res.manifold_count)]) if len(oc) else np.zeros(0, int)
    dc = res.contacts[idx.astype(np.int64)]
    if not (np.array_equal(dc["f1"], oc["f1"]) and np.array_equal(dc["f2"], oc["f2"])):
        return "feature ids"
    for name in ("world1", "world2", "normal", "depth"):
        if not np.allclose(dc[name], oc[name], rtol=RTOL, atol=ATOL):
            return name
    if res.counts["epa_overflow"] or res.counts["ref_panics"]:
        return f"counters {res.counts['epa_overflow']} {res.counts['ref_panics']}"
    return None


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    ctx, orc = Context(0), Oracle()
    t0 = time.time()
    n_scenes = n_pairs = n_contacts = 0
    bad = []
    seed = seed0
    while time.time() - t0 < budget:
        rng = np.random.default_rng(seed)
        s = random_scene(rng, seed)
        ctx.set_hulls(s.hulls)
        res = ctx.world_update(s)
        why = compare(res, s, orc)
        n_scenes += 1
        n_pairs += len(res.pairs)
        n_contacts += len(res.contacts)
        if why:
            bad.append((seed, why))
            os.makedirs("gpurun_out", exist_ok=True)
            np.savez_compressed(f"gpurun_out/fuzz_bad_{seed}.npz", pos=s.pos, rot=s.rot, shape_type=s.shape_type, shape_param=s.shape_param,
                                groups=s.groups, query_limit=s.query_limit, ang_pred=s.ang_pred, margin=s.margin,
                                query_kind=s.query_kind if s.query_kind is not None else np.zeros(0, np.uint8))
            if len(bad) >= 10:
                break
        seed += 1
    print(json.dumps({"scenes": n_scenes, "pairs": n_pairs, "contacts": n_contacts, "mismatches": bad, "seconds": round(time.time() - t0, 1),
                      "seed0": seed0}))


if __name__ == "__main__":
    main()
