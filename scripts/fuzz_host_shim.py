"""Differential fuzzing WITHOUT a GPU: the per-pair device source (proximity detectors of csrc/proximity.cu, GJK / EPA of
csrc/gjk.cuh, the contact generators / clipping / manifold of csrc/narrow.cu), compiled for the host through tests/host_shim/, against the oracle on many small random worlds (sizes, densities,
margins, degenerate placements, scales, far-away coordinates: the generator of scripts/fuzz_parity.py).
python scripts/fuzz_host_shim.py [seconds] [seed0]  ->  one JSON summary line.  Test infrastructure only."""
import copy
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
sys.path.insert(0, "scripts")
os.environ["FUZZ_SENSORS"] = "0"
from fuzz_parity import random_scene  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402
from sim_scenario import step_poses  # noqa: E402
from test_device_source_on_host import (ShimEdges, _build_shim, _sorted_rows, compare_narrow, shim_contact_sm_sm, shim_narrow_phase,  # noqa: E402
                                        shim_narrow_phase_capsules, shim_proximity)


def persist_rounds(nar, orc, s, cb, rng, steps=4):
    """A fixed edge set over `steps` updates (stepping-world per-pair state).  Returns (edge updates, None) or (.., reason)."""
    t = s.shape_type
    cb = cb[~((t[cb[:, 0]] == 3) & (t[cb[:, 1]] == 3))]
    if len(cb) == 0:
        return 0, None
    dev, ref = ShimEdges(nar, cb), orc.edges(cb)
    n = 0
    for step in range(steps):
        if step == 0:
            which = np.arange(len(cb), dtype=np.uint32)
        else:
            idx = step_poses(s, s.pos, s.rot, rng, 0.5)
            moved = np.zeros(s.n, dtype=bool)
            moved[idx] = True
            which = np.nonzero(moved[cb[:, 0]] | moved[cb[:, 1]])[0].astype(np.uint32)
        ed, eo = dev.update(s, which), ref.update(s, which)
        n += len(which)
        if dev.flags[0] or dev.flags[1] or dev.overflow[0]:
            return n, f"step {step}: overflow / panic {dev.flags[:2].tolist()} {int(dev.overflow[0])}"
        if not np.array_equal(_sorted_rows(ed), _sorted_rows(eo)):
            return n, f"step {step}: events"
        (dc, doff, dids, ddir), (oc, ooff, oids, odir) = dev.fetch(), ref.fetch()
        if not (np.array_equal(doff, ooff) and np.array_equal(dids, oids)):
            return n, f"step {step}: manifold sizes / ids"
        for name in ("f1", "f2", "world1", "world2", "normal", "depth"):
            if not np.array_equal(dc[name].view(np.uint32), oc[name].view(np.uint32)):
                return n, f"step {step}: {name}"
        has = odir[:, 3] != 0
        if not (np.array_equal(ddir[:, 3] != 0, has) and np.array_equal(ddir[has].view(np.uint32), odir[has].view(np.uint32))):
            return n, f"step {step}: last_gjk_dir"
    return n, None

F = np.float32


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
    prox, gjk, orc = _build_shim("libprox_host.so", "proximity_host.cpp"), _build_shim("libgjk_host.so", "gjk_host.cpp"), Oracle()
    nar = _build_shim("libnarrow_host.so", "narrow_host.cpp")
    n_narrow = n_contacts = narrow_inexact = n_persist = n_capsule_pairs = capsule_inexact = 0
    t0 = time.time()
    n_scenes = n_prox = n_gjk = n_epa = 0
    inexact = 0
    bad = []
    seed = seed0
    while time.time() - t0 < budget and len(bad) < 10:
        rng = np.random.default_rng(seed)
        s = random_scene(rng, seed)
        pairs = orc.broad_phase(orc.compute_aabbs(s), s.groups)
        if len(pairs):
            both = np.concatenate([pairs, pairs[:, ::-1]])
            for margins in (None, rng.uniform(0, 1.0, size=len(both)).astype(F)):
                got, want = shim_proximity(prox, s, both, margins), orc.proximity(s, both, margins)
                n_prox += len(both)
                if not np.array_equal(got, want):
                    bad.append((seed, "proximity", int((got != want).sum())))
            t = s.shape_type
            cv = both[(t[both[:, 0]] != 0) & (t[both[:, 1]] != 0) & (t[both[:, 0]] != 3) & (t[both[:, 1]] != 3)]
            if len(cv):
                got, flags = shim_contact_sm_sm(gjk, s, cv)
                want, stats = orc.contact_sm_sm(s, cv)
                n_gjk += len(cv)
                n_epa += int(stats[2])
                if flags[0] or flags[1] or not np.array_equal(got[:, 9], want[:, 9]) or not np.allclose(got, want, rtol=1e-4, atol=1e-5):
                    bad.append((seed, "gjk/epa", int(flags[0]), int(flags[1])))
                inexact += int((got.view(np.uint32) != want.view(np.uint32)).any(axis=1).sum())
            # the whole narrow phase in the reference's callback orientation
            cb = orc.broad_phase(orc.compute_aabbs(s), s.groups, mode=0)
            got, want = shim_narrow_phase(nar, s, cb), orc.narrow_phase(s, cb)
            n_narrow += len(cb)
            n_contacts += len(want[0])
            try:
                narrow_inexact += compare_narrow(got, want, f"seed {seed}")
            except AssertionError as ex:
                bad.append((seed, "narrow phase", str(ex)[:120]))
            if seed % 3 == 0:  # a third of the balls / cuboids / hulls become capsules: the staged device functions of csrc/capsule.cuh
                sc = copy.copy(s)
                sc.shape_type, sc.shape_param = s.shape_type.copy(), s.shape_param.copy()
                pick = (sc.shape_type != 3) & (rng.random(sc.n) < 0.35)
                sc.shape_type[pick] = 4
                sc.shape_param[pick, 0] = rng.uniform(0.05, 0.6, size=int(pick.sum())).astype(F)
                sc.shape_param[pick, 1] = rng.uniform(0.02, 0.35, size=int(pick.sum())).astype(F)
                sc.shape_param[pick, 2:] = 0
                cbc = orc.broad_phase(orc.compute_aabbs(sc), sc.groups, mode=0)
                if len(cbc):
                    cbc = np.concatenate([cbc, cbc[:, ::-1]])
                    got, want = shim_narrow_phase_capsules(nar, sc, cbc), orc.narrow_phase(sc, cbc)
                    n_capsule_pairs += int((want[2] >= 7).sum())
                    try:
                        capsule_inexact += compare_narrow(got, want, f"seed {seed} capsules")
                    except AssertionError as ex:
                        bad.append((seed, "capsules", str(ex)[:120]))
            if seed % 4 == 0:  # stepping-world state per pair over a few updates (moves the scene: last check of this world)
                k, why = persist_rounds(nar, orc, s, cb, rng)
                n_persist += k
                if why:
                    bad.append((seed, "persistent manifold", why))
        n_scenes += 1
        seed += 1
    print(json.dumps({"scenes": n_scenes, "proximity_pairs": n_prox, "gjk_pairs": n_gjk, "epa_runs": n_epa, "gjk_rows_not_bit_exact": inexact,
                      "narrow_phase_pairs": n_narrow, "contacts": n_contacts, "contact_fields_not_bit_exact": narrow_inexact,
                      "persistent_edge_updates": n_persist,
                      "capsule_pairs": n_capsule_pairs, "capsule_contact_fields_not_bit_exact": capsule_inexact,
                      "mismatches": bad, "seconds": round(time.time() - t0, 1), "seed0": seed0}))


if __name__ == "__main__":
    main()
