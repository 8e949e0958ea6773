"""Differential fuzzing WITHOUT a GPU: the per-pair device source (proximity detectors of csrc/proximity.cu, GJK / EPA of
csrc/gjk.cuh, the contact generators / clipping / manifold of csrc/narrow.cu), compiled for the host through tests/host_shim/, against the oracle on many small random worlds (sizes, densities,
margins, degenerate placements, scales, far-away coordinates: the generator of scripts/fuzz_parity.py).
python scripts/fuzz_host_shim.py [seconds] [seed0]  ->  one JSON summary line.  Test infrastructure only."""
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
from test_device_source_on_host import _build_shim, compare_narrow, shim_contact_sm_sm, shim_narrow_phase, shim_proximity  # noqa: E402

F = np.float32


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
    prox, gjk, orc = _build_shim("libprox_host.so", "proximity_host.cpp"), _build_shim("libgjk_host.so", "gjk_host.cpp"), Oracle()
    nar = _build_shim("libnarrow_host.so", "narrow_host.cpp")
    n_narrow = n_contacts = narrow_inexact = 0
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
        n_scenes += 1
        seed += 1
    print(json.dumps({"scenes": n_scenes, "proximity_pairs": n_prox, "gjk_pairs": n_gjk, "epa_runs": n_epa, "gjk_rows_not_bit_exact": inexact,
                      "narrow_phase_pairs": n_narrow, "contacts": n_contacts, "contact_fields_not_bit_exact": narrow_inexact,
                      "mismatches": bad, "seconds": round(time.time() - t0, 1), "seed0": seed0}))


if __name__ == "__main__":
    main()
