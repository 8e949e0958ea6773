"""Proximity-sensor figures on their own: bench.bench_proximity on the cfg3 scene, plus the oracle's detectors
timed on a bounded sample of the same sensor pairs (one host core).  Usage: python scripts/bench_proximity.py [N] [fraction]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from ncollide_b200.scenes import config_scene, with_sensors  # noqa: E402
from ncollide_b200.world import Context  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
    scene = config_scene(3, n)
    ctx = Context(0)
    ctx.set_hulls(scene.hulls)
    out = bench.bench_proximity(ctx, scene, frac)
    # CPU beside it: the oracle's detectors on a sample of the sensor pairs of the same world
    import copy

    from oracle.pyoracle import Oracle

    s = with_sensors(copy.copy(scene), frac, 5)
    ctx.set_scene(s)
    res = ctx.world_fetch(ctx.world_update_device(s.margin))
    pp = res.pairs[res.pair_algo == 6][:200_000]
    orc = Oracle()
    t0 = time.perf_counter()
    want = orc.proximity(s, pp)
    dt = time.perf_counter() - t0
    assert np.array_equal(want, res.proximity[res.pair_algo == 6][: len(pp)])
    out["cpu_oracle"] = {"pairs": int(len(pp)), "ms": dt * 1e3, "Mpairs_per_s": len(pp) / dt / 1e6, "cores": 1, "parity": "statuses equal on the sample"}
    print(json.dumps(out))
