"""World ray queries: first_interference_with_ray for R rays against an N-object world through the C ABI
(host buffers in, rows out), and the oracle on a small sample.  python scripts/bench_world_rays.py [N] [R] [R_cpu]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from ncollide_b200.scenes import config_scene  # noqa: E402
from ncollide_b200.world import Context, SteppingWorld  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
R_cpu = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
s = config_scene(3, n)
side = float(s.pos.max())
rng = np.random.default_rng(3)
o = rng.uniform(0, side, size=(R, 3)).astype(np.float32)
d = rng.normal(size=(R, 3)).astype(np.float32)
w = SteppingWorld(Context(0), s)
w.update(fetch=False)
res = {}
for name, max_toi, first in (("first_hit_within_20", 20.0, True), ("all_hits_within_20", 20.0, False), ("first_hit_unbounded", float(np.finfo(np.float32).max), True)):
    times = []
    for it in range(4):
        t0 = time.perf_counter()
        idx, toi, normal, feat = w.ray_cast(o, d, max_toi, first_only=first)
        times.append(time.perf_counter() - t0)
    res[name] = {"ms": 1e3 * min(times[1:]), "rows": int(len(idx)), "Mrays_per_s": R / min(times[1:]) / 1e6}
from oracle.pyoracle import Oracle  # noqa: E402

orc = Oracle().sim(s)
orc.step()
t0 = time.perf_counter()
idx, toi, normal, feat = orc.ray_cast(o[:R_cpu], d[:R_cpu], 20.0, first_only=True)
dt = time.perf_counter() - t0
res["oracle_cpu_first_hit_within_20"] = {"rays": R_cpu, "ms": 1e3 * dt, "Mrays_per_s": R_cpu / dt / 1e6}
a = w.ray_cast(o[:R_cpu], d[:R_cpu], 20.0, first_only=True)
res["sample_equal"] = bool(np.array_equal(a[0], idx) and np.allclose(a[1], toi, rtol=1e-4, atol=1e-5))
print(json.dumps({"n_objects": n, "n_rays": R, **res}))
