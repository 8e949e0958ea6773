"""Work distribution of the device GJK (support evaluations per convex pair, exit kinds) on the cfg3 scene, from the device SOURCE compiled
for the host (tests/host_shim/gjk_host.cpp: shim_gjk_work_stats).  Design data for k_cc_gjk.  python scripts/gjk_work_stats.py [N]"""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from ncollide_b200 import _ffi  # noqa: E402
from ncollide_b200.scenes import config_scene  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402
from test_device_source_on_host import _build_shim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
s = config_scene(3, n)
orc = Oracle()
pairs = orc.broad_phase(orc.compute_aabbs(s), s.groups, mode=1)
t = s.shape_type
cv = np.ascontiguousarray(pairs[(t[pairs[:, 0]] != 0) & (t[pairs[:, 1]] != 0)])
lib = _build_shim("libgjk_host.so", "gjk_host.cpp")
oc, keep = _ffi.pack_objects(s)
hc, keep2 = _ffi.pack_hull_library(s.hulls)
st = np.zeros((len(cv), 3), dtype=np.uint32)
lib.shim_gjk_work_stats(C.byref(oc), C.byref(hc), C.c_uint64(len(cv)), _ffi.ptr(cv), _ffi.ptr(st))
ev, kind = st[:, 0], st[:, 1]
names = {0: "intersection (-> EPA)", 1: "closest points (-> manifold)", 3: "no intersection (done)"}
out = {"scene": s.name, "convex_pairs": int(len(cv)), "support_evals": {"mean": float(ev.mean()), "hist": {str(k): float((ev == k).mean()) for k in range(1, 13)}},
       "exits": {names[k]: {"share": float((kind == k).mean()), "mean_evals": float(ev[kind == k].mean()), "hist": {str(j): float((ev[kind == k] == j).mean()) for j in range(1, 9)}}
                 for k in (0, 1, 3)},
       "share_of_evals_spent_by_exit": {names[k]: float(ev[kind == k].sum() / ev.sum()) for k in (0, 1, 3)}}
print(json.dumps(out, indent=1))
