"""Work distribution of the device EPA (steps, polytope size per penetrating convex pair) on the cfg3 scene, computed on the CPU from
the device SOURCE compiled for the host (tests/host_shim/gjk_host.cpp: shim_epa_work_stats).  Design data for restructuring k_cc_epa
(its capacity choices, a small-polytope fast path, a two-pass split); no GPU involved.  python scripts/epa_work_stats.py [N]"""
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
st = np.zeros((len(cv), 8), dtype=np.uint32)
lib.shim_epa_work_stats(C.byref(oc), C.byref(hc), C.c_uint64(len(cv)), _ffi.ptr(cv), _ffi.ptr(st))
e = st[st[:, 0] > 0]


def pct(a, qs=(50, 75, 90, 95, 99, 99.9, 100)):
    return {str(q): float(np.percentile(a, q)) for q in qs}


out = {
    "scene": s.name, "convex_pairs": int(len(cv)), "epa_pairs": int(len(e)), "epa_fraction": float(len(e) / max(len(cv), 1)),
    "steps": {"mean": float(e[:, 0].mean()), "percentiles": pct(e[:, 0]), "share_of_all_steps_in_the_longest_10pct_of_pairs":
              float(np.sort(e[:, 0])[int(0.9 * len(e)):].sum() / e[:, 0].sum())},
    "vertices_at_exit": pct(e[:, 1]), "faces_at_exit": pct(e[:, 2]), "heap_entries_at_exit": pct(e[:, 3]),
    "peak_heap": pct(e[:, 4]), "peak_silhouette": pct(e[:, 5]), "peak_flood_stack": pct(e[:, 6]),
    "simplex_dim_plus_1": {str(k): float((e[:, 7] == k).mean()) for k in (1, 2, 3, 4)},
    "capacities": {"EPA_MAX_VERTS": 48, "EPA_MAX_FACES": 192, "EPA_MAX_HEAP": 160},
    "compact_store": {"capacities": {"verts": 16, "faces": 48, "heap": 24, "silhouette": 16, "flood_stack": 12},
                      "fits": float(((e[:, 1] <= 16) & (e[:, 2] <= 48) & (e[:, 4] <= 24) & (e[:, 5] <= 16) & (e[:, 6] <= 12) & (e[:, 7] == 4)).mean())},
    "fits": {f"<= {v} verts and <= {f} faces": float(((e[:, 1] <= v) & (e[:, 2] <= f)).mean()) for v, f in ((8, 16), (12, 32), (16, 48), (24, 80), (32, 128))},
}
print(json.dumps(out))
