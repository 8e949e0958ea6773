"""Wall time of ncb2d_contact over repeated calls (1 M pairs, pageable host buffers), to separate first-call costs from the steady state."""
import time

import numpy as np

from ncollide_b200 import dim2
from ncollide_b200.world import Context

ctx = Context(0)
rng = np.random.default_rng(21)
n = 1_000_000
sh = dim2.Shapes2D()
for t in (0, 1, 2):
    for _ in range(64):
        if t == 0:
            sh.ball(rng.uniform(0.2, 0.6))
        elif t == 1:
            sh.cuboid(rng.uniform(0.2, 0.6), rng.uniform(0.2, 0.6))
        else:
            k = int(rng.integers(3, 13))
            ang = np.sort(rng.uniform(0, 2 * np.pi, size=k)) + np.arange(k) * 1e-3
            sh.polygon(np.stack([0.5 * np.cos(ang), 0.35 * np.sin(ang)], axis=1))
typ, par, pts, nrm = sh.arrays()
t1, t2 = rng.choice([0, 1, 2], size=n), rng.choice([0, 1, 2], size=n)
p1, p2 = t1 * 64 + rng.integers(0, 64, size=n), t2 * 64 + rng.integers(0, 64, size=n)
c1 = rng.uniform(-50, 50, size=(n, 2))
m1 = dim2.isometry2(c1, rng.uniform(-np.pi, np.pi, size=n))
m2 = dim2.isometry2(c1 + rng.uniform(-1.2, 1.2, size=(n, 2)), rng.uniform(-np.pi, np.pi, size=n))
args = (typ[p1], par[p1], m1, typ[p2], par[p2], m2, pts)
for k in range(6):
    t0 = time.perf_counter()
    found, out, info = dim2.contact(ctx, *args, prediction=0.02, poly_normals=nrm)
    print(f"call {k}: {(time.perf_counter() - t0) * 1e3:.1f} ms, {int(found.sum())} contacts", flush=True)
margins = np.full(n, 0.05, dtype=np.float32)
for k in range(3):
    t0 = time.perf_counter()
    st = dim2.proximity(ctx, *args, margins=margins)
    print(f"proximity call {k}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
ctx.close()
