import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from ncollide_b200 import dim2
from ncollide_b200.world import Context
ctx = Context(0)
rng = np.random.default_rng(21)
sh = dim2.Shapes2D()
for t in (0, 1, 2):
    for _ in range(64):
        if t == 0: sh.ball(rng.uniform(0.2, 0.6))
        elif t == 1: sh.cuboid(rng.uniform(0.2, 0.6), rng.uniform(0.2, 0.6))
        else:
            k = int(rng.integers(3, 13)); ang = np.sort(rng.uniform(0, 2 * np.pi, size=k)) + np.arange(k) * 1e-3
            sh.polygon(np.stack([0.5 * np.cos(ang), 0.35 * np.sin(ang)], axis=1))
n = 1_000_000
side = float(np.sqrt(n * 0.8 / 2.5))
w = dim2.World2D.from_library(sh, rng.integers(0, 192, size=n), rng.uniform(0, side, size=(n, 2)), rng.uniform(-np.pi, np.pi, size=n))
for _ in range(3):
    t0 = time.perf_counter(); r = dim2.world_update(ctx, w); print("total ms", (time.perf_counter() - t0) * 1e3, len(r["pairs"]), len(r["contacts"]), file=sys.stderr)
