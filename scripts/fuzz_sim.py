"""Differential fuzzing of the stepping world (persistent broad phase, warm start, manifold cache, contact ids / events, add /
remove) and of the world queries against the oracle, for a wall-clock budget.  python scripts/fuzz_sim.py [seconds] [seed0]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from ncollide_b200.scenes import make_world_scene, with_sensors  # noqa: E402
from ncollide_b200.world import Context, SteppingWorld  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402
from sim_scenario import drive, drive_add_remove  # noqa: E402

RTOL, ATOL = 1e-4, 1e-5
F = np.float32


class Dev:
    def __init__(self, ctx, s):
        self.w = SteppingWorld(ctx, s)

    def set_positions(self, *a):
        self.w.set_positions(*a)

    def remove(self, h):
        self.w.remove(h)

    def add(self, sc):
        return self.w.add(sc)

    def step(self):
        r = self.w.update()
        keep = r["algo"] != 0
        cnt = np.diff(r["off"].astype(np.int64))
        sel = np.repeat(keep, cnt)
        off = np.concatenate([[0], np.cumsum(cnt[keep])]).astype(np.uint32)
        return {"pairs": r["pairs"][keep], "algo": r["algo"][keep], "off": off, "contacts": r["contacts"][sel], "ids": r["ids"][sel],
                "events": r["events"], "counts": r["counts"], "bp_pairs": len(r["pairs"]), "prox": r["prox"][keep], "prox_events": r["prox_events"]}


def ev_sorted(e):
    return e[np.lexsort((e[:, 1], e[:, 0], e[:, 2]))] if len(e) else e


def same(dev, orc):
    for t, (a, b) in enumerate(zip(dev, orc)):
        if a["bp_pairs"] != b["bp_pairs"]:
            return f"step {t}: broad-phase pairs"
        for k in ("pairs", "algo", "off", "ids"):
            if not np.array_equal(a[k], b[k]):
                return f"step {t}: {k}"
        if not np.array_equal(ev_sorted(a["events"]), ev_sorted(b["events"])):
            return f"step {t}: events"
        if not np.array_equal(a["prox"], b["prox"]):  # proximity sensors
            return f"step {t}: proximity statuses"
        pa, pb = np.asarray(a["prox_events"]).reshape(-1, 4), np.asarray(b["prox_events"]).reshape(-1, 4)
        if not np.array_equal(pa[np.lexsort(pa.T[::-1])] if len(pa) else pa, pb[np.lexsort(pb.T[::-1])] if len(pb) else pb):
            return f"step {t}: proximity events"
        for f in ("f1", "f2"):
            if not np.array_equal(a["contacts"][f], b["contacts"][f]):
                return f"step {t}: {f}"
        for f in ("world1", "world2", "normal", "depth"):
            if not np.allclose(a["contacts"][f], b["contacts"][f], rtol=RTOL, atol=ATOL):
                return f"step {t}: {f}"
        if a["counts"]["epa_overflow"] or a["counts"]["ref_panics"]:
            return f"step {t}: counters"
    return None


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
    ctx, orc = Context(0), Oracle()
    t0 = time.time()
    n_worlds = n_steps = n_rays = 0
    bad = []
    seed = seed0
    while time.time() - t0 < budget and len(bad) < 5:
        rng = np.random.default_rng(seed)
        n = int(rng.integers(20, 900))
        kinds = [(1, 1, 1), (0, 1, 1), (1, 0, 1), (1, 1, 0)][rng.integers(0, 4)]
        side = max(1.0, (n ** (1 / 3)) * 1.1 / rng.choice([0.8, 1.5, 3.0]))
        hull_kinds = kinds[2] > 0
        s = make_world_scene(n, seed, kinds, side=side, n_hulls=int(rng.integers(2, 16)), plane=bool(rng.random() < 0.3),
                             linear=float(rng.choice([0.0, 0.02, 0.1])), angular=float(rng.choice([0.0, 0.02, 0.2])),
                             margin=float(rng.choice([0.0, 0.02, 0.08])))
        sensors = rng.random() < 0.4
        if sensors:
            with_sensors(s, float(rng.choice([0.1, 0.5, 1.0])), seed + 7, margin=float(rng.choice([0.0, 0.05, 0.3])))
            if not s.query_kind.any():
                s.query_kind = None
        if rng.random() < 0.5:
            extra_kinds = kinds if hull_kinds else (kinds[0], kinds[1], 0)
            extra = make_world_scene(max(6, n // 5), seed + 1, extra_kinds, side=side, hull_library=s.hulls,
                                     linear=float(s.query_limit[0]), angular=float(rng.choice([0.0, 0.05])), margin=s.margin)
            if sensors or rng.random() < 0.2:  # sensors among the added objects (also into a world that had none)
                with_sensors(extra, 0.5, seed + 9, margin=0.1)
            if s.n // 10 >= 1:
                a = drive_add_remove(Dev(ctx, s), s, extra, steps=7, seed=seed)
                b = drive_add_remove(orc.sim(s), s, extra, steps=7, seed=seed)
            else:
                a = b = []
        else:
            d, o = Dev(ctx, s), orc.sim(s)
            a, b = drive(d, s, steps=6, seed=seed), drive(o, s, steps=6, seed=seed)
            ro = rng.uniform(-1, side + 1, size=(64, 3)).astype(F)
            rd = rng.normal(size=(64, 3)).astype(F)
            for first in (False, True):
                x, y = d.w.ray_cast(ro, rd, 3 * side, first_only=first), o.ray_cast(ro, rd, 3 * side, first_only=first)
                if not (np.array_equal(x[0], y[0]) and np.array_equal(x[3], y[3]) and np.allclose(x[1], y[1], rtol=RTOL, atol=ATOL)
                        and np.allclose(x[2], y[2], rtol=RTOL, atol=ATOL)):
                    bad.append((seed, f"ray query first={first}"))
            pts = rng.uniform(0, side, size=(64, 3)).astype(F)
            if not np.array_equal(d.w.query(2, pts), o.query(2, pts)):
                bad.append((seed, "point query"))
            n_rays += 128
        why = same(a, b)
        if why:
            bad.append((seed, why))
        n_worlds += 1
        n_steps += len(a)
        seed += 2
    print(json.dumps({"worlds": n_worlds, "steps": n_steps, "query_rays": n_rays, "mismatches": bad, "seconds": round(time.time() - t0, 1), "seed0": seed0}))


if __name__ == "__main__":
    main()
