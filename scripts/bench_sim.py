"""Stepping-world benchmark: N objects, a fraction moves every step; times SteppingWorld.update on the
device (CUDA-event-free wall clock around the C-ABI call incl. the pose upload of the moved objects, results left on the
device) against the oracle's reference-faithful stepping world on a smaller N.
Usage: python scripts/bench_sim.py [N] [steps] [move_fraction] [N_cpu]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from ncollide_b200.scenes import config_scene  # noqa: E402
from ncollide_b200.world import Context, SteppingWorld  # noqa: E402
from sim_scenario import step_poses  # noqa: E402


def run(make_sim, scene, steps, frac, fetch):
    rng = np.random.default_rng(11)
    pos, rot = scene.pos.copy(), scene.rot.copy()
    sim = make_sim(scene)
    t0 = time.perf_counter()
    r = sim.step() if not fetch else sim.update(fetch=False)
    first = time.perf_counter() - t0
    times, info = [], None
    for _ in range(steps):
        idx = step_poses(scene, pos, rot, rng, frac)
        t0 = time.perf_counter()
        sim.set_positions(idx, pos[idx], rot[idx])
        info = sim.step() if not fetch else sim.update(fetch=False)
        times.append(time.perf_counter() - t0)
    return first, times, info


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
    n_cpu = int(sys.argv[4]) if len(sys.argv) > 4 else 100_000
    s = config_scene(3, n)
    ctx = Context(0)
    first, times, info = run(lambda sc: SteppingWorld(ctx, sc), s, steps, frac, True)
    dev = {"n": n, "first_update_ms": 1e3 * first, "step_ms_median": 1e3 * float(np.median(times)), "step_ms_all": [round(1e3 * t, 2) for t in times],
           "counts": info["counts"]}
    from oracle.pyoracle import Oracle

    sc = config_scene(3, n_cpu)
    o = Oracle()
    first, times, info = run(lambda x: o.sim(x), sc, min(steps, 4), frac, False)
    cpu = {"n": n_cpu, "first_update_ms": 1e3 * first, "step_ms_median": 1e3 * float(np.median(times)), "pairs": int(len(info["pairs"])),
           "contacts": int(len(info["contacts"]))}
    print(json.dumps({"device": dev, "oracle_cpu": cpu, "move_fraction": frac}))
