"""Persistent broad phase stepping benchmark: N proxies, a fraction moves every step.
Times BroadPhase.update through the C ABI with host buffers (set_bounding_volumes + update + events),
and the oracle's reference-faithful DBVTBroadPhase on the same workload (smaller N by default).
Usage: python scripts/bench_bp.py [N] [steps] [move_fraction] [N_cpu]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from ncollide_b200.world import BroadPhase, Context  # noqa: E402

F32 = np.float32


def workload(n, steps, frac, seed=7):
    rng = np.random.default_rng(seed)
    side = (n * (4 * 0.45) ** 3 / 4.0) ** (1 / 3)  # ~4 AABB neighbours per proxy, like the world scenes
    c = rng.uniform(0, side, size=(n, 3)).astype(F32)
    e = rng.uniform(0.25, 0.5, size=(n, 3)).astype(F32)
    moves = []
    for _ in range(steps):
        k = int(n * frac)
        idx = rng.choice(n, size=k, replace=False).astype(np.uint32)
        d = rng.normal(0, 0.05, size=(k, 3)).astype(F32)
        moves.append((idx, d))
    return c, e, moves


def run_device(n, steps, frac):
    c, e, moves = workload(n, steps, frac)
    bp = BroadPhase(0.02, ctx=Context(0))
    t0 = time.perf_counter()
    hs = bp.create_proxies(np.concatenate([c - e, c + e], axis=1))
    st, sp = bp.update_events()
    t_first = time.perf_counter() - t0
    times, ev = [], []
    for idx, d in moves:
        c[idx] += d
        boxes = np.concatenate([c[idx] - e[idx], c[idx] + e[idx]], axis=1)
        t0 = time.perf_counter()
        bp.deferred_set_bounding_volumes(hs[idx], boxes)
        st, sp = bp.update_events()
        times.append(time.perf_counter() - t0)
        ev.append((len(st), len(sp)))
    return {"n": n, "first_update_s": t_first, "step_ms_median": 1e3 * float(np.median(times)), "events_last": ev[-1],
            "pairs": bp.num_interferences()}


def run_oracle(n, steps, frac):
    from oracle.pyoracle import Oracle

    c, e, moves = workload(n, steps, frac)
    o = Oracle()
    lib, bp = o.lib, o.broad_phase_persistent(0.02)
    boxes = np.concatenate([c - e, c + e], axis=1)
    t0 = time.perf_counter()
    for b in boxes:
        bp.create_proxy(b)
    bp.update()
    t_first = time.perf_counter() - t0
    times = []
    for idx, d in moves:
        c[idx] += d
        boxes = np.concatenate([c[idx] - e[idx], c[idx] + e[idx]], axis=1)
        t0 = time.perf_counter()
        for h, b in zip(idx.tolist(), boxes):
            bp.deferred_set_bounding_volume(h, b)
        t_set = time.perf_counter() - t0  # mostly ctypes call overhead: reported separately
        t0 = time.perf_counter()
        bp.update()
        times.append((t_set, time.perf_counter() - t0))
    return {"n": n, "first_update_s": t_first, "step_update_ms_median": 1e3 * float(np.median([t[1] for t in times])),
            "step_set_ms_median_ctypes": 1e3 * float(np.median([t[0] for t in times])), "pairs": bp.num_interferences()}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
    n_cpu = int(sys.argv[4]) if len(sys.argv) > 4 else 200_000
    print(json.dumps({"device": run_device(n, steps, frac), "oracle_cpu": run_oracle(n_cpu, min(steps, 4), frac), "move_fraction": frac}))
