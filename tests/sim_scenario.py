"""Shared driver for the stepping-world tests: deterministic pose updates of a random subset per step."""
import numpy as np

F32 = np.float32


def step_poses(scene, pos, rot, rng, frac=0.4):
    """Moves a random `frac` of the objects: small jitters (resting contacts), a few jumps; small random rotations."""
    n = scene.n
    idx = np.sort(rng.choice(n, size=max(1, int(n * frac)), replace=False)).astype(np.uint32)
    d = rng.normal(0, 0.004, size=(len(idx), 3)).astype(F32)
    jump = rng.random(len(idx)) < 0.1
    d[jump] = rng.normal(0, 0.3, size=(int(jump.sum()), 3)).astype(F32)
    pos[idx] = (pos[idx] + d).astype(F32)
    dq = rng.normal(0, 0.01, size=(len(idx), 4)).astype(F32)
    q = (rot[idx] + dq).astype(F32)
    nrm = np.sqrt((q * q).sum(axis=1, dtype=F32), dtype=F32)
    q = (q / nrm[:, None]).astype(F32)
    nrm = np.sqrt((q * q).sum(axis=1, dtype=F32), dtype=F32)
    rot[idx] = (q / nrm[:, None]).astype(F32)
    return idx


def drive(sim, scene, steps, seed, frac=0.4):
    """sim: object with set_positions(handles, pos, rot) and step() -> dict.  Returns the per-step dicts (+ 'moved')."""
    rng = np.random.default_rng(seed)
    pos, rot = scene.pos.copy(), scene.rot.copy()
    log = []
    for t in range(steps):
        moved = np.ones(scene.n, dtype=bool) if t == 0 else np.zeros(scene.n, dtype=bool)
        if t > 0 and t != 3:  # step 3 moves nothing at all
            idx = step_poses(scene, pos, rot, rng, frac)
            sim.set_positions(idx, pos[idx], rot[idx])
            moved[idx] = True
        r = sim.step()
        r["moved"] = moved
        log.append(r)
    return log
