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


def drive_add_remove(sim, scene, extra, steps, seed):
    """Like `drive`, with objects removed and added between updates: step 2 removes a tenth of the world, step 3 adds the
    objects of `extra` (recycled handles first), step 5 removes some of the new ones and moves others."""
    rng = np.random.default_rng(seed)
    n0 = scene.n
    pos = np.concatenate([scene.pos, extra.pos]).copy()  # indexed by handle once handles are known
    rot = np.concatenate([scene.rot, extra.rot]).copy()
    alive = np.zeros(n0 + extra.n, dtype=bool)
    alive[:n0] = True
    pos_h = {h: scene.pos[h].copy() for h in range(n0)}
    rot_h = {h: scene.rot[h].copy() for h in range(n0)}
    log, new_handles = [], None
    for t in range(steps):
        if t == 2:
            gone = np.sort(rng.choice(n0, size=n0 // 10, replace=False)).astype(np.uint32)
            sim.remove(gone)
            for h in gone.tolist():
                del pos_h[h], rot_h[h]
        if t == 3:
            new_handles = np.asarray(sim.add(extra)).astype(np.uint32)
            for k, h in enumerate(new_handles.tolist()):
                pos_h[h], rot_h[h] = extra.pos[k].copy(), extra.rot[k].copy()
        if t == 5:
            sim.remove(new_handles[::3])
            for h in new_handles[::3].tolist():
                del pos_h[h], rot_h[h]
        if t in (1, 4, 5, 6):
            hs = np.array(sorted(pos_h), dtype=np.uint32)
            idx = hs[rng.random(len(hs)) < 0.3]
            p = np.array([pos_h[h] for h in idx.tolist()], dtype=F32) + rng.normal(0, 0.01, size=(len(idx), 3)).astype(F32)
            r = np.array([rot_h[h] for h in idx.tolist()], dtype=F32)
            for h, pp in zip(idx.tolist(), p):
                pos_h[h] = pp
            sim.set_positions(idx, p, r)
        r = sim.step()
        r["new_handles"] = new_handles
        log.append(r)
    return log


def drive_group_changes(sim, scene, steps, seed):
    """Like `drive`, with CollisionObject::set_collision_groups on live objects between updates: objects leave the default groups
    (pairs stop, with ContactEvent::Stopped when they were touching), come back (pairs start again), and some do both while moving."""
    rng = np.random.default_rng(seed)
    pos, rot = scene.pos.copy(), scene.rot.copy()
    n = scene.n
    ALL = 0x3FFFFFFF
    groups = np.tile(np.array([ALL, ALL, 0], dtype=np.uint32), (n, 1))
    log = []
    for t in range(steps):
        if t in (1, 4):  # a tenth of the world stops talking to group 1 members; a few become group-1 members only
            a = np.sort(rng.choice(n, size=n // 10, replace=False)).astype(np.uint32)
            groups[a] = np.array([ALL, ALL & ~2, 0], dtype=np.uint32)
            b = np.sort(rng.choice(n, size=n // 10, replace=False)).astype(np.uint32)
            groups[b] = np.array([2, ALL, 0], dtype=np.uint32)
            hs = np.unique(np.concatenate([a, b]))
            sim.set_collision_groups(hs, groups[hs])
        if t in (2, 5):  # half of the changed objects return to the default groups
            changed = np.nonzero((groups[:, 0] != ALL) | (groups[:, 1] != ALL))[0].astype(np.uint32)
            back = changed[rng.random(len(changed)) < 0.5]
            groups[back] = np.array([ALL, ALL, 0], dtype=np.uint32)
            if len(back):
                sim.set_collision_groups(back, groups[back])
        if t in (1, 3, 5, 6):
            idx = step_poses(scene, pos, rot, rng, 0.3)
            sim.set_positions(idx, pos[idx], rot[idx])
        r = sim.step()
        r["groups"] = groups.copy()
        log.append(r)
    return log
