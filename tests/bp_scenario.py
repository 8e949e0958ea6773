"""Shared multi-step scenario for the persistent broad phase (BroadPhase trait over several updates).

`run_scenario(impl, seed, ...)` drives an implementation of the trait surface through random create / move / remove
steps and records, per step: the started events (oriented), the stopped events, the removal events,
num_interferences, and a sample of `proxy()` boxes.  Implementations: the reference-faithful oracle
(`oracle.pyoracle.OracleBroadPhase`), the device broad phase (`ncollide_b200.world.BroadPhase`), and
`SetModel` below — a numpy statement of the set semantics the device path relies on (stored box rule + all-pairs of
stored boxes + orientation rule), which the CPU suite checks against the oracle.
"""
import numpy as np

F32 = np.float32
ALL = 0x3FFFFFFF


class OracleAdapter:
    def __init__(self, oracle, margin):
        self.bp = oracle.broad_phase_persistent(margin)
        self.groups = np.zeros((0, 3), dtype=np.uint32)

    def _grow(self, top):
        if top > len(self.groups):
            g = np.zeros((max(top, 2 * len(self.groups)), 3), dtype=np.uint32)
            g[: len(self.groups)] = self.groups
            self.groups = g

    def create(self, bvs, groups):
        out = []
        for bv, g in zip(bvs, groups):
            h = self.bp.create_proxy(bv)
            self._grow(h + 1)
            self.groups[h] = g
            out.append(h)
        return out

    def set_bvs(self, handles, bvs):
        for h, bv in zip(handles, bvs):
            self.bp.deferred_set_bounding_volume(int(h), bv)

    def remove(self, handles):
        return self.bp.remove(handles)

    def update(self):
        return self.bp.update(self.groups if len(self.groups) else None)

    def set_groups(self, h, g, requeue=True):
        self.groups[h] = g
        if requeue:
            self.bp.deferred_recompute_all_proximities_with(int(h))

    def recompute_all(self):
        self.bp.deferred_recompute_all_proximities()

    def query(self, kind, q):
        return QUERY_IMPL[type(self).__name__](self, kind, q)

    def num(self):
        return self.bp.num_interferences()

    def proxy(self, h):
        return self.bp.proxy(h)

    def pairs(self):
        return self.bp.pairs()


class DeviceAdapter:
    def __init__(self, ctx, margin):
        from ncollide_b200.world import BroadPhase

        self.bp = BroadPhase(margin, ctx=ctx)

    def create(self, bvs, groups):
        return self.bp.create_proxies(np.asarray(bvs, dtype=F32), None, np.asarray(groups, dtype=np.uint32)).tolist()

    def set_bvs(self, handles, bvs):
        if len(handles):
            self.bp.deferred_set_bounding_volumes(handles, bvs)

    def remove(self, handles):
        return self.bp.remove(handles)

    def update(self):
        return self.bp.update_events()

    def set_groups(self, h, g, requeue=True):
        if requeue:
            self.bp.set_groups(h, g)
        else:
            self.bp._groups[int(h)] = g
            self.bp._any_groups = True

    def recompute_all(self):
        self.bp.deferred_recompute_all_proximities()

    def query(self, kind, q):
        return QUERY_IMPL[type(self).__name__](self, kind, q)

    def num(self):
        return self.bp.num_interferences()

    def proxy(self, h):
        r = self.bp.proxy(h)
        return None if r is None else r[0]

    def pairs(self):
        return self.bp.pairs()


class SetModel:
    """The set semantics of DBVTBroadPhase restated without trees (what csrc/bp_persistent.cu computes)."""

    def __init__(self, margin):
        self.margin = F32(margin)
        self.occupied, self.attached = [], []
        self.free = []  # LIFO
        self.box = np.zeros((0, 6), dtype=F32)
        self.groups = np.zeros((0, 3), dtype=np.uint32)
        self.pending = {}  # handle -> [first seq, box]; survives removal like the reference's queue
        self.seq = 0
        self.front = 0
        self.set = set()

    def _slot(self):
        if self.free:
            return self.free.pop()
        self.occupied.append(False)
        self.attached.append(False)
        self.box = np.vstack([self.box, np.zeros((1, 6), dtype=F32)])
        self.groups = np.vstack([self.groups, np.zeros((1, 3), dtype=np.uint32)])
        return len(self.occupied) - 1

    def _push(self, h, box):
        if h in self.pending:
            self.pending[h][1] = box
        else:
            self.pending[h] = [self.seq, box]
        self.seq += 1

    def _push_front(self, h):
        if not (self.occupied[h] and self.attached[h]):
            return
        self.front += 1
        if h in self.pending:
            self.pending[h][0] = min(self.pending[h][0], -self.front)
        else:
            self.pending[h] = [-self.front, self.box[h].copy()]

    def set_groups(self, h, g, requeue=True):
        self.groups[h] = g
        if requeue:
            self._push_front(int(h))

    def recompute_all(self):
        for h in range(len(self.occupied)):
            self._push_front(h)

    def query(self, kind, q):
        return _model_query(self, kind, q)

    def create(self, bvs, groups):
        out = []
        for bv, g in zip(bvs, groups):
            h = self._slot()
            self.occupied[h], self.attached[h] = True, False
            self.groups[h] = g
            self._push(h, np.asarray(bv, dtype=F32).copy())
            out.append(h)
        return out

    def set_bvs(self, handles, bvs):
        for h, bv in zip(handles, bvs):
            h = int(h)
            bv = np.asarray(bv, dtype=F32)
            if h >= len(self.occupied) or not self.occupied[h]:
                raise RuntimeError("Attempting to set the bounding volume of an object that does not exist.")
            if self.attached[h]:
                s = self.box[h]
                if np.all(s[:3] <= bv[:3]) and np.all(s[3:] >= bv[3:]):
                    continue
            loose = bv.copy()
            loose[:3] = bv[:3] + (-self.margin)
            loose[3:] = bv[3:] + self.margin
            self._push(h, loose)

    def remove(self, handles):
        hs = set(int(h) for h in handles)
        gone = sorted(p for p in self.set if p[0] in hs or p[1] in hs)
        self.set -= set(gone)
        for h in handles:
            h = int(h)
            self.occupied[h] = self.attached[h] = False
            self.free.append(h)
        return np.array(gone, dtype=np.uint32).reshape(-1, 2)

    def _allowed(self, a, b):
        g1, g2 = self.groups[a], self.groups[b]
        return (g1[0] & g2[2]) == 0 and (g2[0] & g1[2]) == 0 and (g1[0] & g2[1]) != 0 and (g2[0] & g1[1]) != 0

    def update(self):
        upd = {}
        for h, (seq, box) in self.pending.items():
            if self.occupied[h]:
                upd[h] = seq
                self.box[h] = box
                self.attached[h] = True
        self.pending = {}
        self.seq = self.front = 0
        if not upd:
            return np.zeros((0, 2), np.uint32), np.zeros((0, 2), np.uint32)
        alive = [h for h in range(len(self.occupied)) if self.occupied[h]]
        b = self.box[alive]
        new = set()
        if len(alive) > 1:
            ov = np.all(b[:, None, :3] <= b[None, :, 3:], axis=2) & np.all(b[:, None, 3:] >= b[None, :, :3], axis=2)
            ii, jj = np.nonzero(np.triu(ov, 1))
            for i, j in zip(ii.tolist(), jj.tolist()):
                a, c = alive[i], alive[j]
                if self._allowed(a, c):
                    new.add((a, c))
        started = []
        for lo, hi in sorted(new - self.set):
            sl, sh = upd.get(lo), upd.get(hi)
            hi_first = sh is not None and (sl is None or sh > sl)
            started.append((hi, lo) if hi_first else (lo, hi))
        stopped = sorted(self.set - new)
        self.set = new
        return np.array(started, dtype=np.uint32).reshape(-1, 2), np.array(stopped, dtype=np.uint32).reshape(-1, 2)

    def num(self):
        return len(self.set)

    def proxy(self, h):
        return self.box[h].copy() if h < len(self.occupied) and self.occupied[h] and self.attached[h] else None

    def pairs(self):
        return np.array(sorted(self.set), dtype=np.uint32).reshape(-1, 2)


def _oracle_query(self, kind, q):
    rows = []
    for i, r in enumerate(q):
        if kind == 0:
            hs = self.bp.interferences_with_bounding_volume(r)
        elif kind == 1:
            hs = self.bp.interferences_with_ray(r[:3], r[3:6], r[6])
        else:
            hs = self.bp.interferences_with_point(r)
        rows += [(i, int(h)) for h in hs]
    return np.array(rows, dtype=np.uint32).reshape(-1, 2)


def _device_query(self, kind, q):
    if kind == 0:
        return self.bp.interferences_with_bounding_volumes(q)
    if kind == 1:
        return self.bp.interferences_with_rays(q[:, :3], q[:, 3:6], q[:, 6])
    return self.bp.interferences_with_points(q)


def _model_query(self, kind, q):
    rows = []
    hs = [h for h in range(len(self.occupied)) if self.occupied[h] and self.attached[h]]
    for i, r in enumerate(q):
        for h in hs:
            b = self.box[h]
            if kind == 0:
                ok = np.all(b[:3] <= r[3:6]) and np.all(b[3:] >= r[:3])
            elif kind == 2:
                ok = not (np.any(r < b[:3]) or np.any(r > b[3:]))
            else:
                ok = _slab(b, r[:3], r[3:6], r[6])
            if ok:
                rows.append((i, h))
    return np.array(rows, dtype=np.uint32).reshape(-1, 2)


def _slab(b, o, d, max_toi):
    tmin, tmax = F32(0), F32(max_toi)
    for i in range(3):
        if d[i] == 0:
            if o[i] < b[i] or o[i] > b[3 + i]:
                return False
        else:
            inv = F32(1) / d[i]
            near, far = (b[i] - o[i]) * inv, (b[3 + i] - o[i]) * inv
            if near > far:
                near, far = far, near
            tmin, tmax = max(tmin, near), min(tmax, far)
            if tmin > tmax:
                return False
    return True


QUERY_IMPL = {"OracleAdapter": _oracle_query, "DeviceAdapter": _device_query, "SetModel": _model_query}


def sort_rows(a):
    a = np.asarray(a, dtype=np.uint32).reshape(-1, 2)
    return a[np.lexsort((a[:, 1], a[:, 0]))]


def run_scenario(impl, seed, n0=300, steps=12, side=10.0, use_groups=True, n_queries=0):
    """Returns a list of per-step records (dicts of numpy arrays / ints)."""
    rng = np.random.default_rng(seed)

    def rand_boxes(k):
        c = rng.uniform(0, side, size=(k, 3)).astype(F32)
        e = rng.uniform(0.2, 0.7, size=(k, 3)).astype(F32)
        return np.concatenate([c - e, c + e], axis=1).astype(F32)

    def rand_groups(k):
        g = np.zeros((k, 3), dtype=np.uint32)
        g[:, 0] = g[:, 1] = ALL
        if use_groups:
            for r in range(k):
                t = rng.integers(0, 6)
                if t == 0:  # member of group 1 only, blacklists group 1
                    g[r] = (1 << 1, ALL, 1 << 1)
                elif t == 1:  # member of group 2, whitelist group 2 + 3
                    g[r] = (1 << 2, (1 << 2) | (1 << 3), 0)
                elif t == 2:
                    g[r] = (1 << 3, ALL, 0)
        return g

    live = {}  # handle -> current tight box
    log = []
    b0 = rand_boxes(n0)
    for h, b in zip(impl.create(b0, rand_groups(n0)), b0):
        live[h] = b
    for step in range(steps):
        rec = {}
        removed_events = np.zeros((0, 2), np.uint32)
        if step > 0 and step % 5 != 4:  # every fifth step has no operation at all
            hs = np.array(sorted(live), dtype=np.uint32)
            # moves: small jitters (often contained in the loosened box), some jumps, some handles twice
            k = max(1, len(hs) // 3)
            mv = rng.choice(hs, size=k, replace=False)
            mv = np.concatenate([mv, mv[: k // 8]])
            newb = []
            for h in mv.tolist():
                b = live[h].copy()
                d = rng.normal(0, 0.03, 3).astype(F32) if rng.random() < 0.7 else rng.normal(0, 1.0, 3).astype(F32)
                b[:3] += d
                b[3:] += d
                live[h] = b
                newb.append(b)
            impl.set_bvs(mv, np.array(newb, dtype=F32))
            # removals (of attached proxies)
            rm = rng.choice(hs, size=min(len(hs) // 10, 25), replace=False)
            removed_events = impl.remove(rm)
            for h in rm.tolist():
                del live[h]
            # creations recycle the freed handles; one of them is moved again before the update, one is
            # removed again before it was ever attached and its slot reused
            nb = rand_boxes(int(rng.integers(5, 40)))
            hn = impl.create(nb, rand_groups(len(nb)))
            for h, b in zip(hn, nb):
                live[h] = b
            h_again = hn[0]
            b = live[h_again].copy()
            b[:3] += F32(0.25)
            b[3:] += F32(0.25)
            live[h_again] = b
            impl.set_bvs(np.array([h_again], dtype=np.uint32), b.reshape(1, 6))
            if len(hn) > 3:
                ev = impl.remove(np.array([hn[1]], dtype=np.uint32))
                assert len(ev) == 0
                del live[hn[1]]
                nb2 = rand_boxes(1)
                (h2,) = impl.create(nb2, rand_groups(1))
                assert h2 == hn[1], "slab handles are recycled LIFO"
                live[h2] = nb2[0]
        if use_groups and step % 4 == 2:
            # the groups of a few attached proxies change; the world re-queues them (glue/update.rs:83-86)
            hs = [h for h in sorted(live) if impl.proxy(h) is not None]
            for h in rng.choice(np.array(hs), size=min(12, len(hs)), replace=False).tolist():
                impl.set_groups(h, rand_groups(1)[0])
        if use_groups and step == steps - 2:
            # pair filter replaced: groups change silently, then deferred_recompute_all_proximities (world.rs:216)
            hs = [h for h in sorted(live) if impl.proxy(h) is not None]
            for h in rng.choice(np.array(hs), size=min(30, len(hs)), replace=False).tolist():
                impl.set_groups(h, rand_groups(1)[0], requeue=False)
            impl.recompute_all()
        started, stopped = impl.update()
        rec["started"] = sort_rows(started)
        rec["stopped"] = sort_rows(stopped)
        rec["removed"] = sort_rows(removed_events)
        rec["num"] = impl.num()
        rec["pairs"] = sort_rows(impl.pairs())
        sample = sorted(live)[:: max(1, len(live) // 40)]
        rec["proxy"] = np.array([impl.proxy(h) for h in sample], dtype=F32)
        rec["live"] = len(live)
        if n_queries:
            # world-level queries between updates: after a removal, so the tree of the last update is stale for it
            hs = np.array(sorted(h for h in live if impl.proxy(h) is not None), dtype=np.uint32)
            gone = impl.remove(hs[:: max(1, len(hs) // 7)][:5])
            for h in hs[:: max(1, len(hs) // 7)][:5].tolist():
                del live[h]
            rec["removed2"] = sort_rows(gone)
            qb = rand_boxes(n_queries)
            o = rng.uniform(0, side, size=(n_queries, 3)).astype(F32)
            d = rng.normal(0, 1, size=(n_queries, 3)).astype(F32)
            d[::5, 0] = 0  # axis-parallel rays exercise the zero-direction branch
            d[::7, 1:] = 0
            t = rng.uniform(1.0, 2 * side, size=(n_queries, 1)).astype(F32)
            t[::3] = np.finfo(F32).max
            pts = np.concatenate([rng.uniform(0, side, size=(n_queries - 4, 3)).astype(F32), qb[:4, :3]])  # on a box corner
            rec["q_aabb"] = sort_rows(impl.query(0, qb))
            rec["q_ray"] = sort_rows(impl.query(1, np.concatenate([o, d, t], axis=1)))
            rec["q_point"] = sort_rows(impl.query(2, pts))
        log.append(rec)
    return log


def assert_same_log(a, b, what=""):
    assert len(a) == len(b)
    for t, (x, y) in enumerate(zip(a, b)):
        for k in ("removed", "started", "stopped", "pairs") + tuple(q for q in ("removed2", "q_aabb", "q_ray", "q_point") if q in x):
            assert np.array_equal(x[k], y[k]), f"{what} step {t}: {k} differ ({len(x[k])} vs {len(y[k])})"
        assert x["num"] == y["num"], f"{what} step {t}: num_interferences"
        assert np.array_equal(x["proxy"].view(np.uint32), y["proxy"].view(np.uint32)), f"{what} step {t}: proxy boxes"
