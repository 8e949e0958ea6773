"""GPU tests of two boundary behaviours (first seen green on the driver's B200 box at the end of round 1):

  * proximity sensors through the spatially sharded multi-GPU update (ncb_world_update_sharded), replayed rank by rank on one device
    like tests/test_gpu_parity.py::test_spatial_shards_partition_the_pair_set — the sensor path re-keys pairs by GLOBAL handle with
    replicated query kinds;
  * the boundary refusing shape types the device does not know."""
import numpy as np
import pytest

from ncollide_b200.scenes import config_scene, make_world_scene, with_sensors

pytestmark = pytest.mark.gpu


def canon(pairs):
    p = np.sort(np.asarray(pairs, dtype=np.uint32).reshape(-1, 2), axis=1)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


@pytest.mark.parametrize("world,mk", [(2, lambda: config_scene(3, 7001)), (4, lambda: make_world_scene(6000, 92, (1, 1, 1), side=9.0, plane=True, n_hulls=24))])
def test_sharded_update_with_sensors(oracle, world, mk):
    from ncollide_b200.world import Context

    s = with_sensors(mk(), 0.3, 17, margin=0.1)
    ctx = Context(0)
    ctx.set_scene(s)
    full = ctx.world_fetch(ctx.world_update_device(s.margin))
    want = {tuple(p): (int(a), int(st), int(c)) for p, a, st, c in zip(full.pairs.tolist(), full.pair_algo, full.proximity, full.manifold_count)}
    assert (full.pair_algo == 6).any()
    seen = {}
    for rank in range(world):
        c = ctx.world_update_sharded(s.margin, rank, world)
        r = ctx.world_fetch(c)
        assert r.proximity is not None
        assert c["n_algo"]["proximity"] == int((r.pair_algo == 6).sum())
        for p, a, st, cnt in zip(map(tuple, r.pairs.tolist()), r.pair_algo, r.proximity, r.manifold_count):
            assert p not in seen, "pair reported by two ranks"
            seen[p] = (int(a), int(st), int(cnt))
    assert seen == want, "union of the ranks' (algorithm, proximity status, manifold size) differs from the single-GPU update"
    # and the single-GPU statuses are the oracle's
    o_c, o_off, o_algo, o_prox = oracle.narrow_phase_kinds(s, full.pairs)
    assert np.array_equal(full.proximity, o_prox) and np.array_equal(full.pair_algo, o_algo)
    ctx.close()


def test_device_rejects_shapes_it_does_not_know():
    """A shape_type the device has no generator for (5 and up: composite shapes) must be refused at the boundary, not masked into
    another shape; so must a convex-hull object that names a hull outside the uploaded library."""
    from ncollide_b200._ffi import NcbError
    from ncollide_b200.world import Context

    s = make_world_scene(50, 1, (1, 1, 0), side=3.0)
    ctx = Context(0)
    ctx.set_scene(s)  # fine
    s.shape_type = s.shape_type.copy()
    s.shape_type[7] = 5
    with pytest.raises(NcbError, match="shape_type"):
        ctx.set_objects(s)
    s.shape_type[7] = 2  # a convex hull ...
    s.shape_param[7, 0] = 1.0e6  # ... that is not in the library
    with pytest.raises(NcbError, match="hull id"):
        ctx.set_objects(s)
    bad = np.array([[0, s.n + 5]], dtype=np.uint32)
    ctx.set_scene(make_world_scene(50, 1, (1, 1, 0), side=3.0))
    with pytest.raises(NcbError, match="out of range"):
        ctx.generate_contacts(bad)
    ctx.close()
