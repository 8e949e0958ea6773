"""Single-GPU timing of one rank's share of a spatially sharded update (rank 0 of `world`) on an N-object scene:
python scripts/bench_shard.py [N] [world]"""
import sys

sys.path.insert(0, ".")
from ncollide_b200.scenes import config_scene  # noqa: E402
from ncollide_b200.world import Context  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
s = config_scene(3, n)
ctx = Context(0)
ctx.set_scene(s)
ctx.profile_enable(True)
for it in range(4):
    c = ctx.world_update_sharded(s.margin, it % world, world)
    ctx.synchronize()
    print(it % world, c["n_pairs"], [(k, round(v, 3)) for k, v, _ in ctx.profile_get()])
