"""Two fresh-world updates of the cfg3 scene with 20 % sensors, for an ncu launch list (scripts usage:
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/prox_launches.csv python scripts/prox_profile.py)."""
import sys

sys.path.insert(0, ".")
from ncollide_b200.scenes import config_scene, with_sensors  # noqa: E402
from ncollide_b200.world import Context  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
s = with_sensors(config_scene(3, n), 0.2, 5)
ctx = Context(0)
ctx.set_scene(s)
for _ in range(2):
    c = ctx.world_update_device(s.margin)
print(c["n_algo"], c["n_proximity"])
