"""Host-buffer ray cast: PCIe copy bandwidth on this box and the chunk-size sweep of ncb_trimesh_ray_cast_uv's pipeline.
python scripts/ray_e2e_sweep.py  (NCB_RAY_CHUNK is read once per process: the sweep re-runs this script per value)"""
import json, os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def one():
    import torch
    from ncollide_b200.scenes import make_ray_scene
    from ncollide_b200.world import Context
    ctx = Context(0)
    n = 1_000_000
    rs = make_ray_scene("terrain", n, n, seed=1004)
    mesh = ctx.trimesh(rs.verts, rs.tris)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    o_t, d_t = pin(rs.origins), pin(rs.dirs)
    out_t = {"toi": torch.empty(n, dtype=torch.float32).pin_memory(), "face": torch.empty(n, dtype=torch.int32).pin_memory(),
             "normal": torch.empty((n, 3), dtype=torch.float32).pin_memory()}
    out = {"toi": out_t["toi"].numpy(), "face": out_t["face"].numpy().view(np.uint32), "normal": out_t["normal"].numpy()}
    for _ in range(3):
        mesh.toi_and_normal_with_ray(None, o_t.numpy(), d_t.numpy(), out=out)
    t0 = time.perf_counter()
    for _ in range(10):
        mesh.toi_and_normal_with_ray(None, o_t.numpy(), d_t.numpy(), out=out)
    ms = (time.perf_counter() - t0) * 100
    res = {"chunk": os.environ.get("NCB_RAY_CHUNK", "default"), "e2e_ms": ms, "Mrays_s": n / ms / 1e3}
    if os.environ.get("RAY_BW"):
        dev = torch.empty(24_000_000, dtype=torch.uint8, device="cuda")
        host = torch.empty(24_000_000, dtype=torch.uint8).pin_memory()
        for name, fn in (("h2d", lambda: dev.copy_(host, non_blocking=True)), ("d2h", lambda: host.copy_(dev, non_blocking=True))):
            fn(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            res[name + "_GBps_24MB"] = 24e6 * 10 / (time.perf_counter() - t0) / 1e9
    print(json.dumps(res))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for i, c in enumerate(("1000000", "524288", "262144", "131072", "65536")):
            env = dict(os.environ, NCB_RAY_CHUNK=c)
            if i == 0:
                env["RAY_BW"] = "1"
            subprocess.run([sys.executable, __file__, "one"], env=env)
