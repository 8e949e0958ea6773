"""TriMesh ray casting, round-2 additions: the 4-wide BVH kernel against the binary-tree kernel (bit for bit), the chunked host
pipeline, RayCast::toi_and_normal_and_uv_with_ray (query/ray/ray_trimesh.rs:52-94) and per-ray max_toi against the oracle."""
import numpy as np
import pytest

from ncollide_b200.scenes import make_ray_scene


@pytest.fixture(scope="module")
def ctx():
    from ncollide_b200.world import Context

    c = Context(0)
    yield c
    c.close()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,posed", [("terrain", False), ("soup", True)])
def test_wide_kernel_equals_binary_kernel(ctx, monkeypatch, kind, posed):
    """300k rays (three chunks of the host pipeline) against 200k triangles: the 4-wide traversal, the binary traversal and the
    single-launch device entry return the same bytes."""
    import ctypes as C

    import torch

    from ncollide_b200 import _ffi
    from ncollide_b200.scenes import transform_rays

    rs = make_ray_scene(kind, 200_000, 300_000, seed=21, random_pose=posed)
    o, d = (transform_rays(rs.pose, rs.origins, rs.dirs) if posed else (rs.origins, rs.dirs))
    pose = np.ascontiguousarray(rs.pose, dtype=np.float32) if posed else None
    monkeypatch.setenv("NCB_RAY_WIDE", "1")
    wide = ctx.trimesh(rs.verts, rs.tris)
    monkeypatch.setenv("NCB_RAY_WIDE", "0")
    binary = ctx.trimesh(rs.verts, rs.tris)
    tw, fw, nw = wide.toi_and_normal_with_ray(pose, o, d)
    tb, fb, nb = binary.toi_and_normal_with_ray(pose, o, d)
    assert (tw >= 0).sum() > 10_000
    assert np.array_equal(fw, fb) and np.array_equal(bits(tw), bits(tb)) and np.array_equal(bits(nw), bits(nb))
    # one launch over device-resident rays
    n = len(o)
    d_o, d_d = torch.from_numpy(np.ascontiguousarray(o)).cuda(), torch.from_numpy(np.ascontiguousarray(d)).cuda()
    d_t, d_f, d_n = torch.empty(n, dtype=torch.float32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty((n, 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctx.check(ctx.lib.ncb_trimesh_ray_cast_device(wide.h, _ffi.ptr(pose), C.c_uint32(n), C.c_void_p(d_o.data_ptr()), C.c_void_p(d_d.data_ptr()),
                                                  C.c_float(np.finfo(np.float32).max), C.c_void_p(d_t.data_ptr()), C.c_void_p(d_f.data_ptr()),
                                                  C.c_void_p(d_n.data_ptr())), "ray_cast_device")
    ctx.synchronize()
    assert np.array_equal(bits(d_t.cpu().numpy()), bits(tw)) and np.array_equal(d_f.cpu().numpy().view(np.uint32), fw)
    assert np.array_equal(bits(d_n.cpu().numpy()), bits(nw))
    assert ctx.traversal_overflows() == 0
    wide.close()
    binary.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["terrain", "soup"])
def test_uv_ray_cast_vs_oracle(ctx, oracle, kind):
    rs = make_ray_scene(kind, 5000, 3000, seed=22)
    rng = np.random.default_rng(5)
    uvs = rng.random((len(rs.verts), 2)).astype(np.float32)
    mesh = ctx.trimesh(rs.verts, rs.tris)
    om = oracle.trimesh(rs.verts, rs.tris)
    # without uvs the reference falls back to toi_and_normal_with_ray: no uv
    toi0, face0, n0, uv0 = mesh.toi_and_normal_and_uv_with_ray(None, rs.origins, rs.dirs)
    assert not uv0.any()
    mesh.set_uvs(uvs)
    toi, face, normal, uv = mesh.toi_and_normal_and_uv_with_ray(None, rs.origins, rs.dirs)
    assert np.array_equal(face, face0) and np.array_equal(bits(toi), bits(toi0)) and np.array_equal(bits(normal), bits(n0))
    btoi, bface, bnormal, buv = om.ray_cast_uv(rs.origins, rs.dirs, uvs=uvs, mode=1)  # brute-force definition
    assert np.array_equal(face, bface) and np.array_equal(bits(toi), bits(btoi))
    assert np.array_equal(bits(uv), bits(buv)), f"{(bits(uv) != bits(buv)).sum()} uv words differ"
    hit = toi >= 0
    assert hit.sum() > 100 and uv[hit].min() >= -1e-5 and uv[hit].max() <= 1 + 1e-5
    # the reference-faithful best-first search with the uv visitor: same outside the tie class
    rtoi, rface, _, ruv = om.ray_cast_uv(rs.origins, rs.dirs, uvs=uvs, mode=0)
    same = face == rface
    assert (~same).sum() <= 5
    assert np.allclose(uv[same], ruv[same], rtol=1e-4, atol=1e-5)
    mesh.close()


@pytest.mark.gpu
def test_per_ray_max_toi(ctx, oracle):
    rs = make_ray_scene("terrain", 20_000, 6000, seed=23)
    mesh = ctx.trimesh(rs.verts, rs.tris)
    om = oracle.trimesh(rs.verts, rs.tris)
    toi, _, _ = mesh.toi_and_normal_with_ray(None, rs.origins, rs.dirs)
    rng = np.random.default_rng(6)
    limits = np.where(toi > 0, toi * rng.choice([0.5, 1.0, 1.5], size=len(toi)).astype(np.float32), np.float32(1.0)).astype(np.float32)
    t2, f2, n2 = mesh.toi_and_normal_with_ray(None, rs.origins, rs.dirs, max_toi=limits)
    bt, bf, bn, _ = om.ray_cast_uv(rs.origins, rs.dirs, max_toi=limits, mode=1)
    assert np.array_equal(f2, bf) and np.array_equal(bits(t2), bits(bt))
    assert ((t2 >= 0) & (t2 > limits)).sum() == 0
    assert ((toi >= 0) & (t2 < 0)).sum() > 100  # the halved limits cut hits away
    mesh.close()


def test_oracle_uv_is_barycentric_interpolation(oracle):
    """ORACLE check (CPU): uv == sum of the vertex uvs weighted by the barycentric coordinates of the hit point, computed here
    independently in f64 from the hit point (ray_trimesh.rs:76-84, ray_triangle.rs:113)."""
    rs = make_ray_scene("terrain", 2000, 800, seed=24)
    rng = np.random.default_rng(7)
    uvs = rng.random((len(rs.verts), 2)).astype(np.float32)
    om = oracle.trimesh(rs.verts, rs.tris)
    toi, face, normal, uv = om.ray_cast_uv(rs.origins, rs.dirs, uvs=uvs, mode=0)
    t0, f0, n0 = om.ray_cast(rs.origins, rs.dirs, mode=0)
    assert np.array_equal(toi, t0) and np.array_equal(face, f0) and np.array_equal(normal, n0)
    T = len(rs.tris)
    checked = 0
    for r in np.nonzero(toi >= 0)[0][:200]:
        tri = rs.tris[face[r] % T]
        a, b, c = (rs.verts[k].astype(np.float64) for k in tri)
        p = rs.origins[r].astype(np.float64) + rs.dirs[r].astype(np.float64) * float(toi[r])
        m = np.stack([b - a, c - a], axis=1)
        vw, *_ = np.linalg.lstsq(m, p - a, rcond=None)
        want = uvs[tri[0]] * (1 - vw.sum()) + uvs[tri[1]] * vw[0] + uvs[tri[2]] * vw[1]
        assert np.allclose(uv[r], want, atol=2e-3), (r, uv[r], want)
        checked += 1
    assert checked > 50
