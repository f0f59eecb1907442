"""The oracle against every known-answer vector the reference's own tests hold for this path
(SURVEY.md §4 / §8c): geometry.rs test_vec3 / test_bound3 / test_bnd2, sphere.rs test_sphere,
and the geometry + rays of primitives.rs test_primitive; plus hand-derived Morton / radix cases
for bvh.rs (its own test asserts nothing)."""
import numpy as np

import oracle_lib as O
from rs_ray_toy_b200 import synth


def test_vec3_kat():
    # geometry.rs:1922-1946
    out = np.zeros(8)
    O.lib().orc_kat_vec3(np.array([3.0, 4.0, -5.0]), np.array([8.1, 10.8, -13.5]), 2.7, out)
    assert out[0] == 50.0
    # the reference compares with nearly-equal PartialEq (geometry.rs SMALL); exact here to 1 ulp
    assert np.allclose(out[1:4], [8.1, 10.8, -13.5], rtol=0, atol=1e-12)
    assert abs(out[4] - 135.0) <= 1e-9
    out2 = np.zeros(8)
    O.lib().orc_kat_vec3(np.array([3.0, 4.0, -5.0]), np.array([9.6, -12.4, 3.7]), 1.0, out2)
    assert np.allclose(out2[5:8], [-47.2, -59.1, -75.6], rtol=0, atol=1e-12)


def test_bound3_kat():
    # geometry.rs:1949-1971
    out = np.zeros(10)
    O.lib().orc_kat_bounds(np.array([0.0, -10.0, 5.0]), np.array([-10.0, 20.0, 10.0]), np.array([-15.0, 10.0, 30.0]), out)
    assert out[:6].tolist() == [-15.0, -10.0, 5.0, 0.0, 20.0, 30.0]
    out = np.zeros(10)
    # bounding sphere of `a` itself: union with a point already inside leaves it unchanged
    O.lib().orc_kat_bounds(np.array([0.0, -10.0, 5.0]), np.array([-10.0, 20.0, 10.0]), np.array([-5.0, 5.0, 7.5]), out)
    assert out[6:9].tolist() == [-5.0, 5.0, 7.5]


def test_sphere_kat():
    # shape/sphere.rs:309-316 — origin lies ON the sphere; intersect_p is true through t1
    s = O.OracleScene(O.TIER_L)
    m, inv = O.make_to_world(world_pos=(1.0, 0.0, 0.0))
    sp = s.add_sphere(m, inv, 1.0, -1.0, 1.0, 360.0)
    g = s.add_geo_sphere(sp)
    s.add_prims(g, 1, -1)
    d = O.normalize([1.0, 0.0, 0.0])
    assert s.prim_intersect_p(0, [0, 0, 0, d[0], d[1], d[2], np.inf])


def _test_primitive_scene(tier):
    # primitives.rs:151-238: explicit cube index list, 3 translated instances, 3 rays
    vi = np.array([0, 4, 6, 4, 6, 2, 3, 2, 6, 2, 6, 7, 7, 6, 4, 6, 4, 5, 5, 1, 3, 1, 3, 7, 1, 0, 2, 0,
                   2, 3, 5, 4, 0, 4, 0, 1], dtype=np.uint32).reshape(12, 3)
    s = O.OracleScene(tier)
    mesh = s.add_mesh(synth.CUBE_P, vi)
    g0 = s.add_geo_triangles(mesh)
    for pos in [(10.0, 10.0, 15.0), (15.0, 10.0, 15.0), (3.0, 3.0, 3.0)]:
        m, inv = O.make_to_world(world_pos=pos)
        s.add_prims(g0, 12, s.add_xform(m, inv))
    rays = []
    for d in [(0.8, 1.0, 0.8), (1.0, 0.9, 0.9), (1.0, 1.0, 1.0)]:
        n = O.normalize(d)
        rays.append([0, 0, 0, n[0], n[1], n[2], np.inf])
    return s, np.array(rays)


def test_primitive_kat():
    s, rays = _test_primitive_scene(O.TIER_L)
    hits = 0
    for i in range(3):
        for k in range(12):
            hits += s.prim_intersect_p(12 * i + k, rays[i])
    assert hits > 0  # the reference's only assertion (primitives.rs:237)
    # Geometry says more: ray 2 (1,1,1) pierces the cube at (3,3,3) (entering + leaving face),
    # ray 0 / ray 1 miss their cubes at (10,10,15) / (15,10,15).
    sf, _ = _test_primitive_scene(O.TIER_F)
    per_ray = [sum(sf.prim_intersect_p(12 * i + k, rays[i]) for k in range(12)) for i in range(3)]
    assert per_ray[2] >= 2 and per_ray[0] == 0 and per_ray[1] == 0


def test_bnd2_tile_iteration():
    """geometry.rs:1974-1980 (test_bnd2): every point `Bounds2i::into_iter` yields is `inside` the bound.  The oracle's
    Bounds2iIter (the iterator its render loop walks a tile with) must also yield the row-major walk of
    [48,64) x [0,16), which is what geometry.rs:1494-1526 produces."""
    import ctypes as C
    L = O.lib()
    L.orc_kat_bounds2i_iter.restype = C.c_uint64
    L.orc_kat_bounds2i_iter.argtypes = [C.c_int64] * 4 + [C.c_void_p, C.c_uint64]
    out = np.zeros((400, 3), dtype=np.int64)
    n = L.orc_kat_bounds2i_iter(48, 0, 64, 16, out.ctypes.data, 400)
    assert n == 256 and (out[:n, 2] == 1).all()
    assert [tuple(r) for r in out[:n, :2]] == [(x, y) for y in range(0, 16) for x in range(48, 64)]
    # ragged last tile of a 640 x 360 image (rows 352..359) and a one-pixel bound
    n = L.orc_kat_bounds2i_iter(624, 352, 640, 360, out.ctypes.data, 400)
    assert n == 128 and tuple(out[0, :2]) == (624, 352) and tuple(out[n - 1, :2]) == (639, 359)
    assert L.orc_kat_bounds2i_iter(5, 7, 6, 8, out.ctypes.data, 400) == 1 and tuple(out[0, :2]) == (5, 7)


def test_left_shift3_and_morton():
    L = O.lib()
    # bvh.rs:17-32 — spreads the low 10 bits to every third position; 1<<10 saturates to 1023
    assert L.orc_kat_left_shift3(1) == 1
    assert L.orc_kat_left_shift3(0b11) == 0b1001
    assert L.orc_kat_left_shift3(1023) == 0b1001001001001001001001001001
    assert L.orc_kat_left_shift3(1024) == L.orc_kat_left_shift3(1023)
    # bvh.rs:34-39 — z in bit 2, y in bit 1, x in bit 0; `as u32` truncates
    assert L.orc_kat_morton(1.9, 0.0, 0.0) == 1
    assert L.orc_kat_morton(0.0, 1.0, 0.0) == 2
    assert L.orc_kat_morton(0.0, 0.0, 1.0) == 4
    assert L.orc_kat_morton(3.0, 3.0, 3.0) == 0b111111
    assert L.orc_kat_morton(-5.0, float("nan"), 0.0) == 0  # saturating casts


def test_radix_sort_is_stable_sort_on_30_bits():
    rng = np.random.default_rng(7)
    n = 5000
    codes = rng.integers(0, 1 << 30, size=n, dtype=np.uint32)
    codes[::7] = codes[0]  # duplicates: stability is observable
    idx = np.arange(n, dtype=np.uint32)
    c2, i2 = codes.copy(), idx.copy()
    O.lib().orc_kat_radix_sort(n, i2.ctypes.data, c2.ctypes.data)
    order = np.argsort(codes, kind="stable")
    assert (c2 == codes[order]).all() and (i2 == idx[order]).all()


def test_m44_inverse_roundtrip():
    m, inv = O.make_to_world((3.0, -2.0, 5.5), (1.0, 2.0, 3.0), 37.0, (1.0, 1.0, 1.0))
    out = np.zeros(16)
    O.lib().orc_m44_inverse(m.reshape(16).copy(), out)
    assert np.allclose(out.reshape(4, 4), inv, atol=1e-12)
    assert np.allclose(m @ inv, np.eye(4), atol=1e-12)
