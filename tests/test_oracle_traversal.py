"""CPU-only checks of the oracle's BVH restatement (bvh.rs): the Tier-F tree against an O(N*R)
brute force that knows no BVH, topology invariants of the flattened array, the literal tier's
quirks showing up where SURVEY.md Appendix A says they do, and the committed golden vectors."""
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
import scenes
from rs_ray_toy_b200 import synth, transform

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_tier_f_matches_brute_force_soup():
    p, idx = scenes.soup(20000)
    rays = synth.bounce_rays(p, idx, 4000)
    s = scenes.oracle_soup(p, idx)
    r = s.intersect(rays)
    bp, bt = s.brute_force(rays)
    assert (bp == r["prim"]).all()
    assert (bt == r["t"]).all()
    assert r["stats"][4] == 0  # no stack overflow (Q27)


def test_tier_f_any_hit_consistent_with_closest():
    p, idx = scenes.soup(20000)
    rays = synth.bounce_rays(p, idx, 4000)
    s = scenes.oracle_soup(p, idx)
    r = s.intersect(rays)
    occ, _ = s.intersect_p(rays)
    assert ((r["prim"] >= 0) == (occ == 1)).all()
    sh = synth.shadow_rays_from(rays, (0.5, 0.5, 2.0))
    occ2, _ = s.intersect_p(sh)
    r2 = s.intersect(sh)
    assert ((r2["prim"] >= 0) == (occ2 == 1)).all()


def test_flattened_tree_invariants():
    p, idx = scenes.soup(5000)
    s = scenes.oracle_soup(p, idx)
    bounds, meta, ordered = s.nodes()
    n = len(meta)
    # depth-first layout: first child is my+1, second child index is stored (bvh.rs:728-751)
    seen = np.zeros(len(ordered), dtype=int)
    stack = [0]
    visited = 0
    while stack:
        i = stack.pop()
        visited += 1
        off, cnt, axis = meta[i]
        if cnt > 0:
            seen[off:off + cnt] += 1
        else:
            assert i + 1 < n and off < n and axis < 3
            for c in (i + 1, off):
                assert (bounds[c, :3] >= bounds[i, :3]).all() and (bounds[c, 3:] <= bounds[i, 3:]).all()
            stack += [i + 1, off]
    assert visited == n
    assert (seen == 1).all()                       # Tier F: every slot in exactly one leaf
    assert sorted(ordered.tolist()) == list(range(5000))  # ... and no primitive lost (Q1 fixed)


def test_literal_tier_drops_primitives_q1():
    # Q1 (bvh.rs:588-607): treelets that split duplicate their head and lose their tail.
    m, inv = scenes.cube_instances(200, extent=20.0)
    lit = scenes.oracle_cubes(m, inv, tier=O.TIER_L)
    _, _, ordered = lit.nodes()
    present = set(ordered.tolist())
    assert len(present) < lit.num_prims            # geometry silently lost
    fix = scenes.oracle_cubes(m, inv, tier=O.TIER_F)
    _, _, ordered_f = fix.nodes()
    assert sorted(ordered_f.tolist()) == list(range(fix.num_prims))


def test_python_transform_matches_oracle_bits():
    rng = np.random.default_rng(3)
    for _ in range(50):
        pos, axis, ang = rng.uniform(-50, 50, 3), rng.normal(size=3), rng.uniform(0, 360)
        m, inv = transform.make_to_world(pos, axis, ang)
        om, oinv = O.make_to_world(pos, axis, ang)
        assert (m == om).all() and (inv == oinv).all()
    prm = synth.instance_params(64, 50.0, 9)
    bm, binv = transform.make_to_world_batch(prm["world_pos"], prm["axis"], prm["angle"])
    for i in range(64):
        om, oinv = O.make_to_world(prm["world_pos"][i], prm["axis"][i], prm["angle"][i])
        assert (bm[i] == om).all() and (binv[i] == oinv).all()


def test_instanced_cubes_and_spheres_match_brute_force():
    m, inv = scenes.cube_instances(300, extent=20.0)
    s = scenes.oracle_cubes(m, inv)
    rays = synth.camera_like_rays(3000, (0.0, 0.0, -60.0), 20.0)
    r = s.intersect(rays)
    bp, bt = s.brute_force(rays)
    assert (bp == r["prim"]).all() and (bt == r["t"]).all()
    assert (r["prim"] >= 0).sum() > 300
    ms, invs = scenes.sphere_instances(500, extent=10.0)
    ss = scenes.oracle_spheres(ms, invs, radius=0.5)
    rays = synth.camera_like_rays(3000, (0.0, 0.0, -30.0), 10.0)
    r = ss.intersect(rays)
    bp, bt = ss.brute_force(rays)
    assert (bp == r["prim"]).all() and (bt == r["t"]).all()
    assert (r["prim"] >= 0).sum() > 300


@pytest.mark.parametrize("name", ["soup_c3_small", "cubes_c2_small", "spheres_c4_small"])
def test_golden_vectors(name):
    """Committed fixtures (tests/golden/make_golden.py wrote them from this oracle): a change in
    the oracle's arithmetic shows up here, on CPU, before any GPU parity run."""
    import golden_cases
    g = np.load(GOLDEN / f"{name}.npz")
    s, rays = golden_cases.build_oracle(name)
    assert (rays == g["rays"]).all()
    r = s.intersect(rays)
    occ, _ = s.intersect_p(golden_cases.shadow_rays(name, rays))
    assert (r["prim"] == g["prim"]).all()
    assert (r["t"] == g["t"]).all()
    assert (r["uv"] == g["uv"]).all()
    assert (occ == g["occluded"]).all()


def test_product_literal_hlbvh_builder_matches_the_oracle():
    """The product's own restatement of BVHAccel::new / HLBVH (csrc/bvh_hlbvh.cpp, host only) against the
    oracle's literal tier: same flattened nodes (bounds, offsets, axes) and same reordered primitive
    list, Q1 duplicates and the Q2-skewed upper tree included."""
    import ctypes as C
    from rs_ray_toy_b200 import capi
    L = capi.lib()
    L.rrt_hlbvh_literal_probe.restype = C.c_int
    L.rrt_hlbvh_literal_probe.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32),
                                          C.c_void_p, C.c_void_p, C.c_void_p]
    for n, edge, max_prims in ((37, 0.2, 4), (5000, 0.02, 4), (20000, 0.01, 4), (3000, 0.05, 2), (3000, 0.05, 7)):
        p, idx = scenes.soup(n, edge=edge, seed=11 + n)
        ref = scenes.oracle_soup(p, idx, tier=O.TIER_L, max_prims=max_prims)
        rb, rm, ro = ref.nodes()
        tri = p[idx]                                   # [n, 3, 3]
        bounds = np.concatenate([tri.min(axis=1), tri.max(axis=1)], axis=1).copy()
        cap = 2 * n + 8
        nb, nm, no = np.zeros((cap, 6)), np.zeros((cap, 3), dtype=np.uint32), np.zeros(n, dtype=np.uint32)
        cnt = C.c_uint32()
        capi.check(L.rrt_hlbvh_literal_probe(n, bounds.ctypes.data, max_prims, cap, C.byref(cnt), nb.ctypes.data,
                                             nm.ctypes.data, no.ctypes.data))
        assert cnt.value == len(rm)
        assert np.array_equal(nm[: cnt.value], rm)
        assert np.array_equal(nb[: cnt.value], rb)
        assert np.array_equal(no, ro)
        if n >= 3000 and max_prims == 4:
            assert len(set(ro.tolist())) < n           # Q1 really dropped primitives here
