"""-m gpu: the CUDA aggregate (through the C ABI) against the CPU oracle and the committed
golden vectors.  Bar (BASELINE.json north_star): closest-hit primitive index bit-exact except
rays whose candidates tie within 1e-6 relative; t within 1e-5 relative (the kernels decide in
f64 with the reference's operation order, so t, u, v are expected to be bit-identical and the
tests record how many are); any-hit flags bit-exact."""
from pathlib import Path

import numpy as np
import pytest

import golden_cases
import oracle_lib as O
import scenes
from rs_ray_toy_b200 import capi, synth

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
REL_T = 1e-5
REL_TIE = 1e-6


def _assert_closest(hits, prim, t, uv=None):
    c = scenes.compare_closest(hits, prim, t, REL_TIE, REL_T)
    assert c["mismatch_excl_ties"] == 0, c
    assert c["t_bad"] == 0, c
    if uv is not None:
        ok = (hits["prim_id"].astype(np.int64) == prim) & (prim >= 0)
        assert np.allclose(hits["u"][ok], uv[ok, 0], rtol=1e-5, atol=1e-9)
        assert np.allclose(hits["v"][ok], uv[ok, 1], rtol=1e-5, atol=1e-9)
    return c


@pytest.mark.parametrize("name", list(golden_cases.CASES))
def test_golden_closest_and_any_hit(ctx, name):
    g = np.load(GOLDEN / f"{name}.npz")
    agg, rays = golden_cases.build_gpu(ctx, name)
    assert (rays == g["rays"]).all()
    hits = agg.intersect(rays)
    c = _assert_closest(hits, g["prim"], g["t"], g["uv"])
    assert c["mismatch"] == 0, c                    # no ties in the fixtures: exact
    if golden_cases.CASES[name]["kind"] == "cubes":
        # rotated instances: the oracle intersects in instance space (primitives.rs:126-139), the
        # device in world space with baked f64 vertices -> same hit, t equal to rounding only
        assert c["max_rel_dt"] < 1e-11, c
    else:
        assert c["t_exact"] == c["hits"], c          # f64 deciding arithmetic: same bits
    occ = agg.intersect_p(golden_cases.shadow_rays(name, rays))
    assert (occ == g["occluded"]).all()


def test_soup_vs_oracle_200k(ctx):
    import oracle_lib as O
    p, idx = scenes.soup(200000)
    rays = synth.bounce_rays(p, idx, 200000, seed=41)
    ref = scenes.oracle_soup(p, idx).intersect(rays)
    agg = scenes.gpu_soup(ctx, p, idx)
    hits = agg.intersect(rays)
    c = _assert_closest(hits, ref["prim"], ref["t"], ref["uv"])
    assert c["hits"] > 50000
    sh = synth.shadow_rays_from(rays, (0.5, 0.5, 1.5))
    occ_ref, _ = scenes.oracle_soup(p, idx).intersect_p(sh)
    assert (agg.intersect_p(sh) == occ_ref).all()


def test_edge_cases(ctx):
    import oracle_lib as O
    p, idx = scenes.soup(5000)
    agg = scenes.gpu_soup(ctx, p, idx)
    ref = scenes.oracle_soup(p, idx)
    # empty batch
    assert agg.intersect(np.zeros((0, 7))).shape == (0,)
    assert agg.intersect_p(np.zeros((0, 7))).shape == (0,)
    # ragged sizes around the block / chunk boundaries, finite t_max, far origins, axis-parallel
    # directions (zero components -> infinite slabs), rays pointing away
    rng = np.random.default_rng(5)
    for n in (1, 31, 127, 129, 1000):
        rays = synth.bounce_rays(p, idx, n, seed=100 + n)
        rays[:, 6] = np.where(rng.random(n) < 0.5, rng.uniform(0.001, 0.3, n), np.inf)
        r = ref.intersect(rays)
        _assert_closest(agg.intersect(rays), r["prim"], r["t"])
    rays = np.zeros((600, 7))
    rays[:, 6] = np.inf
    rays[:200, 0:3] = rng.uniform(-1000, 1000, (200, 3))           # far outside the world box
    tgt = rng.uniform(0, 1, (200, 3))
    d = tgt - rays[:200, 0:3]
    rays[:200, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[200:400, 0:3] = rng.uniform(0, 1, (200, 3))                # axis-parallel
    ax = rng.integers(0, 3, 200)
    rays[200 + np.arange(200), 3 + ax] = rng.choice([-1.0, 1.0], 200)
    rays[400:, 0:3] = rng.uniform(2, 3, (200, 3))                   # pointing away
    rays[400:, 3:6] = np.array([1.0, 0.0, 0.0])
    r = ref.intersect(rays)
    c = _assert_closest(agg.intersect(rays), r["prim"], r["t"])
    assert c["hits"] > 10
    occ_ref, _ = ref.intersect_p(rays)
    assert (agg.intersect_p(rays) == occ_ref).all()
    # unnormalised directions: t scales, the hit does not change
    rays2 = synth.bounce_rays(p, idx, 2000, seed=77)
    rays2[:, 3:6] *= 3.0
    r = ref.intersect(rays2)
    _assert_closest(agg.intersect(rays2), r["prim"], r["t"])


def test_single_primitive_and_one_leaf_scenes(ctx):
    import oracle_lib as O
    from rs_ray_toy_b200.aggregate import GpuAggregate
    for ntri in (1, 3, 4, 5):
        p, idx = scenes.soup(ntri, edge=0.3, seed=9)
        agg = scenes.gpu_soup(ctx, p, idx)
        ref = scenes.oracle_soup(p, idx)
        rays = synth.camera_like_rays(3000, (0.5, 0.5, -2.0), 0.6, seed=3)
        rays[:, 0:3] += 0.0
        rays[:, 3:6] = (np.random.default_rng(1).uniform(0, 1, (3000, 3)) - rays[:, 0:3])
        rays[:, 3:6] /= np.linalg.norm(rays[:, 3:6], axis=1, keepdims=True)
        r = ref.intersect(rays)
        _assert_closest(agg.intersect(rays), r["prim"], r["t"])


def test_empty_scene_is_an_error(ctx):
    from rs_ray_toy_b200 import capi
    from rs_ray_toy_b200.aggregate import GpuAggregate
    a = GpuAggregate(ctx)
    with pytest.raises(capi.RrtError) as e:
        a.commit()
    assert e.value.status == capi.RRT_ERR_EMPTY     # bvh.rs:319 asserts


def test_instanced_cubes_c2_vs_oracle(ctx):
    m, inv = scenes.cube_instances(10000)            # config 2's 120,000 TransformedPrimitives
    agg = scenes.gpu_cubes(ctx, m, inv)
    ref = scenes.oracle_cubes(m, inv)
    assert agg.num_prims == 120000
    assert np.allclose(agg.world_bound(), ref.world_bound(), rtol=0, atol=0)
    rays = synth.camera_like_rays(300000, (0.0, 0.0, -200.0), 50.0)
    r = ref.intersect(rays)
    c = _assert_closest(agg.intersect(rays), r["prim"], r["t"])
    assert c["hits"] > 100000


def test_sphere_field_c4_vs_oracle(ctx):
    m, inv = scenes.sphere_instances(100000)          # config 4's sphere field
    agg = scenes.gpu_spheres(ctx, m, inv)
    ref = scenes.oracle_spheres(m, inv)
    rays = synth.camera_like_rays(300000, (0.0, 0.0, -120.0), 50.0)
    r = ref.intersect(rays)
    c = _assert_closest(agg.intersect(rays), r["prim"], r["t"])
    assert c["hits"] > 100000
    # secondary rays leaving sphere surfaces: the self-hit policy (Q8) must agree
    hit = r["prim"] >= 0
    o = rays[hit, 0:3] + rays[hit, 3:6] * r["t"][hit, None]
    centre = m[r["prim"][hit], 0:3, 3]
    nrm = (o - centre) / 0.5
    rng = np.random.default_rng(2)
    d = synth.random_unit_vectors(o.shape[0], rng)
    d[np.sum(d * nrm, axis=1) < 0] *= -1.0
    sec = np.concatenate([o, d, np.full((o.shape[0], 1), np.inf)], axis=1)
    r2 = ref.intersect(sec)
    _assert_closest(agg.intersect(sec), r2["prim"], r2["t"])
    occ_ref, _ = ref.intersect_p(sec)
    assert (agg.intersect_p(sec) == occ_ref).all()


def test_full_size_properties_c3(ctx):
    """BASELINE config 3 at full size (1M triangles, 16M rays) through size-independent
    properties: (1) a 64k-ray slice equals the oracle; (2) any-hit == (closest-hit found
    something); (3) re-tracing every hit ray with t_max = t*(1+1e-9) returns the same primitive
    and t (idempotence); (4) with t_max = t*(1-1e-6) nothing is hit: no primitive lies in front
    of the reported closest hit."""
    p, idx = scenes.soup(1 << 20)
    agg = scenes.gpu_soup(ctx, p, idx)
    n = 1 << 24
    rays = synth.bounce_rays(p, idx, n)
    hits = agg.intersect(rays)
    hit = hits["prim_id"] != 0xFFFFFFFF
    assert 0.3 < hit.mean() < 1.0
    ref = scenes.oracle_soup(p, idx).intersect(rays[: 1 << 16])
    _assert_closest(hits[: 1 << 16], ref["prim"], ref["t"], ref["uv"])
    occ = agg.intersect_p(rays)
    assert (occ.astype(bool) == hit).all()
    again = rays.copy()
    again[hit, 6] = hits["t"][hit] * (1.0 + 1e-9)
    h2 = agg.intersect(again)
    assert (h2["prim_id"] == hits["prim_id"]).all() and (h2["t"][hit] == hits["t"][hit]).all()
    short = rays[hit].copy()
    short[:, 6] = hits["t"][hit] * (1.0 - 1e-6)
    h3 = agg.intersect(short)
    got = h3["prim_id"] != 0xFFFFFFFF
    assert not got.any(), int(got.sum())   # nothing lies in front of the closest hit


def test_degenerate_rays_hit_nothing(ctx):
    """Rays the reference would refuse (scene.rs:70 asserts d != 0) or that carry NaN / inf: the
    device reports a miss for them and is not disturbed for their neighbours in the batch."""
    p, idx = scenes.soup(5000)
    agg = scenes.gpu_soup(ctx, p, idx)
    rays = synth.bounce_rays(p, idx, 4096 + 64, seed=31)
    good = agg.intersect(rays)
    bad = rays.copy()
    bad[0, 3:6] = 0.0
    bad[1, 3] = np.nan
    bad[2, 0] = np.nan
    bad[3, 1] = np.inf
    bad[4, 3:6] = (np.inf, 0.0, 0.0)
    bad[5, 6] = -1.0          # negative t_max
    bad[6, 6] = 0.0
    h = agg.intersect(bad)
    assert (h["prim_id"][:7] == 0xFFFFFFFF).all()
    assert np.array_equal(h[7:], good[7:])
    assert (agg.intersect_p(bad)[:7] == 0).all()


def test_call_order_and_argument_errors(ctx):
    """No panic / exception crosses the ABI: misuse comes back as a status with a message
    (the reference asserts: bvh.rs:319, scene.rs:70,77)."""
    import ctypes as C
    from rs_ray_toy_b200 import capi
    from rs_ray_toy_b200.aggregate import GpuAggregate
    a = GpuAggregate(ctx)
    with pytest.raises(capi.RrtError) as e:     # intersect before commit
        a.intersect(np.zeros((4, 7)))
    assert e.value.status == capi.RRT_ERR_INVALID
    with pytest.raises(capi.RrtError):          # bad mesh id
        a.add_triangles(7, 0)
    with pytest.raises(capi.RrtError):          # vertex index out of range
        a.add_mesh(np.zeros((3, 3)), np.array([[0, 1, 5]], dtype=np.uint32))
    p, idx = scenes.soup(100)
    m = a.add_mesh(p, idx)
    a.add_triangles(m, 0)
    a.commit()
    with pytest.raises(capi.RrtError):          # committed scenes are immutable
        a.add_triangles(m, 0)
    with pytest.raises(capi.RrtError):
        a.commit()
    L = capi.lib()
    assert L.rrt_intersect(a.h, 4, None, None) == capi.RRT_ERR_INVALID
    assert L.rrt_last_error()
    with pytest.raises(capi.RrtError) as e:     # unknown build flags
        b = GpuAggregate(ctx)
        mb = b.add_mesh(p, idx)
        b.add_triangles(mb, 0)
        b.commit(4, 7)
    assert e.value.status == capi.RRT_ERR_INVALID


def test_concurrent_host_threads_share_one_aggregate(ctx):
    """`Primitive: Send + Sync` (primitives.rs:14): several host threads query one committed
    aggregate at once (the reference's rayon workers do)."""
    import threading
    p, idx = scenes.soup(20000)
    agg = scenes.gpu_soup(ctx, p, idx)
    batches = [synth.bounce_rays(p, idx, 30000, seed=200 + k) for k in range(4)]
    expect = [agg.intersect(b) for b in batches]
    got = [None] * 4
    occ = [None] * 4

    def work(k):
        for _ in range(3):
            got[k] = agg.intersect(batches[k])
            occ[k] = agg.intersect_p(batches[k])

    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for k in range(4):
        assert np.array_equal(got[k], expect[k])
        assert np.array_equal(occ[k].astype(bool), expect[k]["prim_id"] != 0xFFFFFFFF)


def test_device_pointer_entry_points(ctx):
    """rrt_intersect_device / rrt_intersect_p_device on caller-owned device buffers and stream."""
    import torch
    from rs_ray_toy_b200.aggregate import HIT_DTYPE, pack_rays
    p, idx = scenes.soup(30000)
    agg = scenes.gpu_soup(ctx, p, idx)
    rays = synth.bounce_rays(p, idx, 50000, seed=9)
    ref = agg.intersect(rays)
    d_rays = torch.from_numpy(pack_rays(rays).view(np.float64).copy()).cuda()
    d_hits = torch.zeros(50000 * 4, dtype=torch.float64, device="cuda")
    d_occ = torch.zeros(50000, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        agg.intersect_device(50000, d_rays.data_ptr(), d_hits.data_ptr(), s.cuda_stream)
        agg.intersect_p_device(50000, d_rays.data_ptr(), d_occ.data_ptr(), s.cuda_stream)
    s.synchronize()
    hits = d_hits.cpu().numpy().view(HIT_DTYPE).reshape(-1)
    assert np.array_equal(hits, ref)
    assert np.array_equal(d_occ.cpu().numpy().astype(bool), ref["prim_id"] != 0xFFFFFFFF)


@pytest.mark.parametrize("n", [1, 4097, 131072 + 5, 1048576 + 3, 3000001])
def test_host_buffer_pipeline_at_ragged_sizes(ctx, n):
    """rrt_intersect / rrt_intersect_p over host buffers run a pipeline of tapered chunks (128 Ki ... 1 Mi ... 128 Ki rays,
    csrc/capi.cpp): whatever the batch length, every ray's answer is the device-resident call's."""
    import torch
    from rs_ray_toy_b200.aggregate import HIT_DTYPE, pack_rays
    p, idx = scenes.soup(30000)
    agg = scenes.gpu_soup(ctx, p, idx)
    rays = synth.bounce_rays(p, idx, n, seed=21)
    host = agg.intersect(rays)
    occ = agg.intersect_p(rays)
    d_rays = torch.from_numpy(pack_rays(rays).view(np.float64).copy()).cuda()
    d_hits = torch.zeros(n * 4, dtype=torch.float64, device="cuda")
    agg.intersect_device(n, d_rays.data_ptr(), d_hits.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    dev = d_hits.cpu().numpy().view(HIT_DTYPE).reshape(-1)
    assert np.array_equal(host, dev)
    assert np.array_equal(np.asarray(occ).astype(bool), dev["prim_id"] != 0xFFFFFFFF)


# ---- Tier L: the reference's own tree, order and accept rules, every quirk kept ------------------------
def _literal_pair(ctx, kind, n, **kw):
    import oracle_lib as O
    from rs_ray_toy_b200 import capi
    from rs_ray_toy_b200.aggregate import GpuAggregate
    if kind == "soup":
        p, idx = scenes.soup(n, edge=kw.get("edge", 0.02))
        ref = scenes.oracle_soup(p, idx, tier=O.TIER_L)
        a = GpuAggregate(ctx)
        m = a.add_mesh(p, idx)
        a.add_triangles(m, 0)
        rays = synth.bounce_rays(p, idx, kw.get("rays", 20000), seed=17)
    elif kind == "cubes":
        m_, inv = scenes.cube_instances(n, extent=kw.get("extent", 20.0))
        ref = scenes.oracle_cubes(m_, inv, tier=O.TIER_L)
        a = GpuAggregate(ctx)
        mesh = a.add_mesh(synth.CUBE_P, synth.CUBE_VI, synth.CUBE_N, synth.CUBE_NI)
        a.add_triangles(mesh, 0, instances=(m_, inv))
        rays = synth.camera_like_rays(kw.get("rays", 20000), (0.0, 0.0, -60.0), kw.get("extent", 20.0))
    else:
        m_, inv = scenes.sphere_instances(n, extent=kw.get("extent", 12.0))
        ref = scenes.oracle_spheres(m_, inv, 0.5, tier=O.TIER_L)
        a = GpuAggregate(ctx)
        a.add_sphere(radius=0.5, instances=(m_, inv))
        rays = synth.camera_like_rays(kw.get("rays", 20000), (0.0, 0.0, -40.0), kw.get("extent", 12.0))
    a.commit(4, capi.RRT_BUILD_LITERAL)
    return a, ref, rays


@pytest.mark.parametrize("kind,n", [("soup", 3000), ("cubes", 300), ("spheres", 1500)])
def test_literal_tier_is_bit_exact(ctx, kind, n):
    """RRT_BUILD_LITERAL: same HLBVH (Q1/Q2 included), same visiting order, "last accepted hit wins"
    (Q3), instance rays renormalised (Q6), intersect_p with E2 = p2 - p1 (Q4): the primitive index, t
    and the any-hit flags equal the oracle's literal tier bit for bit (f64, same operation order)."""
    agg, ref, rays = _literal_pair(ctx, kind, n)
    assert np.array_equal(agg.world_bound(), ref.world_bound())
    r = ref.intersect(rays)
    h = agg.intersect(rays)
    gp = h["prim_id"].astype(np.int64)
    gp[gp == 0xFFFFFFFF] = -1
    assert np.array_equal(gp, r["prim"])
    hit = r["prim"] >= 0
    assert hit.sum() > 500
    assert np.array_equal(h["t"][hit], r["t"][hit])
    occ_ref, _ = ref.intersect_p(rays)
    assert np.array_equal(agg.intersect_p(rays), occ_ref)
    sh = synth.shadow_rays_from(rays if kind == "soup" else np.concatenate([rays[:, :3] + rays[:, 3:6] * 30.0, rays[:, 3:]], axis=1),
                                (0.5, 0.5, 1.5) if kind == "soup" else (0.0, 30.0, 0.0))
    occ_ref, _ = ref.intersect_p(sh)
    assert np.array_equal(agg.intersect_p(sh), occ_ref)


def test_literal_and_fixed_tiers_differ_where_the_survey_says(ctx):
    """Q1 drops primitives inside treelets that split: on instanced cubes the literal aggregate misses
    geometry the fixed tier finds."""
    import oracle_lib as O
    agg_l, ref_l, rays = _literal_pair(ctx, "cubes", 300)
    m_, inv = scenes.cube_instances(300, extent=20.0)
    agg_f = scenes.gpu_cubes(ctx, m_, inv)
    hl, hf = agg_l.intersect(rays), agg_f.intersect(rays)
    assert (hl["prim_id"] != hf["prim_id"]).sum() > 0


@pytest.mark.parametrize("force", ["0", "1"])
def test_both_node_formats_agree_with_the_oracle(ctx, force, monkeypatch):
    """The fp32 Node64 and the 15-bit-grid Node32 (device_layout.h) are two encodings of the same conservative
    boxes: forced either way (RRT_QUANTISE is read at commit) the hits are the oracle's, on triangles, rotated
    instances and spheres."""
    monkeypatch.setenv("RRT_QUANTISE", force)
    p, idx = scenes.soup(60000)
    rays = synth.bounce_rays(p, idx, 120000, seed=77)
    ref = scenes.oracle_soup(p, idx).intersect(rays)
    agg = scenes.gpu_soup(ctx, p, idx)
    st = agg.stats()
    assert (st["device_bytes"] - 48 * st["n_records"]) // st["n_nodes"] == (32 if force == "1" else 64), st
    c = _assert_closest(agg.intersect(rays), ref["prim"], ref["t"], ref["uv"])
    assert c["t_exact"] == c["hits"], c
    sh = synth.shadow_rays_from(rays, (0.5, 0.5, 1.5))
    occ_ref, _ = scenes.oracle_soup(p, idx).intersect_p(sh)
    assert (agg.intersect_p(sh) == occ_ref).all()
    m, inv = scenes.cube_instances(300, extent=8.0)
    rng = np.random.default_rng(5)
    rays = np.concatenate([rng.uniform(-12, 12, (60000, 3)), synth.random_unit_vectors(60000, rng), np.full((60000, 1), np.inf)], axis=1)
    ref = scenes.oracle_cubes(m, inv).intersect(rays)
    _assert_closest(scenes.gpu_cubes(ctx, m, inv).intersect(rays), ref["prim"], ref["t"])
    m, inv = scenes.sphere_instances(2000, extent=10.0)
    ref = scenes.oracle_spheres(m, inv).intersect(rays)
    _assert_closest(scenes.gpu_spheres(ctx, m, inv).intersect(rays), ref["prim"], ref["t"])


def test_small_details_in_a_huge_box_keep_fp32_nodes(ctx):
    """A grid cell of extent / 32768 would swallow 1e-4-sized triangles scattered over a 1000-unit box: such a scene
    must stay on Node64 (chosen per scene at commit), and either way the answers are the oracle's."""
    p, idx = scenes.soup(20000, edge=1e-4)
    p = p * 1000.0
    p = p.astype(np.float32).astype(np.float64)
    rays = synth.bounce_rays(p, idx, 50000, seed=9)
    agg = scenes.gpu_soup(ctx, p, idx)
    st = agg.stats()
    assert (st["device_bytes"] - 48 * st["n_records"]) // st["n_nodes"] == 64, st
    ref = scenes.oracle_soup(p, idx).intersect(rays)
    _assert_closest(agg.intersect(rays), ref["prim"], ref["t"], ref["uv"])
    # config-3 shape: the grid is 0.3% of a leaf's edge -> quantised
    p, idx = scenes.soup(20000)
    st = scenes.gpu_soup(ctx, p, idx).stats()
    assert (st["device_bytes"] - 48 * st["n_records"]) // st["n_nodes"] == 32, st


@pytest.mark.parametrize("quantise", ["0", "1"])
def test_device_lbvh_build_gives_the_oracles_answers(ctx, quantise, monkeypatch):
    """RRT_BUILD_DEVICE_LBVH: the tree is built on the GPU (Morton keys, radix sort, binary radix tree).  Tier-F
    answers do not depend on the tree, so every check of the host-built aggregate must hold unchanged — triangles
    (fp32 records), rotated instances (f64 records), spheres, both node formats, leaf sizes 1 / 4 / 8."""
    from rs_ray_toy_b200 import capi
    monkeypatch.setenv("RRT_QUANTISE", quantise)
    p, idx = scenes.soup(200000)
    rays = synth.bounce_rays(p, idx, 200000, seed=43)
    oracle = scenes.oracle_soup(p, idx)
    ref = oracle.intersect(rays)
    sh = synth.shadow_rays_from(rays, (0.5, 0.5, 1.5))
    occ_ref, _ = oracle.intersect_p(sh)
    for max_prims in (4, 1, 8):
        agg = scenes.gpu_soup(ctx, p, idx, max_prims, capi.RRT_BUILD_DEVICE_LBVH)
        info = agg.build_info()
        assert info["device_lbvh"] and info["node_bytes"] == (32 if quantise == "1" else 64), info
        assert 0 < info["tree_device_usec"] < 2_000_000
        st = agg.stats()
        assert st["n_leaves"] == st["n_nodes"] + 1 and st["n_records"] == 200000
        c = _assert_closest(agg.intersect(rays), ref["prim"], ref["t"], ref["uv"])
        assert c["t_exact"] == c["hits"], c
        assert (agg.intersect_p(sh) == occ_ref).all()
    rng = np.random.default_rng(6)
    rays = np.concatenate([rng.uniform(-12, 12, (60000, 3)), synth.random_unit_vectors(60000, rng), np.full((60000, 1), np.inf)], axis=1)
    m, inv = scenes.cube_instances(300, extent=8.0)
    ref = scenes.oracle_cubes(m, inv).intersect(rays)
    _assert_closest(scenes.gpu_cubes(ctx, m, inv, build_flags=capi.RRT_BUILD_DEVICE_LBVH).intersect(rays), ref["prim"], ref["t"])
    m, inv = scenes.sphere_instances(2000, extent=10.0)
    ref = scenes.oracle_spheres(m, inv).intersect(rays)
    _assert_closest(scenes.gpu_spheres(ctx, m, inv, build_flags=capi.RRT_BUILD_DEVICE_LBVH).intersect(rays), ref["prim"], ref["t"])


def test_device_lbvh_small_scenes_fall_back_to_the_host_builder(ctx):
    from rs_ray_toy_b200 import capi
    p, idx = scenes.soup(12)
    rays = synth.bounce_rays(p, idx, 2000, seed=3)
    ref = scenes.oracle_soup(p, idx).intersect(rays)
    agg = scenes.gpu_soup(ctx, p, idx, 4, capi.RRT_BUILD_DEVICE_LBVH)
    assert agg.build_info()["tree_device_usec"] == 0
    _assert_closest(agg.intersect(rays), ref["prim"], ref["t"])


def test_device_lbvh_duplicates_and_a_tree_too_deep_for_the_stack(ctx):
    """Equal centroids are split by position on the device too; and a scene whose Morton keys are 1, 2, 4, ... 2^62
    plus thousands of copies at key 0 makes a radix tree deeper than the 64-entry traversal stack — the commit
    then falls back to the SAH builder instead of failing.  Answers are the oracle's in both cases."""
    from rs_ray_toy_b200 import capi
    rng = np.random.default_rng(12)
    tri = np.array([[0.0, 0.0, 0.0], [1e-3, 0.0, 0.0], [0.0, 1e-3, 0.0]])
    # 6000 identical triangles + 2000 random ones
    p_rand, idx_rand = scenes.soup(2000)
    p = np.concatenate([np.tile(tri + 0.25, (6000, 1)), p_rand])
    idx = np.arange(len(p), dtype=np.uint32).reshape(-1, 3)
    rays = synth.bounce_rays(p_rand, idx_rand, 20000, seed=2)
    rays[:5000, 0:3] = (0.2503, 0.2503, 1.0)
    rays[:5000, 3:6] = (0.0, 0.0, -1.0)
    ref = scenes.oracle_soup(p, idx).intersect(rays)
    agg = scenes.gpu_soup(ctx, p, idx, 4, capi.RRT_BUILD_DEVICE_LBVH)
    assert agg.build_info()["tree_device_usec"] > 0
    c = _assert_closest(agg.intersect(rays), ref["prim"], ref["t"])
    assert c["hits"] >= 5000
    # one tiny triangle per Morton bit + 5000 copies at the origin corner
    pts = [np.zeros(3)]
    for j in range(21):
        for axis in range(3):
            q = np.zeros(3)
            q[axis] = 2.0 ** j / 2.0 ** 21
            pts.append(q)
    pts = np.array(pts) * 0.5
    small = tri * 1e-9
    p = np.concatenate([np.tile(small, (5000, 1))] + [small + q for q in pts])
    idx = np.arange(len(p), dtype=np.uint32).reshape(-1, 3)
    o = np.concatenate([rng.uniform(-0.1, 0.6, (4000, 2)), np.full((4000, 1), 1.0)], axis=1)
    rays = np.concatenate([o, np.tile([0.0, 0.0, -1.0], (4000, 1)), np.full((4000, 1), np.inf)], axis=1)
    rays[:500, 0:2] = 2e-13
    ref = scenes.oracle_soup(p, idx).intersect(rays)
    agg = scenes.gpu_soup(ctx, p, idx, 4, capi.RRT_BUILD_DEVICE_LBVH)
    _assert_closest(agg.intersect(rays), ref["prim"], ref["t"])
    assert agg.stats()["max_depth"] + 2 <= 64


def _general_sphere_scene(ctx, tier_scene):
    """Spheres that need Sphere::intersect in full (sphere.rs:124-259): clipped in z and in phi, under their own
    translated / rotated / non-uniformly scaled object transform, bare and instanced (instances with scale too),
    next to ordinary full spheres.  Built identically on both sides; prim ids follow insertion order."""
    from rs_ray_toy_b200 import transform as T
    from rs_ray_toy_b200.aggregate import GpuAggregate
    rng = np.random.default_rng(21)
    shapes = []
    for k in range(40):
        radius = float(rng.uniform(0.4, 1.2))
        kind = k % 5
        z_min, z_max, phi_max = -radius, radius, 360.0
        if kind in (0, 1, 4):
            z_min, z_max = sorted((float(rng.uniform(-radius, 0.2 * radius)), float(rng.uniform(0.3 * radius, radius))))
        if kind in (1, 2, 4):
            phi_max = float(rng.uniform(60.0, 330.0))
        scale = (1.0, 1.0, 1.0) if kind in (0, 2) else tuple(rng.uniform(0.6, 1.8, 3))
        clipped = kind != 3
        # a clipped sphere is placed through instances (Q5a: see test_clipped_sphere_with_own_transform_is_refused)
        o2w = T.make_to_world(rng.uniform(-6, 6, 3), rng.normal(0, 1, 3), float(rng.uniform(0, 360)), scale) if not clipped \
            else T.make_to_world()
        inst = None
        if k % 2 == 0 or clipped:
            ms, invs = [], []
            for _ in range(3):
                m, inv = T.make_to_world(rng.uniform(-8, 8, 3), rng.normal(0, 1, 3), float(rng.uniform(0, 360)),
                                         tuple(rng.uniform(0.6, 1.6, 3)) if kind != 0 else (1.0, 1.0, 1.0))
                ms.append(m)
                invs.append(inv)
            inst = (np.array(ms), np.array(invs))
        shapes.append((radius, z_min, z_max, phi_max, o2w, inst))
    a = GpuAggregate(ctx)
    s = O.OracleScene(tier_scene)
    for (radius, z_min, z_max, phi_max, o2w, inst) in shapes:
        a.add_sphere(radius=radius, z_min=z_min, z_max=z_max, phi_max=phi_max, obj_to_world=o2w, instances=inst)
        g = s.add_geo_sphere(s.add_sphere(o2w[0], o2w[1], radius, z_min, z_max, phi_max))
        if inst is None:
            s.add_prims(g, 1, -1)
        else:
            for i in range(inst[0].shape[0]):
                s.add_prims(g, 1, s.add_xform(inst[0][i], inst[1][i]))
    # ordinary full spheres through rigid instances keep the world-space fast path in the same tree
    m, inv = scenes.sphere_instances(200, extent=9.0)
    a.add_sphere(radius=0.5, instances=(m, inv))
    g = s.add_geo_sphere(s.add_sphere(None, None, 0.5))
    for i in range(m.shape[0]):
        s.add_prims(g, 1, s.add_xform(m[i], inv[i]))
    return a, s


@pytest.mark.parametrize("flags", [0, "lbvh"])
def test_partial_and_scaled_spheres(ctx, flags):
    """SURVEY §8a6 in full: z / phi clipping with the retry at the far root, any affine object and instance transform
    (Tier F: one parameter t through all spaces), Q5a kept — closest hit, (u, v) and any-hit against the oracle."""
    build = capi.RRT_BUILD_DEVICE_LBVH if flags == "lbvh" else capi.RRT_BUILD_FAST
    a, s = _general_sphere_scene(ctx, O.TIER_F)
    a.commit(4, build)
    s.build(4)
    rng = np.random.default_rng(8)
    n = 200000
    o = rng.uniform(-14, 14, (n, 3))
    tgt = rng.uniform(-9, 9, (n, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1)[:, None]
    rays = np.concatenate([o, d, np.full((n, 1), np.inf)], axis=1)
    rays[: n // 4, 6] = rng.uniform(2.0, 20.0, n // 4)       # finite t_max too
    ref = s.intersect(rays)
    hits = a.intersect(rays)
    c = _assert_closest(hits, ref["prim"], ref["t"], ref["uv"])
    assert c["hits"] > n // 10, c
    # the clipped shapes are really clipped: a fair share of rays passes through a sphere's removed part
    occ_ref, _ = s.intersect_p(rays)
    assert (a.intersect_p(rays) == occ_ref).all()
    # rays leaving the surfaces (inside and outside: the far root and the self-hit floor, Q8)
    hit = ref["prim"] >= 0
    p = rays[hit, 0:3] + rays[hit, 3:6] * ref["t"][hit, None]
    d2 = synth.random_unit_vectors(p.shape[0], rng)
    sec = np.concatenate([p, d2, np.full((p.shape[0], 1), np.inf)], axis=1)
    r2 = s.intersect(sec)
    _assert_closest(a.intersect(sec), r2["prim"], r2["t"], r2["uv"])
    occ2, _ = s.intersect_p(sec)
    assert (a.intersect_p(sec) == occ2).all()


def test_clipped_sphere_with_own_transform_is_refused(ctx):
    """The reference clips a sphere's first root by the z / phi of the point on the ray it was HANDED (Q5a, sphere.rs:157).
    If the sphere has an object transform of its own, that accepts points outside the shape's bound, found only when the
    ray happens to cross the BVH leaf's box: the reference's own answer depends on the tree.  Such a sphere gets its
    placement from instances[] instead; the same sphere with the same placement as an instance commits."""
    from rs_ray_toy_b200 import transform as T
    from rs_ray_toy_b200.aggregate import GpuAggregate
    xf = T.make_to_world((1.0, 2.0, 3.0), (0.0, 1.0, 0.0), 30.0)
    a = GpuAggregate(ctx)
    a.add_sphere(radius=1.0, z_min=-0.3, z_max=0.8, obj_to_world=xf)
    with pytest.raises(capi.RrtError) as e:
        a.commit(4)
    assert e.value.status == capi.RRT_ERR_UNSUPPORTED
    b = GpuAggregate(ctx)
    b.add_sphere(radius=1.0, z_min=-0.3, z_max=0.8, instances=(xf[0][None], xf[1][None]))
    b.commit(4)
    assert b.num_prims == 1


def test_instance_update_answers_like_a_fresh_scene(ctx):
    """rrt_scene_update_instances (dynamic scenes): instanced cubes and spheres are moved, the tree is
    made anew on the device, and every ray is answered like the oracle's scene built over the new transforms; an
    integrator made over the old scene blocks the update."""
    from rs_ray_toy_b200 import capi
    rng = np.random.default_rng(16)
    rays = np.concatenate([rng.uniform(-12, 12, (80000, 3)), synth.random_unit_vectors(80000, rng), np.full((80000, 1), np.inf)], axis=1)
    m0, inv0 = scenes.cube_instances(300, extent=8.0)
    agg = scenes.gpu_cubes(ctx, m0, inv0)
    before = agg.intersect(rays)
    m1, inv1 = scenes.cube_instances(300, extent=8.0, seed=991)
    k = 120                                   # the first 120 cubes move, the rest stay
    agg.update_instances(0, m1[:k], inv1[:k])
    m_new, inv_new = np.concatenate([m1[:k], m0[k:]]), np.concatenate([inv1[:k], inv0[k:]])
    ref = scenes.oracle_cubes(m_new, inv_new).intersect(rays)
    after = agg.intersect(rays)
    _assert_closest(after, ref["prim"], ref["t"])
    assert (after["prim_id"] != before["prim_id"]).mean() > 0.01
    assert agg.build_info()["device_lbvh"]
    with pytest.raises(capi.RrtError):        # out of range
        agg.update_instances(290, m1[:20], inv1[:20])
    ms, invs = scenes.sphere_instances(2000, extent=10.0)
    sph = scenes.gpu_spheres(ctx, ms, invs)
    ms2, invs2 = scenes.sphere_instances(2000, extent=10.0, seed=77)
    sph.update_instances(0, ms2, invs2)
    ref = scenes.oracle_spheres(ms2, invs2).intersect(rays)
    _assert_closest(sph.intersect(rays), ref["prim"], ref["t"])
