"""Scene builders shared by the parity tests: the SAME arrays go to the CPU oracle and to
the CUDA library (through its C ABI), so any disagreement is arithmetic, not input."""
import numpy as np

import oracle_lib as O
from rs_ray_toy_b200 import synth, transform


def soup(n_tris, edge=0.01, seed=synth.SEED_C3_SOUP):
    return synth.soup_triangles(n_tris, edge, seed)


def oracle_soup(p, idx, tier=O.TIER_F, max_prims=4):
    return O.soup_scene(p, idx, tier, max_prims)


def gpu_soup(ctx, p, idx, max_prims=4, build_flags=0):
    from rs_ray_toy_b200.aggregate import soup_aggregate
    return soup_aggregate(ctx, p, idx, max_prims, build_flags)


def cube_instances(n, extent=50.0, seed=synth.SEED_C2_INSTANCES):
    prm = synth.instance_params(n, extent, seed)
    return transform.make_to_world_batch(prm["world_pos"], prm["axis"], prm["angle"])


def oracle_cubes(m, inv, tier=O.TIER_F, max_prims=4):
    """config 2: samples/cube.obj instanced n times (renderprocess.rs:1228-1282)."""
    s = O.OracleScene(tier)
    mesh = s.add_mesh(synth.CUBE_P, synth.CUBE_VI, synth.CUBE_N, synth.CUBE_NI)
    g0 = s.add_geo_triangles(mesh)
    for i in range(m.shape[0]):
        s.add_prims(g0, 12, s.add_xform(m[i], inv[i]))
    s.build(max_prims)
    return s


def gpu_cubes(ctx, m, inv, max_prims=4, build_flags=0):
    from rs_ray_toy_b200.aggregate import GpuAggregate
    a = GpuAggregate(ctx)
    mesh = a.add_mesh(synth.CUBE_P, synth.CUBE_VI, synth.CUBE_N, synth.CUBE_NI)
    a.add_triangles(mesh, 0, instances=(m, inv))
    return a.commit(max_prims, build_flags)


def sphere_instances(n, extent=50.0, seed=synth.SEED_C4_SPHERES):
    prm = synth.instance_params(n, extent, seed, rotate=False)
    return transform.make_to_world_batch(prm["world_pos"], prm["axis"], prm["angle"])


def oracle_spheres(m, inv, radius=0.5, tier=O.TIER_F, max_prims=4):
    """config 4: an identity sphere placed through instances[] (Q5a)."""
    s = O.OracleScene(tier)
    sp = s.add_sphere(None, None, radius)
    g = s.add_geo_sphere(sp)
    for i in range(m.shape[0]):
        s.add_prims(g, 1, s.add_xform(m[i], inv[i]))
    s.build(max_prims)
    return s


def gpu_spheres(ctx, m, inv, radius=0.5, max_prims=4, build_flags=0):
    from rs_ray_toy_b200.aggregate import GpuAggregate
    a = GpuAggregate(ctx)
    a.add_sphere(radius=radius, instances=(m, inv))
    return a.commit(max_prims, build_flags)


def compare_closest(hits, ref_prim, ref_t, rel_tie=1e-6, rel_t=1e-5):
    """north_star's bar: primitive index bit-exact except rays whose two best candidates tie
    within 1e-6 relative (none of those appear unless t differs); t within 1e-5 relative.
    Returns a dict of counts; asserts nothing."""
    gp = hits["prim_id"].astype(np.int64)
    gp[gp == 0xFFFFFFFF] = -1
    same = gp == ref_prim
    hit = ref_prim >= 0
    both = hit & (gp >= 0)
    dt = np.zeros(len(gp))
    dt[both] = np.abs(hits["t"][both] - ref_t[both]) / np.maximum(np.abs(ref_t[both]), 1e-300)
    ties = (~same) & both & (dt <= rel_tie)
    bad = np.nonzero((~same) & ~ties)[0][:5]
    return {
        "examples": [(int(i), int(gp[i]), int(ref_prim[i]), float(hits["t"][i]), float(ref_t[i])) for i in bad],
        "n": len(gp),
        "mismatch": int((~same).sum()),
        "mismatch_excl_ties": int(((~same) & ~ties).sum()),
        "t_bad": int((both & same & (dt > rel_t)).sum()),
        "t_exact": int((both & same & (hits["t"] == ref_t)).sum()),
        "hits": int(hit.sum()),
        "max_rel_dt": float(dt[both & same].max()) if (both & same).any() else 0.0,
    }


def oracle_c5(n_tris, edge, xres, yres, nsamp, seed_render=1, max_depth=5, nthreads=None, crop=None, want_dump=False,
              textured=False):
    """The oracle-side twin of synth.scene_c5_api: same soup, materials, lights and camera, fed
    through the oracle's C API (no scene.json for million-triangle meshes)."""
    import ctypes as C
    import oracle_scene as S
    p, idx = synth.soup_triangles(n_tris, edge, synth.SEED_C5_SOUP)
    half = n_tris // 2
    s = O.OracleScene(O.TIER_F)
    m0 = s.add_mesh(p[: 3 * half], idx[:half])
    m1 = s.add_mesh(p[3 * half:], idx[half:] - 3 * half)
    g0 = s.add_geo_triangles(m0, 0)
    s.add_prims(g0, half, -1)
    g1 = s.add_geo_triangles(m1, 1)
    s.add_prims(g1, n_tris - half, -1)
    s.build(4)
    mats = np.zeros((2, S.MAT_ROW))
    mats[:, 26:38] = -1
    mats[0, 0] = 0
    mats[0, 1:4] = (0.6, 0.55, 0.5)
    mats[1, 0] = 1
    mats[1, 1:4] = (0.3, 0.4, 0.6)
    mats[1, 4:7] = (0.3, 0.3, 0.3)
    mats[1, 20] = 0.15
    mats[:, 21:23] = -1.0
    lights = np.zeros((2, S.LIGHT_ROW))
    lights[0, 0] = 0
    lights[0, 1:4] = (4.0, 4.0, 4.0)
    lights[0, 7:23] = np.eye(4).reshape(16)
    lights[1, 0] = 1
    lights[1, 1:4] = (2.0, 2.0, 2.0)
    lights[1, 4:7] = np.subtract((0.3, 1.0, -0.6), (0, 0, 0))
    lights[1, 7:23] = np.eye(4).reshape(16)
    L = O.lib()
    L.orc_set_materials.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.orc_set_lights.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    if textured:
        rows = np.zeros((4, S.TEX_ROW))
        for r, (kind, vals, mp, m8, t1, t2, w2t) in zip(rows, synth.c5_texture_rows()):
            r[0], r[2], r[4], r[5], r[6] = kind, mp, t1, t2, -1
            for k, v in enumerate(vals):
                r[8 + 3 * k: 11 + 3 * k] = [v, 0.0, 0.0] if np.isscalar(v) else v
            r[20:28] = m8
            r[28:44] = np.asarray(w2t, dtype=np.float64).reshape(16)
        mats[0, 26 + S.T_KD] = 2
        mats[1, 26 + S.T_ROUGH] = 3
        L.orc_set_textures.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_set_textures(s.h, 4, rows.ctypes.data)
    L.orc_set_materials(s.h, 2, mats.ctypes.data)
    L.orc_set_lights(s.h, 2, lights.ctypes.data)
    prm = np.zeros(48)
    prm[0], prm[1], prm[2] = xres, yres, 35.0
    prm[3], prm[4], prm[5], prm[6], prm[7], prm[8] = 0, 0.5, 0.5, 2.0, 1.0, np.inf
    prm[9:12] = (0.5, 0.5, -2.5)
    prm[12:15] = (0.5, 0.5, 0.5)
    prm[15:18] = (0.0, 1.0, 0.0)
    prm[18], prm[19], prm[20], prm[21], prm[22] = 0.0, 1.0, 50.0, 3.0, 1.0
    prm[23], prm[24], prm[25] = nsamp, 0, seed_render
    prm[26], prm[27], prm[28] = 0, max_depth, 1.0
    prm[29], prm[30] = 1, 0
    if crop is not None:
        prm[31] = 1.0
        prm[32:36] = crop
    prm[36] = 1.0 if want_dump else 0.0
    return S.oracle_render(s, prm, np.array(synth.DGAUSS_LENS, dtype=np.float64), nthreads, want_dump)
