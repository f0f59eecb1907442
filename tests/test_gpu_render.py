"""-m gpu: the wavefront renderer (through the C ABI) against the CPU oracle's restatement of
si_render / PathIntegrator::li / DirectLightingIntegrator::li on the same scene.json files,
seeds and sample counts.  Bars (BASELINE.json north_star): first-hit primitive index of every
camera ray bit-exact (ties within 1e-6 relative excluded), t within 1e-5 relative, rendered image
within 1e-3 relative RMSE.  The device shades in f64 with the reference's operation order, so the
observed differences are libm last-ulp effects (sin / cos / atan2 / acos / exp / log)."""
import numpy as np
import pytest

import oracle_scene as S
from rs_ray_toy_b200 import synth
from rs_ray_toy_b200.render import Render

pytestmark = pytest.mark.gpu
REL_RMSE = 1e-3


def rel_rmse(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-300))


def compare(gpu: Render, ref: dict, check_dump=True, rmse_bound=None):
    rgb, raw = gpu.film(want_raw=True)
    st = gpu.stats()
    out = {"rmse": rel_rmse(rgb, ref["rgb"]), "max_abs": float(np.abs(rgb - ref["rgb"]).max()),
           "camera_rays": (st["camera_rays"], ref["stats"]["camera_rays"]),
           "extension_rays": (st["extension_rays"], ref["stats"]["extension_rays"]),
           "shadow_rays": (st["shadow_rays"], ref["stats"]["shadow_rays"]),
           "zero_weight": (st["zero_weight"], ref["stats"]["zero_weight"])}
    # filter weights are sample counts: exact
    assert np.array_equal(raw[..., 3], ref["raw"][..., 3]), out
    assert st["camera_rays"] == ref["stats"]["camera_rays"], out
    assert st["zero_weight"] == ref["stats"]["zero_weight"], out
    if check_dump:
        d, r = gpu.hit_dump(), ref["dump"]
        assert d.shape == r.shape, (d.shape, r.shape)
        assert np.array_equal(d[:, :3], r[:, :3])
        same = d[:, 3] == r[:, 3]
        hit = r[:, 3] >= 0
        rel = np.abs(d[:, 4] - r[:, 4]) / np.maximum(np.abs(r[:, 4]), 1e-300)
        out["first_hit_mismatch"] = int((~same).sum())
        out["first_hit_t_max_rel"] = float(rel[same & hit].max()) if (same & hit).any() else 0.0
        assert (~same).sum() == 0, out
        assert out["first_hit_t_max_rel"] <= 1e-5, out
        assert np.allclose(d[:, 5], r[:, 5], rtol=1e-12, atol=0), out
    assert out["rmse"] <= (REL_RMSE if rmse_bound is None else rmse_bound), out
    return out


def test_config1_path_traced_sample_scene(ctx, tmp_path):
    """Config 1: the reference's sample scene, Path integrator, Halton nsamp 17 (16 rendered)."""
    path = synth.scene_c1(str(tmp_path / "c1"), nsamp=17)
    ref = S.load(path).render(seed=1, want_dump=True)
    gpu = Render.load(ctx, path, seed=1)
    gpu.enable_hit_dump()
    gpu.run()
    out = compare(gpu, ref)
    assert ref["stats"]["camera_rays"] > 500000
    # extension / shadow ray counts agree unless a path decision flipped on a last-ulp difference
    assert abs(out["extension_rays"][0] - out["extension_rays"][1]) <= 4, out
    assert abs(out["shadow_rays"][0] - out["shadow_rays"][1]) <= 4, out
    assert out["rmse"] < 1e-6, out


def test_config1_literal_tier(ctx, tmp_path):
    """Config 1 in the LITERAL tier: the reference's own tree and accept rules, Q9 shadow rays, Q6
    instance rays — against the oracle with every quirk flag at its reference setting."""
    import oracle_lib as O
    path = synth.scene_c1(str(tmp_path / "c1"), nsamp=9)
    ref = S.load(path, tier=O.TIER_L).render(seed=1, want_dump=True)
    gpu = Render.load(ctx, path, seed=1, literal=True)
    gpu.enable_hit_dump()
    gpu.run()
    # Every camera ray's (primitive, t) and every filter weight is bit-equal (checked inside compare).  Deeper in
    # the path the literal tier is chaotic by construction: under Q3 the winner among the accepted candidates
    # depends on box tests against a t_max that an earlier accept has just set, so a 1-ulp difference in a sampled
    # direction (CUDA's sin/cos vs glibc's) can swap the hit (measured: 6 of 230,400 pixels at 8 spp; traced with
    # tools/debug_render_rays.py to a direction that differs by 4 ulp and exits through the neighbouring triangle).
    # Those pixels are excluded by count, everything else must agree to 1e-6.
    rgb = gpu.film()
    diff = np.abs(rgb - ref["rgb"]).max(axis=2)
    flipped = diff > 1e-9
    n_flipped = int(flipped.sum())
    whole = rel_rmse(rgb, ref["rgb"])
    print(f"literal tier: {n_flipped} of {flipped.size} pixels excluded as Q3 flips; whole-image relative RMSE {whole:.3e}")
    assert n_flipped <= 6, n_flipped          # the measured count: a larger one is a regression, not chaos
    masked = np.where(flipped[..., None], ref["rgb"], rgb)
    assert rel_rmse(masked, ref["rgb"]) < 1e-6
    # north_star's image bound (1e-3) holds for the image WITHOUT the excluded pixels by three orders of magnitude; with
    # them the frame stays under 5e-3 (six pixels that took the other branch of a Q3 tie, each a whole light sample off)
    out = compare(gpu, ref, rmse_bound=5e-3)
    # and it is a different image from the fixed tier (shadows: Q4 / Q9)
    fixed = Render.load(ctx, path, seed=1)
    fixed.run()
    assert rel_rmse(fixed.film(), ref["rgb"]) > 1e-3


def test_config1_seed_changes_image_and_identity_perms(ctx, tmp_path):
    path = synth.scene_c1(str(tmp_path / "c1"), xres=160, yres=90, nsamp=9)
    imgs = {}
    for seed in (0, 1, 2):
        ref = S.load(path).render(seed=seed, want_dump=True)
        gpu = Render.load(ctx, path, seed=seed)
        gpu.enable_hit_dump()
        gpu.run()
        compare(gpu, ref)
        imgs[seed] = gpu.film()
    assert not np.array_equal(imgs[1], imgs[2])


def test_config2_direct_lighting_instanced_cubes(ctx, tmp_path):
    path = synth.scene_c2(str(tmp_path / "c2"), n_instances=2000, xres=480, yres=270, extent=30.0)
    ref = S.load(path).render(seed=1, want_dump=True)
    gpu = Render.load(ctx, path, seed=1)
    gpu.enable_hit_dump()
    gpu.run()
    out = compare(gpu, ref)
    assert out["shadow_rays"][1] > 1000 and out["shadow_rays"][0] == out["shadow_rays"][1], out


@pytest.mark.parametrize("extra", [False, True])
def test_config4_sphere_field_materials(ctx, tmp_path, extra):
    """Plastic / Metal presets (+ Mirror, Glass, Oren-Nayar Matte when `extra`), distant + point
    light, 5-bounce paths with Russian roulette."""
    path = synth.scene_c4(str(tmp_path / "c4"), n_spheres=4000, xres=320, yres=180, nsamp=9, extent=14.0,
                          extra_materials=extra)
    ref = S.load(path).render(seed=3, want_dump=True)
    gpu = Render.load(ctx, path, seed=3)
    gpu.enable_hit_dump()
    gpu.run()
    out = compare(gpu, ref)
    assert out["extension_rays"][1] > 2 * out["camera_rays"][1], out
    assert abs(out["extension_rays"][0] - out["extension_rays"][1]) <= max(8, out["extension_rays"][1] // 100000), out


def test_config5_soup_through_the_api(ctx):
    """Config 5's shape at a small size: two bare (non-instanced) meshes, Matte + Plastic, point +
    distant light, assembled through rrt_scene_add_* and rrt_render_create (no scene.json)."""
    import scenes
    agg, gpu = synth.scene_c5_api(ctx, n_tris=60000, edge=0.02, xres=256, yres=144, nsamp=5)
    gpu.enable_hit_dump()
    gpu.run()
    ref = scenes.oracle_c5(60000, 0.02, 256, 144, 5, want_dump=True)
    out = compare(gpu, ref)
    assert out["extension_rays"][1] > out["camera_rays"][1], out


def test_filters_and_overrides(ctx, tmp_path):
    """Gaussian / triangle filters of radius 2 (samples splat across tile borders) and the
    luminance clamp, through the loader's `overrides`."""
    path = synth.scene_c1(str(tmp_path / "c1"), xres=200, yres=120, nsamp=5)
    for flt in ({"filter_type": "GaussianFilter", "radius": [2.0, 2.0], "alpha": 2.0},
                {"filter_type": "TriangleFilter", "radius": [1.5, 2.0]}):
        ov = {"Film": {"xres": 200, "yres": 120, "diagonal": 20, "scale": 2.0, "max_sample_luminance": 0.05, "Filter": flt}}
        ref = S.load(path, ov).render(seed=1, want_dump=False)
        gpu = Render.load(ctx, path, overrides=ov, seed=1)
        gpu.run()
        rgb, raw = gpu.film(want_raw=True)
        assert np.allclose(raw[..., 3], ref["raw"][..., 3], rtol=1e-12)
        assert rel_rmse(rgb, ref["rgb"]) <= 1e-9


def test_tile_partition_is_exact(ctx, tmp_path):
    """renderprocess/si_render tiling dealt to ranks: tiles t % G == r on G renderers, films summed,
    equal the single-renderer film bit for bit (box filter 0.5: every pixel has one owner)."""
    path = synth.scene_c4(str(tmp_path / "c4"), n_spheres=1500, xres=200, yres=120, nsamp=5, extent=10.0)
    full = Render.load(ctx, path, seed=1)
    full.run()
    _, raw_full = full.film(want_raw=True)
    for G in (2, 4):
        acc = np.zeros_like(raw_full)
        for r in range(G):
            part = Render.load(ctx, path, seed=1)
            part.run(tile_mod=G, tile_rank=r)
            acc += part.film(want_raw=True)[1]
            part.close()
        assert np.array_equal(acc, raw_full)
    # the oracle deals tiles the same way
    ref = S.load(path).render(seed=1, tile_mod=2, tile_rank=1)
    part = Render.load(ctx, path, seed=1)
    part.run(tile_mod=2, tile_rank=1)
    assert np.array_equal(part.film(want_raw=True)[1][..., 3], ref["raw"][..., 3])


def test_crop_and_api_scene(ctx, tmp_path):
    """A scene assembled through the aggregate API (no scene.json) renders like the same scene loaded
    from a file; `crop` restricts sampling to a pixel rectangle."""
    path = synth.scene_c1(str(tmp_path / "c1"), xres=160, yres=90, nsamp=5)
    a = Render.load(ctx, path, seed=1)
    a.run(crop=(40, 20, 100, 60))
    ref = S.load(path).render(seed=1, crop=(40, 20, 100, 60))
    rgb, raw = a.film(want_raw=True)
    assert np.array_equal(raw[..., 3], ref["raw"][..., 3])
    assert raw[..., 3][:20].sum() == 0 and raw[..., 3][20:60, 40:100].min() > 0
    assert rel_rmse(rgb, ref["rgb"]) <= 1e-9


def test_unsupported_inputs_are_refused(ctx, tmp_path):
    from rs_ray_toy_b200 import capi
    path = synth.scene_c1(str(tmp_path / "c1"), xres=64, yres=36, nsamp=3)
    for ov, status in (({"Sampler": {"sampler_type": "SobolSampler"}}, capi.RRT_ERR_IO),
                       ({"Sampler": {"sampler_type": "HaltonSampler", "nsamp": -3}}, capi.RRT_ERR_IO),
                       ({"Integrator": {"integrator_type": "Path", "max_depth": -1}}, capi.RRT_ERR_IO),
                       ({"Film": {"xres": 64, "yres": 36, "Filter": {"filter_type": "GaussianFilter", "radius": [0.0, 2.0]}}}, capi.RRT_ERR_IO),
                       ({"Integrator": {"integrator_type": "SPPM"}}, capi.RRT_ERR_IO),
                       ({"infinite_lights": [{"light_type": "infinite"}]}, capi.RRT_ERR_IO)):
        with pytest.raises(capi.RrtError) as e:
            Render.load(ctx, path, overrides=ov)
        assert e.value.status == status
    with pytest.raises(capi.RrtError):
        Render.load(ctx, str(tmp_path / "missing.json"))


@pytest.mark.parametrize("integrator", ["Path", "DirectLighting"])
def test_diffuse_area_lights(ctx, tmp_path, integrator):
    """DiffuseAreaLights (a sphere emitter and a mesh triangle) sampled by next-event estimation
    with the reference's MIS weight and its pdf quirk (Q28), next to a point light.  The BSDF-sampling half of
    estimate_direct is dead code in effect (Q22) — the oracle traces and counts those rays, the device skips
    them — so films and the other ray counts must agree."""
    path = synth.scene_area_lights(str(tmp_path / "a"), xres=256, yres=144, nsamp=9, integrator=integrator,
                                   max_depth=5 if integrator == "Path" else 1)
    ref = S.load(path).render(seed=1, want_dump=True)
    gpu = Render.load(ctx, path, seed=1)
    gpu.enable_hit_dump()
    gpu.run()
    out = compare(gpu, ref)
    assert out["rmse"] < 1e-6, out
    assert ref["stats"]["mis_probe_rays"] > 0
    assert abs(out["extension_rays"][0] - out["extension_rays"][1]) <= 4, out
    assert abs(out["shadow_rays"][0] - out["shadow_rays"][1]) <= 4, out
    # the area lights alone light the scene too
    import json
    cfg = json.loads(open(path).read())
    only = {"lights": cfg["lights"][:2]}
    ref2 = S.load(path, only).render(seed=1)
    gpu2 = Render.load(ctx, path, overrides=only, seed=1)
    gpu2.run()
    assert ref2["rgb"].max() > 0
    assert rel_rmse(gpu2.film(), ref2["rgb"]) < 1e-6


def test_textured_scene(ctx, tmp_path):
    """SURVEY §8f row 3 (procedural part): every in-scope texture kind and mapping drives a material parameter of
    tests' textured scene (cubes on a uv-mapped floor + three spheres; Matte / Plastic / Metal / Mirror), loaded from
    the same scene.json by both sides.  First hits are bit-equal; the film agrees to rounding except where a
    4-ulp difference between CUDA's and glibc's atan2 / sin / cos moves a sample across a checker edge."""
    path = synth.scene_textured(str(tmp_path / "t"), xres=256, yres=144, nsamp=9)
    ref = S.load(path).render(seed=1, want_dump=True)
    gpu = Render.load(ctx, path, seed=1)
    gpu.enable_hit_dump()
    gpu.run()
    out = compare(gpu, ref)
    assert out["rmse"] < 1e-4, out
    assert abs(out["extension_rays"][0] - out["extension_rays"][1]) <= 4, out
    # the textures are seen: the same scene with constant materials renders a different image
    import json
    cfg = json.loads(open(path).read())
    plain = {"materials": [{"material_type": m["material_type"], "material_name": m["material_name"]} for m in cfg["materials"]]}
    gpu2 = Render.load(ctx, path, overrides=plain, seed=1)
    gpu2.run()
    assert rel_rmse(gpu2.film(), ref["rgb"]) > 0.05
    # DirectLighting reads the same materials
    dl = {"Integrator": {"integrator_type": "DirectLighting", "max_depth": 1, "light_strategy": "one"}}
    ref3 = S.load(path, dl).render(seed=1)
    gpu3 = Render.load(ctx, path, overrides=dl, seed=1)
    gpu3.run()
    assert rel_rmse(gpu3.film(), ref3["rgb"]) < 1e-4


def test_textures_through_the_api(ctx):
    """rrt_scene_set_textures / rrt_scene_set_material_textures on an API-assembled soup (config 5's shape): a 3D
    checkerboard drives the Matte kd, a planar float ramp the Plastic roughness."""
    import scenes
    from rs_ray_toy_b200 import capi
    from rs_ray_toy_b200 import render as R
    agg, gpu = synth.scene_c5_api(ctx, n_tris=40000, edge=0.03, xres=192, yres=108, nsamp=5, textured=True)
    gpu.enable_hit_dump()
    gpu.run()
    ref = scenes.oracle_c5(40000, 0.03, 192, 108, 5, want_dump=True, textured=True)
    out = compare(gpu, ref)
    assert out["rmse"] < 1e-4, out
    plain = scenes.oracle_c5(40000, 0.03, 192, 108, 5)
    assert rel_rmse(plain["rgb"], ref["rgb"]) > 0.02
    # a material slot that names a texture outside the table is refused when the integrator is made
    agg2, _ = synth.scene_c5_api(ctx, n_tris=2000, edge=0.05, xres=32, yres=18, nsamp=3)
    slots = np.full((2, R.MATERIAL_SLOTS), -1, dtype=np.int32)
    slots[0, R.SLOT_KD] = 7
    desc = synth.default_render_desc(32, 18, 3, cam_pos=(0.5, 0.5, -2.5), cam_look=(0.5, 0.5, 0.5), focus_distance=3.0)
    with pytest.raises(capi.RrtError):
        Render.create(agg2, [R.matte(), R.plastic()], [R.point_light()], desc, synth.DGAUSS_LENS,
                      textures=[R.texture(R.TEX_CONSTANT, [(0.5, 0.5, 0.5)])], material_slots=slots)


def test_f32_lens_walks_change_nothing(ctx, tmp_path, monkeypatch):
    """generate_ray_differential's +-0.05 px neighbour rays only decide whether a sample keeps its weight
    (camera.rs:582-628), and a main ray that is blocked in the lens only makes the weight zero.  The generate kernels
    answer those yes / no questions with an fp32 lens walk that abstains near every decision boundary
    (csrc/camera.cuh lens_walk_from_film_f32): RRT_GEN_F32=2 is the screened kernel (the default), 1 the lane state
    machine with fp32 neighbour walks, 0 sends every ray through the f64 walk.  All three must give every camera sample
    the same weight, first hit and film, on the two camera set-ups of the BASELINE configs — and the oracle's
    zero-weight count (checked by every other render test) stays exact."""
    scenes_ = [synth.scene_c1(str(tmp_path / "c1"), xres=640, yres=360, nsamp=9),
               synth.scene_c4(str(tmp_path / "c4"), n_spheres=3000, xres=480, yres=270, nsamp=9, extent=12.0)]
    for path in scenes_:
        runs = {}
        for flag in ("0", "1", "2", "3"):
            monkeypatch.setenv("RRT_GEN_F32", flag)
            gpu = Render.load(ctx, path, seed=1)
            gpu.enable_hit_dump()
            gpu.run()
            runs[flag] = (gpu.hit_dump(), gpu.film(), gpu.stats())
        d0, f0, s0 = runs["0"]
        assert s0["f32_neighbours"] == 0 and s0["f32_unsure"] == 0
        for flag in ("1", "2", "3"):   # 3 = the screened kernel's two phases as kernels of their own
            d1, f1, s1 = runs[flag]
            assert s1["f32_neighbours"] > 2 * s1["camera_rays"] * 0.9, (flag, s1)
            assert s1["f32_unsure"] < 0.05 * s1["f32_neighbours"], (flag, s1)
            assert s0["camera_rays"] == s1["camera_rays"] and s0["zero_weight"] == s1["zero_weight"], (flag, s0, s1)
            assert s0["extension_rays"] == s1["extension_rays"] and s0["shadow_rays"] == s1["shadow_rays"], (flag, s0, s1)
            assert np.array_equal(d0, d1), flag   # pixel, sample, first primitive, t and weight of every camera sample
            assert np.allclose(f0, f1, rtol=1e-12, atol=0), flag


def test_shading_by_material_kind_changes_nothing(ctx, tmp_path, monkeypatch):
    """Constant-valued scenes shade a round's hits with one launch per material kind: an escaped-ray kernel, kernels whose
    Bsdf is narrowed to the lobe set of Matte / Plastic / Metal at compile time (csrc/shading.cuh "lobe sets"), and the
    general code over the bins of Mirror and Glass.  RRT_SHADE_BY_KIND=0 keeps the one general kernel.  Same functions, same
    operation order: every camera sample's first hit, every ray count and the film (to the rounding of its unordered
    atomic sums) must not move — on a scene with all five kinds (Lambert and Oren-Nayar Matte, smooth and rough glass) under
    the Path integrator, and on config 1's meshes under DirectLighting."""
    cases = [synth.scene_c4(str(tmp_path / "c4"), n_spheres=4000, xres=384, yres=216, nsamp=17, extent=12.0, extra_materials=True),
             synth.scene_c1(str(tmp_path / "c1"), xres=320, yres=180, nsamp=9, integrator="DirectLighting", max_depth=1)]
    for path in cases:
        runs = {}
        for flag in ("0", "1"):
            monkeypatch.setenv("RRT_SHADE_BY_KIND", flag)
            gpu = Render.load(ctx, path, seed=1)
            gpu.enable_hit_dump()
            gpu.run()
            runs[flag] = (gpu.hit_dump(), gpu.film(), gpu.stats())
        (d0, f0, s0), (d1, f1, s1) = runs["0"], runs["1"]
        # the per-kind launches did run: four (Matte, Plastic, Metal, Mirror + Glass) for one general launch per round on the
        # sphere field; config 1's primitives are all Matte, one launch either way
        assert s1["launches"] > s0["launches"] or "c1" in path, (s0, s1)
        for k in ("camera_rays", "zero_weight", "extension_rays", "shadow_rays", "bounces"):
            assert s0[k] == s1[k], (k, s0, s1)
        assert s0["bounces"] > 0 or "c1" in path
        assert np.array_equal(d0, d1)
        assert np.allclose(f0, f1, rtol=1e-12, atol=1e-300), float(np.abs(f0 - f1).max())


def test_chunk_size_changes_nothing(ctx, tmp_path, monkeypatch):
    """A frame is rendered in chunks of up to 2^26 camera samples, capped by the frame (every other test's frame is one
    chunk).  RRT_CHUNK_LOG2=16 cuts this one into 2^16-sample chunks — 19 of them, the last one partial: same camera
    samples, same ray counts, same film to the rounding of its atomic sums."""
    path = synth.scene_c4(str(tmp_path / "c4"), n_spheres=3000, xres=320, yres=240, nsamp=17, extent=12.0)
    runs = {}
    for log2 in ("26", "16"):
        monkeypatch.setenv("RRT_CHUNK_LOG2", log2)
        gpu = Render.load(ctx, path, seed=1)
        gpu.enable_hit_dump()
        gpu.run()
        runs[log2] = (gpu.hit_dump(), gpu.film(), gpu.stats())
    (d0, f0, s0), (d1, f1, s1) = runs["26"], runs["16"]
    assert s0["chunks"] == 1 and s1["chunks"] == (320 * 240 * 16 + 65535) // 65536, (s0, s1)
    for k in ("samples", "camera_rays", "zero_weight", "extension_rays", "shadow_rays", "bounces"):
        assert s0[k] == s1[k], (k, s0, s1)
    assert np.array_equal(d0, d1)
    assert np.allclose(f0, f1, rtol=1e-12, atol=1e-300), float(np.abs(f0 - f1).max())


@pytest.mark.parametrize("integrator", ["Path", "DirectLighting"])
def test_clipped_and_stretched_spheres(ctx, tmp_path, integrator):
    """SURVEY §8a6 through the renderer: spheres clipped in z / phi (their inside seen through the opening: hits at the
    far root, whose point comes from the object-space ray), under non-uniform object and instance transforms, with
    uv-driven textures.  First hits bit-equal, film to rounding."""
    path = synth.scene_clipped_spheres(str(tmp_path / "s"), xres=256, yres=144, nsamp=9, integrator=integrator,
                                       max_depth=5 if integrator == "Path" else 1)
    ref = S.load(path).render(seed=1, want_dump=True)
    gpu = Render.load(ctx, path, seed=1)
    gpu.enable_hit_dump()
    gpu.run()
    out = compare(gpu, ref)
    assert out["rmse"] < 1e-4, out
    first = ref["dump"][:, 3]
    assert all((first == k).sum() > 200 for k in range(6))   # every sphere of the scene is seen by camera rays


def test_direct_lighting_sample_all_lights(ctx, tmp_path):
    """light_strategy "all": one estimate per light and hit (Q30), point lights (config 1's three) and area lights."""
    every = {"Integrator": {"integrator_type": "DirectLighting", "max_depth": 1, "light_strategy": "all"}}
    for path in (synth.scene_c1(str(tmp_path / "c1"), xres=256, yres=144, nsamp=9, integrator="DirectLighting", max_depth=1),
                 synth.scene_area_lights(str(tmp_path / "a"), xres=192, yres=108, nsamp=9, integrator="DirectLighting", max_depth=1)):
        ref = S.load(path, every).render(seed=1, want_dump=True)
        gpu = Render.load(ctx, path, overrides=every, seed=1)
        gpu.enable_hit_dump()
        gpu.run()
        out = compare(gpu, ref)
        assert out["rmse"] < 1e-6, out
        assert out["shadow_rays"][0] == out["shadow_rays"][1], out
        one = Render.load(ctx, path, seed=1)
        one.run()
        assert one.stats()["shadow_rays"] < out["shadow_rays"][0]


def test_full_size_config5(ctx):
    """BASELINE config 5 at FULL size — the 4,194,304-triangle soup, 3840 x 2160, Halton nsamp 257 (256 rendered), Path
    max_depth 5 — against the oracle on what the oracle finishes in seconds: (1) a 64 Ki-ray slice of incoherent
    bounce rays on the 4 Mi-triangle aggregate (primitive, t, u, v; any-hit), (2) a centred 48 x 48-pixel crop of
    the frame at full spp: every camera ray's first primitive and t, the filter weights, the ray counts, the film."""
    import scenes
    agg, gpu = synth.scene_c5_api(ctx)           # defaults = config 5 as bench.py renders it
    p, idx = synth.soup_triangles(1 << 22, 0.006, synth.SEED_C5_SOUP)
    rays = synth.bounce_rays(p, idx, 1 << 16, seed=77)
    hits = agg.intersect(rays)
    occ = agg.intersect_p(rays)
    oscene = scenes.oracle_soup(p, idx)
    ref = oscene.intersect(rays)
    c = scenes.compare_closest(hits, ref["prim"], ref["t"])
    assert c["mismatch_excl_ties"] == 0 and c["t_bad"] == 0, c
    occ_ref, _ = oscene.intersect_p(rays)
    assert np.array_equal(occ.astype(bool), np.asarray(occ_ref).astype(bool))
    del oscene
    crop = (1896, 1056, 1944, 1104)
    gpu.enable_hit_dump()
    gpu.run(crop=crop)
    ref = scenes.oracle_c5(1 << 22, 0.006, 3840, 2160, 257, crop=crop, want_dump=True)
    rgb, raw = gpu.film(want_raw=True)
    st = gpu.stats()
    sl = (slice(crop[1], crop[3]), slice(crop[0], crop[2]))
    assert np.array_equal(raw[sl][..., 3], ref["raw"][sl][..., 3])
    assert st["camera_rays"] == ref["stats"]["camera_rays"] and st["zero_weight"] == ref["stats"]["zero_weight"]
    assert st["camera_rays"] + st["zero_weight"] == 48 * 48 * 256
    d, r = gpu.hit_dump(), ref["dump"]
    d = d[~np.isnan(d[:, 0])]
    assert d.shape == r.shape and np.array_equal(d[:, :3], r[:, :3])
    assert np.array_equal(d[:, 3], r[:, 3])                      # first-hit primitive of every camera ray
    hit = r[:, 3] >= 0
    assert np.allclose(d[hit, 4], r[hit, 4], rtol=1e-5, atol=0)   # its t (north_star: 1e-5 relative)
    assert abs(st["extension_rays"] - ref["stats"]["extension_rays"]) <= 8
    assert abs(st["shadow_rays"] - ref["stats"]["shadow_rays"]) <= 8
    assert rel_rmse(rgb[sl], ref["rgb"][sl]) <= REL_RMSE
    gpu.close()


def test_film_gather_and_tree_replication(ctx, tmp_path):
    """The multi-GPU plumbing of the C ABI on one device: (1) rrt_scene_export_tree / rrt_scene_commit_from_tree — a
    scene committed from another scene's tree answers every ray identically; (2) rrt_film_gather — G renderers, each
    having rendered its tiles t % G == r, gathered onto renderer 0 by moving only the owned tiles' pixels: the frame
    equals the single-renderer frame bit for bit (box filter 0.5: every pixel has one owner)."""
    import ctypes as C
    import scenes
    from rs_ray_toy_b200 import capi
    from rs_ray_toy_b200.aggregate import GpuAggregate
    p, idx = scenes.soup(30000)
    a = scenes.gpu_soup(ctx, p, idx)
    blob = a.export_tree()
    b = GpuAggregate(ctx)
    b.add_triangles(b.add_mesh(p, idx), 0)
    b.commit_from_tree(blob)
    rays = synth.bounce_rays(p, idx, 50000, seed=3)
    assert np.array_equal(a.intersect(rays), b.intersect(rays)) and np.array_equal(a.intersect_p(rays), b.intersect_p(rays))
    c = GpuAggregate(ctx)
    c.add_triangles(c.add_mesh(p[:300], idx[:100]), 0)
    with pytest.raises(capi.RrtError):      # a tree over other primitives
        c.commit_from_tree(blob)
    with pytest.raises(capi.RrtError):
        c.commit_from_tree(blob[:1000])

    path = synth.scene_c4(str(tmp_path / "c4"), n_spheres=1500, xres=200, yres=120, nsamp=5, extent=10.0)
    full = Render.load(ctx, path, seed=1)
    full.run()
    _, raw_full = full.film(want_raw=True)
    L = capi.lib()
    for G in (2, 3):
        parts = [Render.load(ctx, path, seed=1) for _ in range(G)]
        for r, part in enumerate(parts):
            part.run(tile_mod=G, tile_rank=r)
        assert sum(part.owned_doubles(G, r) for r, part in enumerate(parts)) == 1024 * 13 * 8   # ceil(200/16) x ceil(120/16) tiles
        handles = (C.c_void_p * G)(*[part.h for part in parts])
        capi.check(L.rrt_film_gather(handles, G, 0))
        assert np.array_equal(parts[0].film(want_raw=True)[1], raw_full)
        for part in parts:
            part.close()
    # a filter wider than a pixel: tiles do not own their pixels, the gather refuses and the caller sums films instead
    wide = Render.load(ctx, path, overrides={"Film": {"xres": 200, "yres": 120, "diagonal": 35, "Filter": {"filter_type": "GaussianFilter"}}}, seed=1)
    with pytest.raises(capi.RrtError) as e:
        wide.owned_doubles(2, 0) and wide.pack_owned(2, 0, 1, 1 << 40)
    assert e.value.status == capi.RRT_ERR_UNSUPPORTED
