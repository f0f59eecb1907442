"""DirectLighting with its specular recursion, the IntersectDebug integrator and the StratifiedSampler on the device
(whitted_kernel in csrc/render.cu) against the oracle — including config 1 exactly as the reference ships it."""
import json

import numpy as np
import pytest

import oracle_lib as O
import oracle_scene as S
from rs_ray_toy_b200 import capi, synth
from rs_ray_toy_b200.aggregate import Context
from rs_ray_toy_b200.render import Render

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return Context(0)


def rel_rmse(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-300))


def check(gpu, ref, rmse=1e-9, ray_slack=0):
    rgb, raw = gpu.film(want_raw=True)
    st = gpu.stats()
    assert np.array_equal(raw[..., 3], ref["raw"][..., 3])            # filter weights: sample counts, exact
    assert st["camera_rays"] == ref["stats"]["camera_rays"] and st["zero_weight"] == ref["stats"]["zero_weight"]
    d, r = gpu.hit_dump(), ref["dump"]
    assert d.shape == r.shape and np.array_equal(d[:, :4], r[:, :4])  # pixel, sample, first primitive of every camera ray
    hit = r[:, 3] >= 0
    assert np.allclose(d[hit, 4], r[hit, 4], rtol=1e-5, atol=0)
    assert abs(st["extension_rays"] - ref["stats"]["extension_rays"]) <= ray_slack, (st, ref["stats"])
    assert abs(st["shadow_rays"] - ref["stats"]["shadow_rays"]) <= ray_slack, (st, ref["stats"])
    e = rel_rmse(rgb, ref["rgb"])
    assert e <= rmse, e
    return e


@pytest.mark.parametrize("tier", ["F", "L"])
def test_config1_as_shipped(ctx, tmp_path, tier):
    """samples/scene.json with NO overrides: Debug integrator, StratifiedSampler 4 x 4 (15 rendered, Q10), three point
    lights at the origin (Q17), an unreferenced ImageTexture declaration.  Tier L = the reference's own tree and accept
    rules."""
    path = synth.scene_c1_as_shipped(str(tmp_path / "c1"))
    ref = S.load(path, tier=O.TIER_L if tier == "L" else O.TIER_F).render(seed=1, want_dump=True)
    gpu = Render.load(ctx, path, seed=1, literal=(tier == "L"))
    gpu.enable_hit_dump()
    gpu.run()
    assert ref["stats"]["camera_rays"] > 900000 and ref["stats"]["shadow_rays"] > 100000
    check(gpu, ref)
    gpu.close()


def specular_scene(tmp_path, name):
    # config 4's sphere field with the Mirror / smooth Glass / rough Glass / Oren-Nayar presets, dense enough for
    # reflections of reflections
    return synth.scene_c4(str(tmp_path / name), n_spheres=1500, xres=200, yres=120, nsamp=5, extent=10.0, extra_materials=True)


AREA_AND_POINT = [{"light_type": "diffuse", "spectrum": {"values": [60, 50, 40]},
                   "light_shape": {"shape_type": "sphere", "radius": 3.0, "world_pos": [0.0, 14.0, -6.0]}},
                  {"light_type": "point", "spectrum": {"values": [3000, 3000, 3000]}}]


@pytest.mark.parametrize("strategy", ["one", "all"])
@pytest.mark.parametrize("lights", ["delta", "area"])
def test_direct_lighting_specular_recursion(ctx, tmp_path, strategy, lights):
    """integrator/mod.rs:150-301 under DirectLighting, max_depth 4: mirrors (a chain) and smooth glass (a fork at every
    hit).  With an area light the values of the light draws matter, so the depth-first dimension accounting of the
    wavefront is what is being checked."""
    path = specular_scene(tmp_path, "s")
    ov = {"Integrator": {"integrator_type": "DirectLighting", "max_depth": 4, "light_strategy": strategy}}
    if lights == "area":
        ov["lights"] = AREA_AND_POINT
    ref = S.load(path, ov).render(seed=1, want_dump=True)
    gpu = Render.load(ctx, path, overrides=ov, seed=1)
    gpu.enable_hit_dump()
    gpu.run()
    assert ref["stats"]["extension_rays"] > 1.2 * ref["stats"]["camera_rays"]      # the recursion really runs
    check(gpu, ref, ray_slack=4)
    shallow = S.load(path, dict(ov, Integrator=dict(ov["Integrator"], max_depth=1))).render(seed=1)
    assert rel_rmse(shallow["rgb"], ref["rgb"]) > 1e-3                              # and changes the picture
    gpu.close()


@pytest.mark.parametrize("sampler", ["HaltonSampler", "StratifiedSampler"])
def test_debug_integrator_with_recursion(ctx, tmp_path, sampler):
    path = specular_scene(tmp_path, "d")
    smp = {"sampler_type": sampler, "nsamp": 5} if sampler == "HaltonSampler" else \
        {"sampler_type": sampler, "xsamp": 3, "ysamp": 2, "dimension": 3, "jitter": True}
    ov = {"Integrator": {"integrator_type": "Debug", "max_depth": 4}, "Sampler": smp, "lights": AREA_AND_POINT}
    ref = S.load(path, ov).render(seed=3, want_dump=True)
    gpu = Render.load(ctx, path, overrides=ov, seed=3)
    gpu.enable_hit_dump()
    gpu.run()
    check(gpu, ref, ray_slack=4)
    gpu.close()


def test_stratified_sampler_with_direct_lighting(ctx, tmp_path):
    """PixelSampler<Stratified>: `dimension` 1 leaves the lens sample and every light draw to the U[-1, 1) overflow
    stream (Q12) — negative lens coordinates and all; jitter off puts every film sample at its stratum's centre."""
    path = synth.scene_area_lights(str(tmp_path / "a"), xres=160, yres=90, nsamp=9, integrator="DirectLighting", max_depth=1)
    for smp in ({"sampler_type": "StratifiedSampler", "xsamp": 2, "ysamp": 3, "dimension": 1},
                {"sampler_type": "StratifiedSampler", "xsamp": 4, "ysamp": 4, "dimension": 6, "jitter": False},
                {"sampler_type": "StratifiedSampler"}):
        for strategy in ("one", "all"):
            ov = {"Sampler": smp, "Integrator": {"integrator_type": "DirectLighting", "max_depth": 1, "light_strategy": strategy}}
            ref = S.load(path, ov).render(seed=5, want_dump=True)
            gpu = Render.load(ctx, path, overrides=ov, seed=5)
            gpu.enable_hit_dump()
            gpu.run()
            check(gpu, ref)
            gpu.close()


def test_limits_are_refused_not_approximated(ctx, tmp_path):
    path = specular_scene(tmp_path, "r")
    # 15 hits x 12 Halton dimensions: beyond the 128-dimension tables
    with pytest.raises(capi.RrtError):
        Render.load(ctx, path, overrides={"Integrator": {"integrator_type": "DirectLighting", "max_depth": 5, "light_strategy": "all"}})
    # Path + StratifiedSampler: BSDF sampling from the U[-1, 1) overflow stream
    with pytest.raises(capi.RrtError):
        Render.load(ctx, path, overrides={"Sampler": {"sampler_type": "StratifiedSampler"}})
    with pytest.raises(capi.RrtError):
        Render.load(ctx, path, overrides={"Sampler": {"sampler_type": "StratifiedSampler", "xsamp": 32, "ysamp": 32},
                                          "Integrator": {"integrator_type": "Debug"}})
    with pytest.raises(capi.RrtError):
        Render.load(ctx, path, overrides={"Integrator": {"integrator_type": "Debug", "max_depth": 12}})


@pytest.mark.parametrize("integrator", ["Path", "DirectLighting-one", "DirectLighting-all", "Debug"])
def test_environment_light_and_image_textures(ctx, tmp_path, integrator):
    """an InfiniteAreaLight in `lights` (next-event estimation + the live BSDF-sampling half of
    estimate_direct, whose ray only has to escape) and in `infinite_lights` (escaped camera rays and specular bounces of
    the Path integrator; DirectLighting reads the FIRST entry of `lights` on a miss), and ImageTextures through the
    reference's MIPMap (EWA, trilinear, three wrap modes; Q31 / Q32 included) — against the oracle.  Transcendentals
    differ in the last place between the device and libm, and MIPMap levels / texel indices are floors of them: the bar
    is north_star's image tolerance, the counts are exact."""
    path = synth.scene_env_and_images(str(tmp_path / "e"), xres=192, yres=108, nsamp=9)
    kind, _, strategy = integrator.partition("-")
    ov = {"Integrator": {"integrator_type": kind, "max_depth": 4, "light_strategy": strategy or "one"}}
    ref = S.load(path, ov).render(seed=1, want_dump=True)
    gpu = Render.load(ctx, path, overrides=ov, seed=1)
    gpu.enable_hit_dump()
    gpu.run()
    rgb, raw = gpu.film(want_raw=True)
    st = gpu.stats()
    assert np.array_equal(raw[..., 3], ref["raw"][..., 3])
    assert st["camera_rays"] == ref["stats"]["camera_rays"]
    d, r = gpu.hit_dump(), ref["dump"]
    assert np.array_equal(d[:, :4], r[:, :4])
    assert abs(st["extension_rays"] - ref["stats"]["extension_rays"]) <= 8
    # the device's shadow queue also carries the BSDF-sampled probes the oracle counts apart
    assert abs(st["shadow_rays"] - ref["stats"]["shadow_rays"] - ref["stats"]["mis_probe_rays"]) <= 8
    if kind != "Debug":
        assert ref["stats"]["mis_probe_rays"] > 0
    assert ref["rgb"].mean() > 1e-3
    assert rel_rmse(rgb, ref["rgb"]) <= 1e-3
    # the environment really is what lights the miss pixels (Path, DirectLighting): the corner of the frame is sky
    if kind != "Debug":
        assert ref["rgb"][:8, :8].mean() > 0
    gpu.close()


def test_image_inputs_are_checked(ctx, tmp_path):
    import json as js
    path = synth.scene_env_and_images(str(tmp_path / "c"), xres=64, yres=36, nsamp=3)
    cfg = js.loads(open(path).read())
    # a referenced image that does not exist: an I/O error; an unreferenced one: ignored (the sample scene has one)
    bad = [dict(t) for t in cfg["rgb_texture"]]
    bad[0]["filename"] = "nowhere.png"
    with pytest.raises(capi.RrtError) as e:
        Render.load(ctx, path, overrides={"rgb_texture": bad})
    assert e.value.status == capi.RRT_ERR_IO
    extra = cfg["rgb_texture"] + [{"texture_name": "unused", "texture_type": "ImageTexture", "filename": "nowhere.png"}]
    Render.load(ctx, path, overrides={"rgb_texture": extra}).close()
    # an EWA-filtered image with a single MIPMap level: the reference panics at the first filtered lookup
    synth.write_test_png(str(tmp_path / "c" / "small.png"), 100, 60, seed=9)
    small = [dict(t) for t in cfg["rgb_texture"]]
    small[0]["filename"] = "small.png"
    with pytest.raises(capi.RrtError) as e:
        Render.load(ctx, path, overrides={"rgb_texture": small})
    assert e.value.status == capi.RRT_ERR_UNSUPPORTED
    with pytest.raises(capi.RrtError):    # the literal tier's shadow queue cannot carry the probe ray (Q4)
        Render.load(ctx, path, literal=True)
