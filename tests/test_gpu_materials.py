"""TranslucentMaterial, DisneyMaterial and the Debug material on the device (the eight-lobe kernels of csrc/render.cu:
shade_kernel<.., BIG> and whitted_kernel<.., BIG>) against the oracle — SURVEY §8f row 3's remaining materials."""
import json

import numpy as np
import pytest

import oracle_scene as S
from rs_ray_toy_b200 import capi, synth
from rs_ray_toy_b200 import render as R
from rs_ray_toy_b200.aggregate import Context
from rs_ray_toy_b200.render import Render

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return Context(0)


def rel_rmse(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-300))


def compare(ctx, path, ov, rmse, slack, env=False):
    ref = S.load(path, ov).render(seed=1, want_dump=True)
    gpu = Render.load(ctx, path, overrides=ov, seed=1)
    gpu.enable_hit_dump()
    gpu.run()
    rgb, raw = gpu.film(want_raw=True)
    st = gpu.stats()
    assert np.array_equal(raw[..., 3], ref["raw"][..., 3])            # filter weights: sample counts, exact
    assert st["camera_rays"] == ref["stats"]["camera_rays"] and st["zero_weight"] == ref["stats"]["zero_weight"]
    d, r = gpu.hit_dump(), ref["dump"]
    assert d.shape == r.shape and np.array_equal(d[:, :4], r[:, :4])  # pixel, sample, first primitive of every camera ray
    assert abs(st["extension_rays"] - ref["stats"]["extension_rays"]) <= slack, (st, ref["stats"])
    probes = ref["stats"]["mis_probe_rays"] if env else 0            # the device's shadow queue carries them too
    assert abs(st["shadow_rays"] - ref["stats"]["shadow_rays"] - probes) <= slack, (st, ref["stats"])
    assert ref["stats"]["asserts"] == 0
    assert np.isfinite(ref["rgb"]).all() and ref["rgb"].mean() > 1e-4
    e = rel_rmse(rgb, ref["rgb"])
    assert e <= rmse, e
    gpu.close()
    return ref, e


@pytest.mark.parametrize("integrator", ["Path", "DirectLighting-one", "DirectLighting-all", "Debug"])
def test_translucent_disney_debug_materials(ctx, tmp_path, integrator):
    """Every lobe the two materials can build — LambertianTransmission, DisneyDiffuse / FakeSS / Retro / Sheen / Clearcoat
    (typed neither reflection nor transmission, Q37), DisneyFresnel over the separable-G distribution, the thin and the
    solid MicrofacetTransmission — with constant and texture-driven parameters, through each integrator.  pow / log10 /
    sin / cos differ in the last place between the device and libm, so a handful of paths may choose another lobe or
    Russian-roulette outcome: the ray counts get that much slack, the image is held to 1e-3 relative RMSE."""
    kind, _, strategy = integrator.partition("-")
    path = synth.scene_more_materials(str(tmp_path / "m"), xres=192, yres=108, nsamp=9, integrator=kind)
    ov = {"Integrator": {"integrator_type": kind, "max_depth": 5, "light_strategy": strategy or "one"}}
    ref, e = compare(ctx, path, ov, rmse=1e-3, slack=8)
    if kind == "Path":
        assert ref["stats"]["bounces"] > 2000
    print(f"{integrator}: relative RMSE {e:.3e}")


def test_materials_under_an_environment_light(ctx, tmp_path):
    """The same scene lit by an InfiniteAreaLight: estimate_direct's BSDF-sampling half samples the eight-lobe Bsdf
    (Bsdf::sample_f's Q15 rule adds the other lobes' pdfs only when the chosen lobe is not typed reflective — the
    clearcoat and the transmissive lobes)."""
    path = synth.scene_more_materials(str(tmp_path / "e"), xres=160, yres=90, nsamp=9, env=True)
    ov = {"Integrator": {"integrator_type": "Path", "max_depth": 4}}
    ref, e = compare(ctx, path, ov, rmse=1e-3, slack=8, env=True)
    assert ref["stats"]["mis_probe_rays"] > 0
    print(f"env: relative RMSE {e:.3e}")


def test_loader_records_are_what_the_constructors_build(ctx, tmp_path):
    """rrt_material as the loader fills it == as render.disney / render.translucent fill it (the records a caller of
    rrt_scene_set_materials would pass), and a two-material subset renders."""
    path = synth.scene_more_materials(str(tmp_path / "a"), xres=96, yres=54, nsamp=5)
    cfg = json.loads(open(path).read())
    keep = {"m_disney_metal": R.disney(color=(0.9, 0.7, 0.3), metallic=0.85, anisotropic=0.6, roughness=0.35),
            "m_trans": R.translucent(kd=(0.3, 0.5, 0.4), ks=(0.3, 0.3, 0.3), reflect=(0.5,) * 3, transmit=(0.6, 0.5, 0.7), roughness=0.2)}
    cfg["materials"] = [m for m in cfg["materials"] if m["material_name"] in keep]
    cfg["Aggregate"]["primitives"] = [p for p in cfg["Aggregate"]["primitives"] if p["material_name"] in keep]
    p2 = tmp_path / "a" / "two.json"
    p2.write_text(json.dumps(cfg))
    a = Render.load(ctx, str(p2), seed=1)
    a.run()
    want = a.film()
    a.close()
    # the loader's records are what the constructors build
    _, mats, _ = R.json_texture_probe(str(p2))
    for m, name in zip(mats, [m["material_name"] for m in cfg["materials"]]):
        assert bytes(m) == bytes(keep[name])
    assert np.isfinite(want).all() and want.mean() > 1e-5


def test_bssrdf_and_unknown_kinds_are_refused(ctx, tmp_path):
    path = synth.scene_more_materials(str(tmp_path / "r"), xres=64, yres=36, nsamp=3)
    cfg = json.loads(open(path).read())
    cfg["rgb_texture"].append({"texture_name": "c_sd", "texture_type": "BilerpTexture", "v00": {"values": [0.1, 0.2, 0.3]},
                               "v01": {"values": [0.1, 0.2, 0.3]}})
    cfg["materials"][2]["scatter_distance"] = "c_sd"              # not thin: disney.rs:588-606 builds a SeparableBSSRDF
    p2 = tmp_path / "r" / "sss.json"
    p2.write_text(json.dumps(cfg))
    with pytest.raises(capi.RrtError, match="BSSRDF"):
        Render.load(ctx, str(p2), seed=1)
    with pytest.raises(RuntimeError, match="BSSRDF"):
        S.load(str(p2)).render(seed=1)
    cfg["materials"][2]["thin"] = True                            # thin: scatter_distance is never read
    p3 = tmp_path / "r" / "thin.json"
    p3.write_text(json.dumps(cfg))
    Render.load(ctx, str(p3), seed=1).close()
