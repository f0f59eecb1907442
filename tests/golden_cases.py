"""Definitions of the committed golden cases (tests/golden/*.npz) — small seeded versions of
BASELINE.json configs 2, 3 and 4.  `make_golden.py` writes them with the CPU oracle; the
`-m gpu` parity tests replay them through the CUDA library without the oracle in the loop."""
import numpy as np

import oracle_lib as O
import scenes
from rs_ray_toy_b200 import synth

CASES = {
    "soup_c3_small": dict(kind="soup", n=30000, rays=8000),
    "cubes_c2_small": dict(kind="cubes", n=400, rays=8000),
    "spheres_c4_small": dict(kind="spheres", n=2000, rays=8000),
}


def inputs(name):
    c = CASES[name]
    if c["kind"] == "soup":
        p, idx = scenes.soup(c["n"])
        rays = synth.bounce_rays(p, idx, c["rays"])
        return dict(p=p, idx=idx), rays
    if c["kind"] == "cubes":
        m, inv = scenes.cube_instances(c["n"], extent=25.0)
        rays = synth.camera_like_rays(c["rays"], (0.0, 0.0, -80.0), 25.0)
        return dict(m=m, inv=inv), rays
    m, inv = scenes.sphere_instances(c["n"], extent=15.0)
    rays = synth.camera_like_rays(c["rays"], (0.0, 0.0, -40.0), 15.0)
    return dict(m=m, inv=inv), rays


def shadow_rays(name, rays):
    kind = CASES[name]["kind"]
    if kind == "soup":
        return synth.shadow_rays_from(rays, (0.5, 0.5, 1.5))
    # camera rays share one origin; cast the shadow rays from random points inside the scene box
    extent, light = {"cubes": (25.0, (0.0, 40.0, 0.0)), "spheres": (15.0, (0.0, 25.0, 0.0))}[kind]
    rng = np.random.Generator(np.random.PCG64(21))
    o = np.zeros_like(rays)
    o[:, 0:3] = (rng.random((rays.shape[0], 3)) * 2 - 1) * extent
    return synth.shadow_rays_from(o, light)


def build_oracle(name, tier=O.TIER_F):
    geo, rays = inputs(name)
    kind = CASES[name]["kind"]
    if kind == "soup":
        return scenes.oracle_soup(geo["p"], geo["idx"], tier), rays
    if kind == "cubes":
        return scenes.oracle_cubes(geo["m"], geo["inv"], tier), rays
    return scenes.oracle_spheres(geo["m"], geo["inv"], 0.5, tier), rays


def build_gpu(ctx, name):
    geo, rays = inputs(name)
    kind = CASES[name]["kind"]
    if kind == "soup":
        return scenes.gpu_soup(ctx, geo["p"], geo["idx"]), rays
    if kind == "cubes":
        return scenes.gpu_cubes(ctx, geo["m"], geo["inv"]), rays
    return scenes.gpu_spheres(ctx, geo["m"], geo["inv"], 0.5), rays
