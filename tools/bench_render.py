"""Experiment driver (GPU box): times the wavefront renderer on configs 2 / 4 / 5 (reduced sizes by
flag) and prints Msamples/s with the ray counts.  Not part of the product or the tests."""
import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rs_ray_toy_b200 import synth  # noqa: E402
from rs_ray_toy_b200.aggregate import Context  # noqa: E402
from rs_ray_toy_b200.render import Render  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c4")
ap.add_argument("--scale", type=float, default=1.0, help="resolution scale")
ap.add_argument("--nsamp", type=int, default=0)
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--textured", action="store_true", help="config 5 with a 3D checkerboard kd and a textured roughness")
a = ap.parse_args()
ctx = Context(0)
d = tempfile.mkdtemp()
t0 = time.time()
if a.config == "c4":
    path = synth.scene_c4(d, n_spheres=a.n or 100000, xres=int(1920 * a.scale), yres=int(1080 * a.scale), nsamp=a.nsamp or 65)
    r = Render.load(ctx, path, seed=1)
elif a.config == "c2":
    path = synth.scene_c2(d, n_instances=a.n or 10000, xres=int(1920 * a.scale), yres=int(1080 * a.scale), nsamp=a.nsamp or 2)
    r = Render.load(ctx, path, seed=1)
elif a.config == "c1":
    path = synth.scene_c1(d, nsamp=a.nsamp or 17)
    r = Render.load(ctx, path, seed=1)
else:
    agg, r = synth.scene_c5_api(ctx, n_tris=a.n or (1 << 22), xres=int(3840 * a.scale), yres=int(2160 * a.scale), nsamp=a.nsamp or 257,
                                  textured=a.textured)
setup = time.time() - t0
for rep in range(a.reps):
    r.clear()
    t0 = time.time()
    r.run()
    dt = time.time() - t0
    s = r.stats()
    print(json.dumps({"config": a.config, "rep": rep, "setup_s": round(setup, 2), "render_s": round(dt, 3),
                      "Msamples_per_s": round(s["samples"] / dt / 1e6, 2),
                      "Mrays_per_s": round((s["extension_rays"] + s["shadow_rays"]) / dt / 1e6, 2), **s}), flush=True)
img = r.film()
print("mean rgb", img.mean(axis=(0, 1)).tolist(), "max", float(img.max()))
