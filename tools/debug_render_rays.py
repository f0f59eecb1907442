"""Parity debugging aid (GPU box): renders config 1 on both sides, finds the pixels whose film values differ,
re-renders those pixels in the oracle with its ray log on, and replays every logged ray through the GPU aggregate
(rrt_intersect / rrt_intersect_p) to find the first ray on which the two sides part.

usage: python tools/debug_render_rays.py [--literal] [--nsamp 9] [--max-pixels 12]
"""
import argparse
import ctypes as C
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np

import oracle_lib as O
import oracle_scene as S
from rs_ray_toy_b200 import synth
from rs_ray_toy_b200.aggregate import Context, GpuAggregate
from rs_ray_toy_b200.render import Render

ap = argparse.ArgumentParser()
ap.add_argument("--literal", action="store_true")
ap.add_argument("--nsamp", type=int, default=9)
ap.add_argument("--max-pixels", type=int, default=12)
ap.add_argument("--gpu-log", default=None, help="x,y: with RRT_LIB pointing at a -DRRT_DEBUG_PIXEL_X/Y build, print both sides' rays of that pixel")
args = ap.parse_args()

if args.gpu_log:
    x, y = map(int, args.gpu_log.split(","))
    ctx = Context(0)
    path = synth.scene_c1(tempfile.mkdtemp(prefix="rrt_dbg_"), nsamp=args.nsamp)
    tier = O.TIER_L if args.literal else O.TIER_F
    ref_scene = S.load(path, tier=tier)
    L = O.lib()
    L.orc_raylog_take.restype = C.c_uint64
    L.orc_raylog_take.argtypes = [C.c_void_p, C.c_uint64]
    L.orc_raylog_begin()
    ref_scene.render(seed=1, nthreads=1, crop=(x, y, x + 1, y + 1))
    log = np.zeros((4096, 10))
    n = L.orc_raylog_take(log.ctypes.data, 4096)
    for r in log[:n]:
        print("ORC", "closest" if r[0] == 0 else "shadow ", " ".join(float(v).hex() for v in r[1:8]), "->", int(r[8]), float(r[9]).hex())
    sys.stdout.flush()
    gpu = Render.load(ctx, path, seed=1, literal=args.literal)
    gpu.run(crop=(x, y, x + 1, y + 1))
    print("gpu pixel", gpu.film()[y, x])
    sys.exit(0)

ctx = Context(0)
path = synth.scene_c1(tempfile.mkdtemp(prefix="rrt_dbg_"), nsamp=args.nsamp)
tier = O.TIER_L if args.literal else O.TIER_F
ref_scene = S.load(path, tier=tier)
ref = ref_scene.render(seed=1)
gpu = Render.load(ctx, path, seed=1, literal=args.literal)
gpu.run()
img, raw = gpu.film(want_raw=True)
diff = np.abs(img - ref["rgb"]).max(axis=2)
rms = float(np.sqrt(np.mean(ref["rgb"] ** 2)))
print("image rms", rms, "rel rmse", float(np.sqrt(np.mean((img - ref["rgb"]) ** 2))) / rms)
ys, xs = np.nonzero(diff > 1e-9 * max(rms, 1e-300))
order = np.argsort(-diff[ys, xs])
print("differing pixels:", len(ys))
for k in order[:20]:
    print("  px", xs[k], ys[k], "gpu", img[ys[k], xs[k]], "ref", ref["rgb"][ys[k], xs[k]])

agg = GpuAggregate.__new__(GpuAggregate)
agg.ctx, agg.L, agg.h, agg.committed = ctx, ctx.L, gpu.scene_h, True
L = O.lib()
L.orc_raylog_take.restype = C.c_uint64
L.orc_raylog_take.argtypes = [C.c_void_p, C.c_uint64]
for k in order[: args.max_pixels]:
    x, y = int(xs[k]), int(ys[k])
    L.orc_raylog_begin()
    ref_scene.render(seed=1, nthreads=1, crop=(x, y, x + 1, y + 1))
    n = L.orc_raylog_take(None, 0)
    # take() cleared the pointer; log again to fetch (two-step keeps the C side trivial)
    L.orc_raylog_begin()
    ref_scene.render(seed=1, nthreads=1, crop=(x, y, x + 1, y + 1))
    log = np.zeros((n, 10))
    L.orc_raylog_take(log.ctypes.data, n)
    closest = log[log[:, 0] == 0]
    shadow = log[log[:, 0] == 1]
    print(f"pixel ({x},{y}): {len(closest)} closest + {len(shadow)} shadow rays logged")
    if len(closest):
        h = agg.intersect(np.ascontiguousarray(closest[:, 1:8]))
        gp = h["prim_id"].astype(np.int64)
        gp[gp == 0xFFFFFFFF] = -1
        bad = np.nonzero((gp != closest[:, 8].astype(np.int64)) | ((gp >= 0) & (h["t"] != closest[:, 9])))[0]
        for i in bad[:6]:
            print("   closest ray", i, "gpu", gp[i], repr(float(h["t"][i])), "ref", int(closest[i, 8]), repr(float(closest[i, 9])),
                  "ray", [repr(float(v)) for v in closest[i, 1:8]])
    if len(shadow):
        occ = agg.intersect_p(np.ascontiguousarray(shadow[:, 1:8]))
        bad = np.nonzero(occ.astype(np.int64) != shadow[:, 8].astype(np.int64))[0]
        for i in bad[:6]:
            print("   shadow ray", i, "gpu", int(occ[i]), "ref", int(shadow[i, 8]), "ray", [repr(float(v)) for v in shadow[i, 1:8]])
agg.h = None
