"""Diagnostic (GPU box): the device-side timeline of one rrt_intersect call over host buffers (RRT_HOST_TRACE=1)."""
import os
import sys
from pathlib import Path

os.environ["RRT_HOST_TRACE"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

from rs_ray_toy_b200 import synth  # noqa: E402
from rs_ray_toy_b200.aggregate import HIT_DTYPE, RAY_DTYPE, Context, pack_rays, soup_aggregate  # noqa: E402

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
ctx = Context(0)
p, idx = synth.soup_triangles(1 << 20)
agg = soup_aggregate(ctx, p, idx, 4)
rays = synth.bounce_rays(p, idx, n_rays)
h_rays = ctx.pinned_empty(n_rays, RAY_DTYPE)
h_rays[:] = pack_rays(rays)
h_hits = ctx.pinned_empty(n_rays, HIT_DTYPE)
for i in range(3):
    print(f"--- call {i}", file=sys.stderr, flush=True)
    agg.intersect(h_rays, out=h_hits)
