"""Experiment driver (GPU box): commit time and closest-hit throughput of the host SAH builder against the
device LBVH builder (RRT_BUILD_DEVICE_LBVH) on the config-3 / config-5 soups.  Not part of the product or the tests."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from rs_ray_toy_b200 import capi, synth
from rs_ray_toy_b200.aggregate import RAY_DTYPE, Context, pack_rays, soup_aggregate

ctx = Context(0)
for n_tris, edge, seed in ((1 << 20, 0.01, synth.SEED_C3_SOUP), (1 << 22, 0.006, synth.SEED_C5_SOUP)):
    p, idx = synth.soup_triangles(n_tris, edge, seed)
    n_rays = 1 << 23
    rays = pack_rays(synth.bounce_rays(p, idx, n_rays, seed=synth.SEED_C3_RAYS))
    d_rays = torch.from_numpy(rays.view(np.float64)).cuda()
    d_hits = torch.empty(n_rays * 4, dtype=torch.float64, device="cuda")
    prims = {}
    for name, flags, leaf in (("host SAH", capi.RRT_BUILD_FAST, 4), ("device LBVH", capi.RRT_BUILD_DEVICE_LBVH, 4),
                              ("device LBVH leaf 2", capi.RRT_BUILD_DEVICE_LBVH, 2), ("device LBVH leaf 1", capi.RRT_BUILD_DEVICE_LBVH, 1)):
        t0 = time.perf_counter()
        agg = soup_aggregate(ctx, p, idx, leaf, flags)
        commit_s = time.perf_counter() - t0
        st, info = agg.stats(), agg.build_info()
        s = torch.cuda.current_stream()
        for _ in range(3):
            agg.intersect_device(n_rays, d_rays.data_ptr(), d_hits.data_ptr(), s.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            agg.intersect_device(n_rays, d_rays.data_ptr(), d_hits.data_ptr(), s.cuda_stream)
        e1.record(s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        prims[name] = (d_hits.view(torch.int64)[0::4] & 0xFFFFFFFF).clone()
        print(json.dumps({"n_tris": n_tris, "builder": name, "commit_s": round(commit_s, 3), "build_usec": st["build_usec"],
                          "tree_device_ms": info["tree_device_usec"] / 1e3, "n_nodes": st["n_nodes"], "max_depth": st["max_depth"],
                          "node_bytes": info["node_bytes"], "Mrays_per_s": round(n_rays / ms / 1e3, 1)}), flush=True)
        del agg
    print("same primitive on every ray:", all(bool((prims["host SAH"] == v).all()) for v in prims.values()))
