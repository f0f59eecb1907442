"""GPU box: does a cropped warm-up make the next full frame of config 5 slow?  (bench.py timed 3.8 s where
tools/bench_render.py times 2.9 s.)  Prints the wall time of: crop, frame, frame, frame."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rs_ray_toy_b200 import synth  # noqa: E402
from rs_ray_toy_b200.aggregate import Context  # noqa: E402

ctx = Context(0)
agg, r = synth.scene_c5_api(ctx)
for label, kw in [("crop", {"crop": (1680, 945, 2160, 1215)}), ("frame", {}), ("frame", {}), ("crop", {"crop": (1680, 945, 2160, 1215)}), ("frame", {})]:
    r.clear()
    t0 = time.perf_counter()
    r.run(**kw)
    dt = time.perf_counter() - t0
    s = r.stats()
    print(label, round(dt, 3), "s", s["samples"], "samples", s["launches"], "launches", s["chunks"], "chunks", flush=True)
