"""Experiment driver: builds variants of librrt_sm100.so with different -DRRT_<knob>=<v> settings (here, on the
build machine: `--build-only`) and times the wavefront renderer on config 4 at half resolution for each (GPU box).
Not part of the product or the tests.

    python tools/sweep_render.py --build-only "" "SHADE_MINBLOCKS=3" "GEN_MINBLOCKS=5,SHADE_MINBLOCKS=4"
    gpurun -- python tools/sweep_render.py "" "SHADE_MINBLOCKS=3" ...
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rs_ray_toy_b200.build import build_library  # noqa: E402

out_dir = ROOT / "rs_ray_toy_b200" / "variants"
out_dir.mkdir(exist_ok=True)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
build_only = "--build-only" in sys.argv
config = os.environ.get("SWEEP_CONFIG", "c4")
scale = os.environ.get("SWEEP_SCALE", "0.5")
for spec in args:
    defines = [f"RRT_{d}" for d in spec.split(",") if d]
    tag = "base" if not defines else "_".join(d.replace("=", "") for d in defines)
    lib = out_dir / f"librrt_{tag}.so"
    if not lib.exists() or build_only:
        build_library(force=True, defines=defines, out=lib)
    if build_only:
        regs = [l.split("Used")[1].split(",")[0].strip() for l in (lib.parent / (lib.stem + "_ptxas.txt")).read_text().splitlines() if "registers" in l]
        print(tag, "built", regs[-2:], flush=True)
        continue
    env = dict(os.environ, RRT_LIB=str(lib))
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "bench_render.py"), "--config", config, "--scale", scale, "--reps", "3"],
                       env=env, capture_output=True, text=True)
    try:
        js = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
        best = max(j["Msamples_per_s"] for j in js[1:])
        print(f"{spec or 'base':45s} Msamples/s={best:8.2f}  render_s={min(j['render_s'] for j in js[1:]):.4f}", flush=True)
    except Exception:
        print(spec, "FAILED", r.stdout[-500:], r.stderr[-1500:], flush=True)
