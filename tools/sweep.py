"""Experiment driver (GPU box): builds variants of librrt_sm100.so with different -D knobs / env
knobs and times the config-3 closest-hit kernel for each.  Not part of the product or the tests.

    python tools/sweep.py "MINBLOCKS=4,REFILL=8" "MINBLOCKS=5,REFILL=4;RRT_SORT_MODE=0" ...
A variant is `DEFINES[;ENV]`: comma-separated -DRRT_<k>=<v> and comma-separated environment pairs.
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
rays = os.environ.get("SWEEP_RAYS", str(1 << 24))
build_only = "--build-only" in sys.argv   # on the build machine: compile the variants, the GPU box only times them
for spec in [a for a in sys.argv[1:] if not a.startswith("--")]:
    defs, _, envs = spec.partition(";")
    defines = [f"RRT_{d}" for d in defs.split(",") if d]
    tag = "base" if not defines else "_".join(d.replace("=", "") for d in defines)
    lib = out_dir / f"librrt_{tag}.so" if defines else ROOT / "rs_ray_toy_b200" / "librrt_sm100.so"
    if not lib.exists():
        build_library(force=True, defines=defines, out=lib)
    regs = [l for l in (lib.parent / (lib.stem + "_ptxas.txt")).read_text().splitlines() if "registers" in l][:2]
    if build_only:
        print(tag, "built", flush=True)
        continue
    env = dict(os.environ, RRT_LIB=str(lib))
    for kv in envs.split(","):
        if kv:
            k, v = kv.split("=")
            env[k] = v
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu-baseline", "--path-config", "none",
                        "--rays", rays], env=env, capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{spec:50s} value={j['value']:9.1f} Mrays/s  e2e={j['e2e']['value']:8.1f}  kernel_ms={j['roofline']['kernel_ms']:.2f} "
              f"launches={j['gpu_launches']} regs={[x.split('Used')[1].split(',')[0].strip() for x in regs]}", flush=True)
    except Exception:
        print(spec, "FAILED", r.stdout[-500:], r.stderr[-1500:], flush=True)
