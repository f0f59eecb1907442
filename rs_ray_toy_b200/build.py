"""Builds librrt_sm100.so in-tree with nvcc for sm_100a (no other architecture, no JIT cache).

`python -m rs_ray_toy_b200.build` or `build_library()`; `__graft_entry__.build()` calls this.
The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "librrt_sm100.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-shared",
    "-Xcompiler", "-fPIC,-O3,-pthread,-Wall,-Wno-unused-function,-ffp-contract=off",
    # no implicit FMA contraction anywhere: the f64 render path must round like the reference's plain Rust
    # arithmetic; the fp32 box tests ask for their FFMAs explicitly (fmaf)
    "--fmad=false",
    "-Xptxas", "-v",
]


def sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*")) + [PKG.parent / "include" / "rrt.h", PKG.parent / "include" / "rrt_test.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, defines=(), out: Path | None = None) -> Path:
    """`defines` / `out` build an experiment variant (tools/sweep.py); the product is the default call."""
    global LIB
    if out is None and not force and not is_stale():
        return LIB
    if out is not None:
        saved, LIB = LIB, Path(out)
        try:
            return _build(verbose, defines)
        finally:
            LIB = saved
    return _build(verbose, defines)


def _build(verbose, defines):
    """One nvcc -c per source, in parallel, objects cached under csrc/../build/<variant>/ (git-ignored); an object is
    rebuilt when its source, any header in csrc/ or include/, or this file is newer.  Then one link."""
    import hashlib
    from concurrent.futures import ThreadPoolExecutor

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        if LIB.exists():
            return LIB  # GPU box without a toolchain: use the prebuilt library that travelled with the repo
        raise RuntimeError("nvcc not found and no prebuilt librrt_sm100.so")
    variant = hashlib.sha1(" ".join(sorted(defines)).encode()).hexdigest()[:10] if defines else "product"
    objdir = PKG / "build" / variant
    objdir.mkdir(parents=True, exist_ok=True)
    headers = [p for p in CSRC.glob("*") if p.suffix in (".h", ".hpp", ".cuh")] + [PKG.parent / "include" / "rrt.h",
                                                                                  PKG.parent / "include" / "rrt_test.h", Path(__file__)]
    hdr_time = max(h.stat().st_mtime for h in headers if h.exists())
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src: Path):
        obj = objdir / (src.name + ".o")
        if obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_time):
            log = objdir / (src.name + ".log")
            return obj, 0, log.read_text() if log.exists() else ""
        cmd = [nvcc, *compile_flags, *[f"-D{d}" for d in defines], "-I", str(PKG.parent / "include"), "-c", "-o", str(obj), str(src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 and obj.exists():
            obj.unlink()
        (objdir / (src.name + ".log")).write_text(r.stdout + r.stderr)
        return obj, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        results = list(ex.map(compile_one, sources()))
    log = "".join(out for _, _, out in results)
    if any(rc != 0 for _, rc, _ in results):
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    tmp = LIB.with_suffix(".so.tmp%d" % os.getpid())
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(tmp), *[str(o) for o, _, _ in results], "-lcudart", "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc link failed")
    if verbose:
        sys.stderr.write(log)
    (LIB.parent / (LIB.stem + "_ptxas.txt")).write_text(log)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
