"""Writes tests/golden/*.npz from the CPU oracle (Tier F).  The Rust reference cannot be built
in this image (nightly toolchain + un-vendored crates, no cargo), so these vectors come from
the line-by-line restatement under oracle/, not from the reference binary: "parity unpinned by
the reference" (SURVEY.md §8c); they pin the oracle against drift and let the GPU box check
parity with no oracle in the loop.   Run:  python tests/golden/make_golden.py"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import golden_cases  # noqa: E402

for name in golden_cases.CASES:
    s, rays = golden_cases.build_oracle(name)
    r = s.intersect(rays)
    bp, bt = s.brute_force(rays)
    assert (bp == r["prim"]).all() and (bt == r["t"]).all(), name
    occ, _ = s.intersect_p(golden_cases.shadow_rays(name, rays))
    np.savez_compressed(HERE / f"{name}.npz", rays=rays, prim=r["prim"], t=r["t"], uv=r["uv"], occluded=occ,
                        stats=r["stats"])
    print(name, "hits", int((r["prim"] >= 0).sum()), "occluded", int(occ.sum()), "stats", r["stats"])
