"""Host logic of the Tier-F commit, without a GPU: the binned-SAH builder (csrc/bvh_sah.cpp) and the packer's slot / record
plan (csrc/bvh_pack_plan.hpp), compiled with g++ into a small checker (tests/cpp/pack_plan_check.cpp).  The checker builds
a tree over a random soup with duplicated boxes (multi-primitive leaves, median splits), validates it (every primitive in
exactly one leaf, boxes nested, subtree totals) and compares the parallel plan with the one-thread depth-first walk the
packer used to run: same slot for every interior node, same first record for every leaf."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "rs_ray_toy_b200" / "csrc"


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    gxx = shutil.which("g++") or "/opt/gcc/bin/g++"
    if not Path(gxx).exists():
        pytest.skip("no g++ on this box")
    exe = tmp_path_factory.mktemp("sah") / "pack_plan_check"
    subprocess.run([gxx, "-O2", "-std=c++17", "-pthread", "-ffp-contract=off", "-I", str(CSRC), "-I", str(ROOT / "include"),
                    "-o", str(exe), str(ROOT / "tests" / "cpp" / "pack_plan_check.cpp"), str(CSRC / "bvh_sah.cpp")], check=True)
    return exe


@pytest.mark.parametrize("n", [2, 3, 17, 5000, 300000])
def test_sah_tree_and_pack_plan(checker, n):
    r = subprocess.run([str(checker), str(n)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "tree VALID" in r.stdout and "IDENTICAL" in r.stdout, r.stdout


def test_sah_tree_does_not_depend_on_the_thread_count(checker):
    """Nodes of >= 2^18 primitives split their binning pass and their partition (a stable two-sided scatter) over threads:
    the tree — node count, depth, every slot of the plan — must be the one a single thread builds."""
    import os
    outs = []
    for threads in ("1", "3", "16"):
        r = subprocess.run([str(checker), "300000"], capture_output=True, text=True, timeout=300, env=dict(os.environ, T=threads))
        assert r.returncode == 0 and "tree VALID" in r.stdout and "IDENTICAL" in r.stdout, r.stdout + r.stderr
        outs.append([ln for ln in r.stdout.splitlines() if ln.startswith("tree") or ln.startswith("IDENTICAL")])
    assert outs[0] == outs[1] == outs[2], outs
