"""CPU-side checks of the drop-in boundary: librrt_sm100.so loads, exports every symbol
include/rrt.h declares, and refuses to work without a GPU (no compute call is made here)."""
import ctypes as C
import subprocess

import pytest

from rs_ray_toy_b200 import capi
from rs_ray_toy_b200.build import build_library


@pytest.fixture(scope="module")
def library():
    build_library()
    return capi.lib()


def test_every_declared_symbol_is_exported(library):
    names = capi.declared_symbols()
    assert len(names) >= 19
    missing = [n for n in names if not hasattr(library, n)]
    assert not missing, missing


def test_record_sizes_match_header():
    from rs_ray_toy_b200.aggregate import HIT_DTYPE, RAY_DTYPE
    assert RAY_DTYPE.itemsize == 64 and HIT_DTYPE.itemsize == 32
    assert HIT_DTYPE.fields["t"][1] == 8 and HIT_DTYPE.fields["u"][1] == 16


def test_no_cpu_fallback(library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is for GPU-less hosts")
    h = C.c_void_p()
    rc = library.rrt_create(0, C.byref(h))
    assert rc != capi.RRT_OK and not h.value
    assert library.rrt_last_error()  # says why


def test_library_has_only_sm100_code():
    r = subprocess.run(["cuobjdump", "-lelf", str(capi.LIB_PATH)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = {line.split(".")[-2] for line in r.stdout.splitlines() if ".cubin" in line}
    assert archs == {"sm_100a"}, archs


def test_product_does_not_touch_the_oracle():
    # the oracle is test infrastructure: nothing under the package may import / load it
    from pathlib import Path
    pkg = Path(capi.__file__).resolve().parent
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.h*")):
        text = f.read_text(errors="replace")
        assert "liboracle" not in text and "oracle_lib" not in text and "orc_" not in text, f
