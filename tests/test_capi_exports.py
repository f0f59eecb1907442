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


def test_rust_bindings_name_exported_symbols_and_c_sizes(library, tmp_path):
    """rust/rrt-sys/src/lib.rs cannot be compiled here (no cargo): check what can be checked — every `pub fn rrt_*` it
    declares is exported by librrt_sm100.so and declared in include/rrt.h (never in rrt_test.h), every product entry
    point is bound, and the `size_of` assertions it carries equal the C compiler's sizeof for include/rrt.h."""
    import re
    from pathlib import Path
    root = Path(capi.__file__).resolve().parent.parent
    rs = (root / "rust" / "rrt-sys" / "src" / "lib.rs").read_text()
    bound = set(re.findall(r"pub fn (rrt_[a-z0-9_]+)\s*\(", rs))
    header = re.sub(r"/\*.*?\*/", "", (root / "include" / "rrt.h").read_text(), flags=re.S)
    declared = set(re.findall(r"\b(rrt_[a-z0-9_]+)\s*\(", header))
    assert bound == declared, (sorted(bound - declared), sorted(declared - bound))
    assert all(hasattr(library, n) for n in bound)
    sizes = dict((m[0], int(m[1])) for m in re.findall(r"size_of::<(rrt_[a-z_]+)>\(\) == (\d+)", rs))
    assert set(sizes) == {"rrt_ray", "rrt_hit", "rrt_material", "rrt_texture", "rrt_light", "rrt_render_desc"}
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "rrt.h"\nint main(void) {\n' +
                   "".join(f'  printf("{n} %zu\\n", sizeof({n}));\n' for n in sizes) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    r = subprocess.run(["gcc", "-I", str(root / "include"), "-o", str(exe), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    assert {k: int(v) for k, v in out.items()} == sizes
    # and the ctypes mirrors the tests use agree too
    from rs_ray_toy_b200 import render as R
    assert (C.sizeof(R.Material), C.sizeof(R.Texture), C.sizeof(R.Light), C.sizeof(R.RenderDesc)) == \
        (sizes["rrt_material"], sizes["rrt_texture"], sizes["rrt_light"], sizes["rrt_render_desc"])
