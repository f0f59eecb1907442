"""ctypes wrapper around oracle/build/liboracle.so — the CPU restatement of the reference.

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs only.  Never imported by the rs_ray_toy_b200 package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = ROOT / "oracle" / "build" / "liboracle.so"

TIER_L = 0x00  # literal reference behaviour, every Appendix-A quirk kept
TIER_F = 0x1FF  # Q1,Q2,Q3,Q4,Q5b,Q5c,Q6,Q8,Q9 switched to the evident intent

_lib = None


def build(force: bool = False) -> Path:
    srcs = list((ROOT / "oracle").glob("*.cpp")) + list((ROOT / "oracle").glob("*.hpp"))
    stale = (not LIB_PATH.exists()) or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists() or os.environ.get("RRT_ORACLE_REBUILD"):
        build()
    else:
        try:
            build()
        except Exception:
            pass  # no compiler on this box: use the prebuilt file
    L = C.CDLL(str(LIB_PATH))
    vp, u32, i32, u64, dbl = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64, C.c_double
    pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.orc_last_error.restype = C.c_char_p
    L.orc_scene_new.restype = vp
    L.orc_scene_new.argtypes = [u32]
    L.orc_scene_free.argtypes = [vp]
    L.orc_add_mesh.restype = i32
    L.orc_add_mesh.argtypes = [vp, u32, vp, u32, vp, u32, vp, vp, u32, vp, vp]
    L.orc_add_sphere.restype = i32
    L.orc_add_sphere.argtypes = [vp, pd, pd, dbl, dbl, dbl, dbl]
    L.orc_add_geo_triangles.restype = i32
    L.orc_add_geo_triangles.argtypes = [vp, i32, i32]
    L.orc_add_geo_sphere.restype = i32
    L.orc_add_geo_sphere.argtypes = [vp, i32, i32]
    L.orc_add_xform.restype = i32
    L.orc_add_xform.argtypes = [vp, pd, pd]
    L.orc_add_prims.argtypes = [vp, i32, i32, i32]
    L.orc_num_prims.restype = u32
    L.orc_num_prims.argtypes = [vp]
    L.orc_build.restype = i32
    L.orc_build.argtypes = [vp, u32]
    L.orc_num_nodes.restype = u32
    L.orc_num_nodes.argtypes = [vp]
    L.orc_num_ordered.restype = u32
    L.orc_num_ordered.argtypes = [vp]
    L.orc_get_nodes.argtypes = [vp, vp, vp, vp]
    L.orc_world_bound.argtypes = [vp, pd]
    L.orc_intersect.restype = i32
    L.orc_intersect.argtypes = [vp, u64, vp, vp, vp, vp, vp, vp, i32]
    L.orc_intersect_p.restype = i32
    L.orc_intersect_p.argtypes = [vp, u64, vp, vp, vp, i32]
    L.orc_brute_force.restype = i32
    L.orc_brute_force.argtypes = [vp, u64, vp, vp, vp, i32]
    L.orc_kat_vec3.argtypes = [pd, pd, dbl, pd]
    L.orc_kat_bounds.argtypes = [pd, pd, pd, pd]
    L.orc_kat_left_shift3.restype = u32
    L.orc_kat_left_shift3.argtypes = [u32]
    L.orc_kat_morton.restype = u32
    L.orc_kat_morton.argtypes = [dbl, dbl, dbl]
    L.orc_kat_radix_sort.argtypes = [u32, vp, vp]
    L.orc_make_to_world.argtypes = [pd, pd, dbl, pd, pd, pd]
    L.orc_m44_inverse.argtypes = [pd, pd]
    L.orc_normalize.argtypes = [pd, pd]
    L.orc_prim_intersect_p.restype = i32
    L.orc_prim_intersect_p.argtypes = [vp, u32, pd]
    L.orc_hardware_threads.restype = i32
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def hardware_threads() -> int:
    return max(1, int(lib().orc_hardware_threads()))


def make_to_world(world_pos=(0, 0, 0), axis=(0, 0, 0), angle=0.0, scale=(1, 1, 1)):
    """`make_to_world`, src/renderprocess.rs:242-252 -> (m[4,4], m_inv[4,4])."""
    m = np.zeros(16)
    inv = np.zeros(16)
    lib().orc_make_to_world(np.asarray(world_pos, np.float64), np.asarray(axis, np.float64), float(angle),
                            np.asarray(scale, np.float64), m, inv)
    return m.reshape(4, 4), inv.reshape(4, 4)


def normalize(v):
    out = np.zeros(3)
    lib().orc_normalize(np.ascontiguousarray(v, dtype=np.float64), out)
    return out


class OracleScene:
    """Mirror of what `make_aggregate` (src/renderprocess.rs:1178-1304) assembles."""

    def __init__(self, tier: int = TIER_F):
        self.L = lib()
        self.h = self.L.orc_scene_new(tier)
        self.tier = tier
        self._keep = []

    def __del__(self):
        try:
            if self.h:
                self.L.orc_scene_free(self.h)
                self.h = None
        except Exception:
            pass

    def add_mesh(self, p, vi, n=None, ni=None, uv=None, uvi=None) -> int:
        p = np.ascontiguousarray(p, dtype=np.float64)
        vi = np.ascontiguousarray(vi, dtype=np.uint32)
        n = None if n is None else np.ascontiguousarray(n, dtype=np.float64)
        ni = None if ni is None else np.ascontiguousarray(ni, dtype=np.uint32)
        uv = None if uv is None else np.ascontiguousarray(uv, dtype=np.float64)
        uvi = None if uvi is None else np.ascontiguousarray(uvi, dtype=np.uint32)
        return self.L.orc_add_mesh(self.h, p.shape[0], _ptr(p), vi.shape[0], _ptr(vi),
                                   0 if n is None else n.shape[0], _ptr(n), _ptr(ni),
                                   0 if uv is None else uv.shape[0], _ptr(uv), _ptr(uvi))

    def add_sphere(self, m=None, inv=None, radius=1.0, z_min=None, z_max=None, phi_max=360.0) -> int:
        m = np.eye(4) if m is None else np.ascontiguousarray(m, dtype=np.float64)
        inv = np.eye(4) if inv is None else np.ascontiguousarray(inv, dtype=np.float64)
        z_min = -radius if z_min is None else z_min
        z_max = radius if z_max is None else z_max
        return self.L.orc_add_sphere(self.h, m.reshape(16), inv.reshape(16), radius, z_min, z_max, phi_max)

    def add_geo_triangles(self, mesh: int, material: int = 0) -> int:
        return self.L.orc_add_geo_triangles(self.h, mesh, material)

    def add_geo_sphere(self, sphere: int, material: int = 0) -> int:
        return self.L.orc_add_geo_sphere(self.h, sphere, material)

    def add_xform(self, m, inv) -> int:
        return self.L.orc_add_xform(self.h, np.ascontiguousarray(m, np.float64).reshape(16),
                                    np.ascontiguousarray(inv, np.float64).reshape(16))

    def add_prims(self, first_geo: int, count: int, xf: int = -1):
        self.L.orc_add_prims(self.h, first_geo, count, xf)

    @property
    def num_prims(self) -> int:
        return self.L.orc_num_prims(self.h)

    def build(self, max_prims_in_node: int = 4):
        if self.L.orc_build(self.h, max_prims_in_node) != 0:
            raise RuntimeError(self.L.orc_last_error().decode())

    def nodes(self):
        n = self.L.orc_num_nodes(self.h)
        no = self.L.orc_num_ordered(self.h)
        bounds = np.zeros((n, 6))
        meta = np.zeros((n, 3), dtype=np.uint32)
        ordered = np.zeros(no, dtype=np.uint32)
        self.L.orc_get_nodes(self.h, _ptr(bounds), _ptr(meta), _ptr(ordered))
        return bounds, meta, ordered

    def world_bound(self):
        out = np.zeros(6)
        self.L.orc_world_bound(self.h, out)
        return out

    def intersect(self, rays, want_geom=False, nthreads=None):
        rays = np.ascontiguousarray(rays, dtype=np.float64)
        n = rays.shape[0]
        prim = np.empty(n, dtype=np.int32)
        t = np.empty(n)
        uv = np.empty((n, 2))
        geom = np.empty((n, 9)) if want_geom else None
        stats = np.zeros(5, dtype=np.uint64)
        nthreads = hardware_threads() if nthreads is None else nthreads
        if self.L.orc_intersect(self.h, n, _ptr(rays), _ptr(prim), _ptr(t), _ptr(uv), _ptr(geom), _ptr(stats),
                                nthreads) != 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        out = {"prim": prim, "t": t, "uv": uv, "stats": stats}
        if want_geom:
            out["geom"] = geom
        return out

    def intersect_p(self, rays, nthreads=None):
        rays = np.ascontiguousarray(rays, dtype=np.float64)
        n = rays.shape[0]
        occ = np.empty(n, dtype=np.uint8)
        stats = np.zeros(5, dtype=np.uint64)
        nthreads = hardware_threads() if nthreads is None else nthreads
        if self.L.orc_intersect_p(self.h, n, _ptr(rays), _ptr(occ), _ptr(stats), nthreads) != 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return occ, stats

    def brute_force(self, rays, nthreads=None):
        rays = np.ascontiguousarray(rays, dtype=np.float64)
        n = rays.shape[0]
        prim = np.empty(n, dtype=np.int32)
        t = np.empty(n)
        nthreads = hardware_threads() if nthreads is None else nthreads
        self.L.orc_brute_force(self.h, n, _ptr(rays), _ptr(prim), _ptr(t), nthreads)
        return prim, t

    def prim_intersect_p(self, prim: int, ray7) -> bool:
        return bool(self.L.orc_prim_intersect_p(self.h, prim, np.ascontiguousarray(ray7, dtype=np.float64)))


def soup_scene(p, idx, tier=TIER_F, max_prims=4) -> OracleScene:
    s = OracleScene(tier)
    m = s.add_mesh(p, idx)
    g0 = s.add_geo_triangles(m, 0)
    s.add_prims(g0, idx.shape[0], -1)
    s.build(max_prims)
    return s
