"""Host-side mirror of the reference's render entry points over the C ABI.

`Render.load(ctx, "scene.json")` is `deploy_render` (src/renderprocess.rs:92-105) up to the
`Box<dyn Integrator>`; `.run()` is `Integrator::render` (src/integrator/mod.rs:21-23,48-139) for
all tiles or for one rank's share; `.film()` is `Film::write_image` (src/film.rs:323-366) up to
the float image.  Everything is computed by librrt_sm100.so.
"""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from . import capi
from .aggregate import Context, GpuAggregate


class RenderDesc(C.Structure):
    """rrt_render_desc (include/rrt.h)."""
    _fields_ = [
        ("xres", C.c_int64), ("yres", C.c_int64),
        ("diagonal_mm", C.c_double), ("scale", C.c_double), ("max_sample_luminance", C.c_double),
        ("filter_kind", C.c_uint32), ("pad0", C.c_uint32),
        ("filter_radius", C.c_double * 2), ("filter_alpha", C.c_double),
        ("cam_pos", C.c_double * 3), ("cam_look", C.c_double * 3), ("cam_up", C.c_double * 3),
        ("shutter_open", C.c_double), ("shutter_close", C.c_double), ("aperture_diameter", C.c_double),
        ("focus_distance", C.c_double),
        ("simple_weighting", C.c_uint32), ("n_lens_values", C.c_uint32),
        ("lens_data", C.POINTER(C.c_double)),
        ("nsamp", C.c_uint64),
        ("sample_at_center", C.c_uint32), ("light_strategy", C.c_uint32),
        ("seed", C.c_uint64),
        ("integrator_kind", C.c_uint32), ("max_depth", C.c_uint32),
        ("rr_threshold", C.c_double),
        ("sampler_kind", C.c_uint32), ("strat_xsamp", C.c_uint32), ("strat_ysamp", C.c_uint32),
        ("strat_dimension", C.c_uint32), ("strat_jitter", C.c_uint32), ("pad1", C.c_uint32),
    ]


class Material(C.Structure):
    """rrt_material."""
    _fields_ = [("kind", C.c_uint32), ("remap_roughness", C.c_uint32),
                ("kd", C.c_double * 3), ("ks", C.c_double * 3), ("kr", C.c_double * 3), ("kt", C.c_double * 3),
                ("metal_eta", C.c_double * 3), ("metal_k", C.c_double * 3),
                ("sigma", C.c_double), ("roughness", C.c_double), ("u_roughness", C.c_double),
                ("v_roughness", C.c_double), ("eta", C.c_double),
                ("metallic", C.c_double), ("specular_tint", C.c_double), ("anisotropic", C.c_double), ("sheen", C.c_double),
                ("sheen_tint", C.c_double), ("clearcoat", C.c_double), ("clearcoat_gloss", C.c_double),
                ("spec_trans", C.c_double), ("flatness", C.c_double), ("diff_trans", C.c_double),
                ("scatter_distance", C.c_double * 3), ("thin", C.c_uint32), ("pad", C.c_uint32)]


class Texture(C.Structure):
    """rrt_texture."""
    _fields_ = [("kind", C.c_uint32), ("mapping", C.c_uint32), ("t1", C.c_int32), ("t2", C.c_int32), ("amount", C.c_int32),
                ("aa", C.c_uint32), ("v", (C.c_double * 3) * 4), ("map", C.c_double * 8), ("world_to_texture", C.c_double * 16)]


MAX_TEXTURES = 32
MATERIAL_SLOTS = 23
(SLOT_KD, SLOT_KS, SLOT_KR, SLOT_KT, SLOT_METAL_ETA, SLOT_METAL_K, SLOT_SIGMA, SLOT_ROUGHNESS, SLOT_U_ROUGHNESS,
 SLOT_V_ROUGHNESS, SLOT_ETA, SLOT_BUMP_MAP, SLOT_METALLIC, SLOT_SPECULAR_TINT, SLOT_ANISOTROPIC, SLOT_SHEEN, SLOT_SHEEN_TINT,
 SLOT_CLEARCOAT, SLOT_CLEARCOAT_GLOSS, SLOT_SPEC_TRANS, SLOT_FLATNESS, SLOT_DIFF_TRANS, SLOT_SCATTER_DISTANCE) = range(23)
TEX_CONSTANT, TEX_BILERP, TEX_SCALE, TEX_MIX, TEX_CHECKER2D, TEX_CHECKER3D, TEX_UV, TEX_WINDY, TEX_WRINKLED = range(9)
TEXMAP_UV, TEXMAP_PLANAR, TEXMAP_SPHERICAL, TEXMAP_CYLINDRICAL = range(4)


def texture(kind, v=(), mapping=TEXMAP_UV, map8=(1, 1, 0, 0, 0, 0, 0, 0), t1=-1, t2=-1, amount=-1, world_to_texture=None,
            aa=0) -> Texture:
    """One row of the texture table; `v` = up to four values (floats or RGB triples); `aa` = 1 for a closed-form
    checkerboard."""
    t = Texture(kind=kind, mapping=mapping, t1=t1, t2=t2, amount=amount, aa=aa)
    for k, val in enumerate(v):
        t.v[k][:] = [float(val), 0.0, 0.0] if np.isscalar(val) else [float(x) for x in val]
    t.map[:] = [float(x) for x in map8]
    t.world_to_texture[:] = (np.eye(4) if world_to_texture is None else np.asarray(world_to_texture, dtype=np.float64)).reshape(16).tolist()
    return t


def texture_host_probe(textures, uv, p, diff=None) -> np.ndarray:
    """rrt_texture_host_probe: every texture of the table evaluated at (uv, p) by the product's evaluator on the host.
    `diff` = dpdx[3] dpdy[3] dudx dvdx dudy dvdy (None: no differentials)."""
    L = lib()
    arr = (Texture * max(1, len(textures)))(*textures)
    out = np.zeros((len(textures), 3))
    uv2 = (C.c_double * 2)(*[float(x) for x in uv])
    p3 = (C.c_double * 3)(*[float(x) for x in p])
    d = None if diff is None else (C.c_double * 10)(*[float(x) for x in diff])
    capi.check(L.rrt_texture_host_probe(len(textures), C.cast(arr, C.c_void_p), uv2, p3, d, out.ctypes.data))
    return out


def differentials_host_probe(p, n, dpdu, dpdv, rx_o, rx_d, ry_o, ry_d) -> np.ndarray:
    """rrt_differentials_host_probe: SurfaceInteraction::compute_differentials with the product's code on the host
    -> dpdx[3] dpdy[3] dudx dvdx dudy dvdy."""
    L = lib()
    a = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float64).reshape(3) for x in (p, n, dpdu, dpdv, rx_o, rx_d, ry_o, ry_d)]))
    out = np.zeros(10)
    capi.check(L.rrt_differentials_host_probe(a.ctypes.data, out.ctypes.data))
    return out


def json_texture_probe(path, overrides=None, max_materials=256):
    """rrt_scene_json_texture_probe -> (list of Texture, list of Material, slots[n_materials, 11])."""
    L = lib()
    tex = (Texture * MAX_TEXTURES)()
    mats = (Material * max_materials)()
    slots = np.full((max_materials, MATERIAL_SLOTS), -1, dtype=np.int32)
    nt, nm = C.c_uint32(), C.c_uint32()
    ov = json.dumps(overrides).encode() if overrides else None
    capi.check(L.rrt_scene_json_texture_probe(str(path).encode(), ov, C.byref(nt), C.cast(tex, C.c_void_p), max_materials,
                                              C.byref(nm), C.cast(mats, C.c_void_p), slots.ctypes.data))
    n = min(nm.value, max_materials)
    return [tex[i] for i in range(nt.value)], [mats[i] for i in range(n)], slots[:n].copy()


class Light(C.Structure):
    """rrt_light."""
    _fields_ = [("kind", C.c_uint32), ("shape_kind", C.c_uint32), ("intensity", C.c_double * 3), ("dir", C.c_double * 3),
                ("to_world", C.c_double * 16), ("shape_to_world", C.c_double * 16), ("shape_to_world_inv", C.c_double * 16),
                ("radius", C.c_double), ("z_min", C.c_double), ("z_max", C.c_double), ("phi_max_deg", C.c_double),
                ("tri_p", C.c_double * 9), ("tri_n", C.c_double * 9), ("tri_has_n", C.c_uint32), ("env_image", C.c_uint32)]


COPPER_N = (0.19998972096819712, 0.922085788777433, 1.0998762520488314)
COPPER_K = (3.9046381767086675, 2.4476332238684626, 2.1376510366555137)


def matte(kd=(0.5, 0.5, 0.5), sigma=0.0) -> Material:
    return Material(kind=0, kd=tuple(kd), sigma=sigma, u_roughness=-1, v_roughness=-1)


def plastic(kd=(0.25,) * 3, ks=(0.25,) * 3, roughness=0.1, remap=False) -> Material:
    return Material(kind=1, remap_roughness=int(remap), kd=tuple(kd), ks=tuple(ks), roughness=roughness, u_roughness=-1,
                    v_roughness=-1)


def metal(eta=COPPER_N, k=COPPER_K, roughness=0.01, u_roughness=-1.0, v_roughness=-1.0, remap=False) -> Material:
    return Material(kind=2, remap_roughness=int(remap), metal_eta=tuple(eta), metal_k=tuple(k), roughness=roughness,
                    u_roughness=u_roughness, v_roughness=v_roughness)


def mirror(kr=(0.9,) * 3) -> Material:
    return Material(kind=3, kr=tuple(kr), u_roughness=-1, v_roughness=-1)


def glass(kr=(1.0,) * 3, kt=(1.0,) * 3, eta=1.5) -> Material:
    return Material(kind=4, kr=tuple(kr), kt=tuple(kt), eta=eta, u_roughness=0.0, v_roughness=0.0)


def translucent(kd=(0.25,) * 3, ks=(0.25,) * 3, roughness=0.1, reflect=(0.25,) * 3, transmit=(0.25,) * 3, remap=False) -> Material:
    """TranslucentMaterial (material/translucent.rs; defaults renderprocess.rs:695-706)."""
    return Material(kind=5, remap_roughness=int(remap), kd=tuple(kd), ks=tuple(ks), kr=tuple(reflect), kt=tuple(transmit),
                    roughness=roughness, u_roughness=-1, v_roughness=-1)


def disney(color=(0.5,) * 3, metallic=0.0, eta=1.5, roughness=0.5, specular_tint=0.0, anisotropic=0.0, sheen=0.0, sheen_tint=0.5,
           clearcoat=0.0, clearcoat_gloss=1.0, spec_trans=0.0, scatter_distance=(0.0,) * 3, thin=False, flatness=0.0,
           diff_trans=1.0) -> Material:
    """DisneyMaterial (material/disney.rs; defaults renderprocess.rs:810-836)."""
    return Material(kind=6, kd=tuple(color), metallic=metallic, eta=eta, roughness=roughness, specular_tint=specular_tint,
                    anisotropic=anisotropic, sheen=sheen, sheen_tint=sheen_tint, clearcoat=clearcoat,
                    clearcoat_gloss=clearcoat_gloss, spec_trans=spec_trans, scatter_distance=tuple(scatter_distance),
                    thin=int(thin), flatness=flatness, diff_trans=diff_trans, u_roughness=-1, v_roughness=-1)


def debug_material() -> Material:
    """DebugMaterial (material/debug_material.rs)."""
    return Material(kind=7, u_roughness=-1, v_roughness=-1)


def point_light(intensity=(1.0, 1.0, 1.0)) -> Light:
    l = Light(kind=0, intensity=tuple(intensity))
    l.to_world[:] = np.eye(4).reshape(16).tolist()
    return l


def distant_light(l=(1.0, 1.0, 1.0), frm=(0.0, 0.0, 0.0), to=(0.0, 0.0, 1.0), to_world=None) -> Light:
    lt = Light(kind=1, intensity=tuple(l), dir=tuple(np.subtract(frm, to).tolist()))
    lt.to_world[:] = (np.eye(4) if to_world is None else np.asarray(to_world)).reshape(16).tolist()
    return lt


def infinite_light(image_index: int, light_to_world=None, l=(1.0, 1.0, 1.0)) -> Light:
    """InfiniteAreaLight over an image added with add_image (lights/infinite.rs); `light_to_world` = (m, inverse)."""
    lt = Light(kind=3, intensity=tuple(l), env_image=image_index)
    m, inv = light_to_world if light_to_world is not None else (np.eye(4), np.eye(4))
    lt.to_world[:] = np.asarray(m).reshape(16).tolist()
    lt.shape_to_world_inv[:] = np.asarray(inv).reshape(16).tolist()
    return lt


def add_image(agg: GpuAggregate, rgb8: np.ndarray) -> int:
    """rrt_scene_add_image: 8-bit RGB rows (top first) for RRT_TEX_IMAGE textures / infinite lights -> image index."""
    L = lib()
    a = np.ascontiguousarray(rgb8, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 3
    idx = C.c_uint32()
    capi.check(L.rrt_scene_add_image(agg.h, a.shape[1], a.shape[0], a.ctypes.data, C.byref(idx)))
    return idx.value


def set_infinite_lights(agg: GpuAggregate, lights) -> None:
    """rrt_scene_set_infinite_lights: Scene::infinite_lights (escaped rays of the Path integrator)."""
    L = lib()
    lts = (Light * max(1, len(lights)))(*lights)
    capi.check(L.rrt_scene_set_infinite_lights(agg.h, len(lights), C.cast(lts, C.c_void_p)))


def area_light_sphere(lemit=(1.0, 1.0, 1.0), radius=1.0, obj_to_world=None) -> Light:
    """DiffuseAreaLight over make_sphere's Sphere (renderprocess.rs:999-1017, 1097-1106)."""
    l = Light()
    l.kind, l.shape_kind = 2, 0
    l.intensity[:] = lemit
    m, inv = obj_to_world if obj_to_world is not None else (np.eye(4), np.eye(4))
    l.to_world[:] = np.eye(4).reshape(16).tolist()
    l.shape_to_world[:] = np.asarray(m, dtype=np.float64).reshape(16).tolist()
    l.shape_to_world_inv[:] = np.asarray(inv, dtype=np.float64).reshape(16).tolist()
    l.radius, l.z_min, l.z_max, l.phi_max_deg = radius, -radius, radius, 360.0
    return l


def area_light_triangle(lemit, p3x3, n3x3=None) -> Light:
    """DiffuseAreaLight over one mesh triangle (renderprocess.rs:1084-1090)."""
    l = Light()
    l.kind, l.shape_kind = 2, 1
    l.intensity[:] = lemit
    l.to_world[:] = np.eye(4).reshape(16).tolist()
    l.tri_p[:] = np.asarray(p3x3, dtype=np.float64).reshape(9).tolist()
    if n3x3 is not None:
        l.tri_n[:] = np.asarray(n3x3, dtype=np.float64).reshape(9).tolist()
        l.tri_has_n = 1
    return l


def _bind(L):
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    pvp = C.POINTER(C.c_void_p)
    L.rrt_scene_set_materials.restype = i32
    L.rrt_scene_set_materials.argtypes = [vp, u32, vp]
    L.rrt_scene_set_lights.restype = i32
    L.rrt_scene_set_lights.argtypes = [vp, u32, vp]
    L.rrt_scene_set_infinite_lights.restype = i32
    L.rrt_scene_set_infinite_lights.argtypes = [vp, u32, vp]
    L.rrt_scene_add_image.restype = i32
    L.rrt_scene_add_image.argtypes = [vp, u32, u32, vp, C.POINTER(u32)]
    L.rrt_scene_add_image_png.restype = i32
    L.rrt_scene_add_image_png.argtypes = [vp, C.c_char_p, C.POINTER(u32)]
    L.rrt_scene_set_textures.restype = i32
    L.rrt_scene_set_textures.argtypes = [vp, u32, vp]
    L.rrt_scene_set_material_textures.restype = i32
    L.rrt_scene_set_material_textures.argtypes = [vp, u32, vp]
    L.rrt_texture_host_probe.restype = i32
    L.rrt_texture_host_probe.argtypes = [u32, vp, vp, vp, vp, vp]
    L.rrt_differentials_host_probe.restype = i32
    L.rrt_differentials_host_probe.argtypes = [vp, vp]
    L.rrt_scene_json_texture_probe.restype = i32
    L.rrt_scene_json_texture_probe.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(u32), vp, u32, C.POINTER(u32), vp, vp]
    L.rrt_scene_load_json.restype = i32
    L.rrt_scene_load_json.argtypes = [vp, C.c_char_p, C.c_char_p, u64, pvp, pvp]
    L.rrt_scene_load_json_tier.restype = i32
    L.rrt_scene_load_json_tier.argtypes = [vp, C.c_char_p, C.c_char_p, u64, u32, pvp, pvp]
    L.rrt_scene_json_probe.restype = i32
    L.rrt_scene_json_probe.argtypes = [C.c_char_p, C.c_char_p, vp, C.POINTER(RenderDesc)]
    L.rrt_render_create.restype = i32
    L.rrt_render_create.argtypes = [vp, C.POINTER(RenderDesc), pvp]
    L.rrt_render_destroy.restype = None
    L.rrt_render_destroy.argtypes = [vp]
    L.rrt_render_run.restype = i32
    L.rrt_render_run.argtypes = [vp, u32, u32, vp]
    L.rrt_render_clear.restype = i32
    L.rrt_render_clear.argtypes = [vp]
    L.rrt_render_read_film.restype = i32
    L.rrt_render_read_film.argtypes = [vp, vp, vp]
    L.rrt_render_film_device.restype = i32
    L.rrt_render_film_device.argtypes = [vp, pvp, C.POINTER(u64)]
    L.rrt_render_read_rgba8.restype = i32
    L.rrt_render_read_rgba8.argtypes = [vp, vp]
    L.rrt_render_write_png.restype = i32
    L.rrt_render_write_png.argtypes = [vp, C.c_char_p]
    L.rrt_rgb_to_png.restype = i32
    L.rrt_rgb_to_png.argtypes = [vp, u32, u32, C.c_char_p, vp]
    L.rrt_render_film_copy.restype = i32
    L.rrt_render_film_copy.argtypes = [vp, vp, i32, vp]
    L.rrt_render_owned_doubles.restype = i32
    L.rrt_render_owned_doubles.argtypes = [vp, u32, u32, C.POINTER(C.c_uint64)]
    L.rrt_render_pack_owned.restype = i32
    L.rrt_render_pack_owned.argtypes = [vp, u32, u32, vp, C.c_uint64, vp]
    L.rrt_render_unpack_owned.restype = i32
    L.rrt_render_unpack_owned.argtypes = [vp, u32, u32, vp, C.c_uint64, vp]
    L.rrt_film_gather.restype = i32
    L.rrt_film_gather.argtypes = [vp, u32, u32]
    L.rrt_render_stats.restype = i32
    L.rrt_render_stats.argtypes = [vp, vp]
    L.rrt_render_hit_dump.restype = i32
    L.rrt_render_hit_dump.argtypes = [vp, i32, vp, u64, C.POINTER(u64)]
    return L


def lib():
    L = capi.lib()
    if not getattr(L, "_render_bound", False):
        _bind(L)
        L._render_bound = True
    return L


def rgb_to_png(rgb, path=None) -> np.ndarray:
    """write_image's quantisation (+ optional PNG file) for a float image [h, w, 3]; host only."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float64)
    out = np.zeros(rgb.shape[:2] + (4,), dtype=np.uint8)
    capi.check(lib().rrt_rgb_to_png(rgb.ctypes.data, rgb.shape[1], rgb.shape[0], str(path).encode() if path else None,
                                    out.ctypes.data))
    return out


def json_probe(path, overrides=None):
    """What the C++ loader reads from a scene file (host only; no GPU needed)."""
    L = lib()
    out = np.zeros(8, dtype=np.uint64)
    desc = RenderDesc()
    ov = json.dumps(overrides).encode() if overrides else None
    capi.check(L.rrt_scene_json_probe(str(path).encode(), ov, out.ctypes.data, C.byref(desc)))
    keys = ["prims", "meshes", "spheres", "instances", "materials", "lights", "max_prims_in_node", "lens_values"]
    return {k: int(v) for k, v in zip(keys, out)}, desc


class Render:
    """The GPU stand-in for `Box<dyn Integrator>` (+ its camera, film and sampler)."""

    def __init__(self, ctx: Context, scene_handle, render_handle, owns_scene: bool, xres: int, yres: int, keep=None):
        self.ctx, self.L = ctx, lib()
        self.scene_h, self.h, self.owns_scene = scene_handle, render_handle, owns_scene
        self.xres, self.yres = xres, yres
        self._keep = keep

    @classmethod
    def load(cls, ctx: Context, path, overrides=None, seed: int = 1, literal: bool = False) -> "Render":
        """`deploy_render(filepath, ..)` up to the integrator: make_scene + make_integrator.
        `literal` selects the Tier-L aggregate and integrator rules (every reference quirk kept)."""
        L = lib()
        sh, rh = C.c_void_p(), C.c_void_p()
        ov = json.dumps(overrides).encode() if overrides else None
        capi.check(L.rrt_scene_load_json_tier(ctx.h, str(path).encode(), ov, seed,
                                              capi.RRT_BUILD_LITERAL if literal else capi.RRT_BUILD_FAST,
                                              C.byref(sh), C.byref(rh)))
        _, desc = json_probe(path, overrides)
        return cls(ctx, sh, rh, True, int(desc.xres), int(desc.yres))

    @classmethod
    def create(cls, agg: GpuAggregate, materials, lights, desc: RenderDesc, lens_data, textures=None,
               material_slots=None) -> "Render":
        """make_integrator for an aggregate assembled through GpuAggregate.  `textures` = rows of the texture table,
        `material_slots[n_materials, 11]` = which texture drives each material parameter (-1: the constant)."""
        L = lib()
        mats = (Material * len(materials))(*materials)
        lts = (Light * max(1, len(lights)))(*lights)
        capi.check(L.rrt_scene_set_materials(agg.h, len(materials), C.cast(mats, C.c_void_p)))
        capi.check(L.rrt_scene_set_lights(agg.h, len(lights), C.cast(lts, C.c_void_p)))
        if textures is not None:
            tex = (Texture * max(1, len(textures)))(*textures)
            capi.check(L.rrt_scene_set_textures(agg.h, len(textures), C.cast(tex, C.c_void_p)))
        if material_slots is not None:
            sl = np.ascontiguousarray(material_slots, dtype=np.int32).reshape(len(materials), MATERIAL_SLOTS)
            capi.check(L.rrt_scene_set_material_textures(agg.h, len(materials), sl.ctypes.data))
        lens = np.ascontiguousarray(lens_data, dtype=np.float64).reshape(-1)
        desc.lens_data = lens.ctypes.data_as(C.POINTER(C.c_double))
        desc.n_lens_values = lens.shape[0]
        rh = C.c_void_p()
        capi.check(L.rrt_render_create(agg.h, C.byref(desc), C.byref(rh)))
        return cls(agg.ctx, agg.h, rh, False, int(desc.xres), int(desc.yres), keep=(agg, lens))

    def close(self):
        if getattr(self, "h", None):
            self.L.rrt_render_destroy(self.h)
            self.h = None
            if self.owns_scene and self.scene_h:
                self.L.rrt_scene_destroy(self.scene_h)
                self.scene_h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, tile_mod: int = 1, tile_rank: int = 0, crop=None):
        """`Integrator::render` for the tiles t with t % tile_mod == tile_rank."""
        c = None
        if crop is not None:
            c = (C.c_int64 * 4)(*[int(x) for x in crop])
        capi.check(self.L.rrt_render_run(self.h, tile_mod, tile_rank, C.cast(c, C.c_void_p) if c is not None else None))

    def clear(self):
        capi.check(self.L.rrt_render_clear(self.h))

    def film(self, want_raw: bool = False):
        """`Film::write_image` up to the float RGB image [yres, xres, 3] (+ raw xyz/weight)."""
        rgb = np.zeros((self.yres, self.xres, 3))
        raw = np.zeros((self.yres, self.xres, 4)) if want_raw else None
        capi.check(self.L.rrt_render_read_film(self.h, rgb.ctypes.data, raw.ctypes.data if want_raw else None))
        return (rgb, raw) if want_raw else rgb

    def rgba8(self) -> np.ndarray:
        """`write_image`'s 8-bit pixels (renderprocess.rs:1501-1530): [yres, xres, 4] uint8."""
        out = np.zeros((self.yres, self.xres, 4), dtype=np.uint8)
        capi.check(self.L.rrt_render_read_rgba8(self.h, out.ctypes.data))
        return out

    def write_png(self, path):
        capi.check(self.L.rrt_render_write_png(self.h, str(path).encode()))

    def film_device(self):
        p, n = C.c_void_p(), C.c_uint64()
        capi.check(self.L.rrt_render_film_device(self.h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def film_copy(self, d_buffer: int, to_render: bool, stream: int = 0):
        """Accumulation film -> device buffer (to_render False) or back (True); 4*xres*yres f64."""
        capi.check(self.L.rrt_render_film_copy(self.h, C.c_void_p(d_buffer), int(to_render), C.c_void_p(stream)))

    def owned_doubles(self, tile_mod: int, tile_rank: int) -> int:
        n = C.c_uint64()
        capi.check(self.L.rrt_render_owned_doubles(self.h, tile_mod, tile_rank, C.byref(n)))
        return int(n.value)

    def pack_owned(self, tile_mod: int, tile_rank: int, d_buffer: int, capacity_doubles: int, stream: int = 0):
        """The pixels of this rank's tiles -> a packed device buffer (the multi-GPU film gather's send side)."""
        capi.check(self.L.rrt_render_pack_owned(self.h, tile_mod, tile_rank, C.c_void_p(d_buffer), capacity_doubles, C.c_void_p(stream)))

    def unpack_owned(self, tile_mod: int, tile_rank: int, d_buffer: int, capacity_doubles: int, stream: int = 0):
        capi.check(self.L.rrt_render_unpack_owned(self.h, tile_mod, tile_rank, C.c_void_p(d_buffer), capacity_doubles, C.c_void_p(stream)))

    def stats(self) -> dict:
        out = np.zeros(16, dtype=np.uint64)
        capi.check(self.L.rrt_render_stats(self.h, out.ctypes.data))
        keys = ["camera_rays", "extension_rays", "shadow_rays", "bounces", "zero_weight", "samples", "launches",
                "render_usec", "setup_usec", "chunks", "f32_neighbours", "f32_unsure"]
        return {k: int(v) for k, v in zip(keys, out)}

    def enable_hit_dump(self, on: bool = True):
        capi.check(self.L.rrt_render_hit_dump(self.h, int(on), None, 0, None))

    def hit_dump(self) -> np.ndarray:
        """(pixel x, pixel y, sample, prim id | -1 miss | -2 zero weight, t, weight) per camera sample."""
        n = C.c_uint64()
        capi.check(self.L.rrt_render_hit_dump(self.h, 1, None, 0, C.byref(n)))
        out = np.zeros((max(1, n.value), 6))
        capi.check(self.L.rrt_render_hit_dump(self.h, 1, out.ctypes.data, n.value, C.byref(n)))
        out = out[: n.value]
        out = out[~np.isnan(out[:, 0])]
        order = np.lexsort((out[:, 2], out[:, 0], out[:, 1]))
        return out[order]
