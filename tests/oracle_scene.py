"""Oracle-side restatement (Python; TEST INFRASTRUCTURE) of the reference's scene description
loader: src/renderprocess.rs (schema subset of SURVEY.md Appendix C) and src/objparser.rs.
It feeds the C++ oracle through oracle_lib.  The product has its own loader in C++
(rs_ray_toy_b200/csrc/scene_json.cpp); the parity tests run both on the same files."""
from __future__ import annotations

import copy
import ctypes as C
import json
import math
from pathlib import Path

import numpy as np

import oracle_lib as O

MAT_KIND = {"MatteMaterial": 0, "PlasticMaterial": 1, "MetalMaterial": 2, "MirrorMaterial": 3, "GlassMaterial": 4,
            "TranslucentMaterial": 5, "DisneyMaterial": 6, "Debug": 7}
FILTER_KIND = {"BoxFilter": 0, "GaussianFilter": 1, "TriangleFilter": 2}


def copper_rgb():
    """COPPER_N / COPPER_K as RGB (material/metal.rs tables through RGBSpectrum::from_sampled,
    spectrum.rs:2701-2727); values in tests/golden/copper_rgb.json (make_copper_rgb.py)."""
    g = json.loads((Path(__file__).resolve().parent / "golden" / "copper_rgb.json").read_text())
    return g["eta"], g["k"]


def parse_obj(path):
    """objparser.rs:83-247: v / vt / vn / f only, first three vertices of a face, 1-based indices;
    `vn` is normalised on read (:130); index triples are kept only when every one is in range."""
    p, uv, n, vi, ni, uvi = [], [], [], [], [], []
    for line in Path(path).read_text().splitlines():
        sp = line.split()
        if not sp:
            continue
        if sp[0] == "v":
            p.append([float(sp[1]), float(sp[2]), float(sp[3])])
        elif sp[0] == "vt":
            uv.append([float(sp[1]), float(sp[2]) if len(sp) > 2 else 0.0])
        elif sp[0] == "vn":
            v = np.array([float(sp[1]), float(sp[2]), float(sp[3])])
            n.append((v / math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])).tolist())
        elif sp[0] == "f":
            if len(sp) < 4:
                raise ValueError("ParseObjError: Failed to get face element")
            el = []
            for tok in sp[1:4]:
                parts = []
                for s in tok.split("/"):
                    try:
                        parts.append(int(s) - 1 if int(s) >= 0 and s.isdigit() else None)
                    except ValueError:
                        parts.append(None)
                parts = (parts + [None, None, None])[:3]
                el.append(parts)
            if all(e[0] is not None for e in el):
                vi += [e[0] for e in el]
                if all(e[1] is not None for e in el) and uv and all(e[1] < len(uv) for e in el):
                    uvi += [e[1] for e in el]
                if all(e[2] is not None for e in el) and n and all(e[2] < len(n) for e in el):
                    ni += [e[2] for e in el]
    ntri = len(vi) // 3
    return dict(
        p=np.array(p, dtype=np.float64).reshape(-1, 3), vi=np.array(vi, dtype=np.uint32).reshape(-1, 3),
        n=np.array(n, dtype=np.float64).reshape(-1, 3) if n else None,
        ni=np.array(ni, dtype=np.uint32).reshape(-1, 3) if len(ni) == 3 * ntri and ntri else None,
        uv=np.array(uv, dtype=np.float64).reshape(-1, 2) if uv else None,
        uvi=np.array(uvi, dtype=np.uint32).reshape(-1, 3) if len(uvi) == 3 * ntri and ntri else None)


def _xyz(cfg, key, default):
    v = cfg.get(key)
    if isinstance(v, list) and len(v) >= 3:
        return [float(v[0]), float(v[1]), float(v[2])]
    return list(default)


def to_world(cfg):
    """make_to_world (renderprocess.rs:242-252)."""
    return O.make_to_world(_xyz(cfg, "world_pos", (0, 0, 0)), _xyz(cfg, "rotation_axis", (0, 0, 0)),
                           float(cfg.get("rotation_angle", 0.0)), _xyz(cfg, "scale", (1, 1, 1)))


def _spectrum(cfg, key, default):
    v = cfg.get(key)
    if isinstance(v, dict) and isinstance(v.get("values"), list):
        return [float(x) for x in v["values"][:3]]
    return [float(default)] * 3 if not isinstance(default, (list, tuple)) else list(default)


TEX_ROW = 48
TEX_CONST, TEX_BILERP, TEX_SCALE, TEX_MIX, TEX_CHECKER2D, TEX_CHECKER3D, TEX_UV, TEX_WINDY, TEX_WRINKLED, TEX_IMAGE = range(10)
WRAP = {"repeat": 0, "black": 1, "clamp": 2}


class Textures:
    """make_textures (renderprocess.rs:298-515) flattened into one table in definition order: float textures first,
    then rgb textures; a texture can only name textures defined before it, an unknown name falls back to a constant
    (get_text_fallback, :282-296).  Row layout: oracle_capi.cpp orc_set_textures.
    A BilerpTexture whose corners agree is stored as the constant it is (the loader reads v10 and v11 from key
    "v01", :326-329,439-442) — the device loader does the same, so both sides drop the 1-ulp sum-of-weights factor."""

    def __init__(self, cfg):
        self.rows, self.f, self.rgb = [], {}, {}
        self.images = []   # (row, filename, do_trilinear, max_aniso, wrap): decoded once the scene exists
        for t in cfg.get("float_texture", []) or []:
            self._add(t, False)
        for t in cfg.get("rgb_texture", []) or []:
            self._add(t, True)
        if len(self.rows) > 32:
            raise ValueError("more than 32 textures")

    def table(self):
        return np.array(self.rows).reshape(-1, TEX_ROW)

    def _row(self, kind, is_rgb):
        r = np.zeros(TEX_ROW)
        r[0], r[1] = kind, 1.0 if is_rgb else 0.0
        r[4:7] = -1
        r[20:24] = (1, 1, 0, 0)
        r[28:44] = np.eye(4).reshape(16)
        self.rows.append(r)
        return r, len(self.rows) - 1

    def const(self, value, is_rgb):
        r, i = self._row(TEX_CONST, is_rgb)
        r[8:11] = value if is_rgb else [value, 0.0, 0.0]
        return i

    def _value(self, t, key, default, is_rgb):
        return _spectrum(t, key, default) if is_rgb else float(t.get(key, default))

    def _child(self, names, name, default, is_rgb):
        return names[name] if name in names else self.const([default] * 3 if is_rgb else default, is_rgb)

    def _mapping(self, r, t):
        mp = t.get("mapping")
        if mp is None:
            return
        kind = mp.get("mapping", "uv")
        if kind == "uv":     # NB du / dv default to 1 when a mapping block is given (renderprocess.rs:582-583)
            r[2] = 0
            r[20:24] = [float(mp.get("su", 1.0)), float(mp.get("sv", 1.0)), float(mp.get("du", 1.0)), float(mp.get("dv", 1.0))]
        elif kind == "planar":
            r[2] = 1
            r[20:23] = _xyz(mp, "v1", (1, 0, 0))
            r[23:26] = _xyz(mp, "v2", (0, 1, 0))
            r[26], r[27] = float(mp.get("udelta", 0.0)), float(mp.get("vdelta", 0.0))
        elif kind in ("spherical", "cylindrical"):   # Transform::inverse(to_world) of the TEXTURE's block (:594-599)
            r[2] = 2 if kind == "spherical" else 3
            _, inv = to_world(t)
            r[28:44] = inv.reshape(16)
        else:
            raise ValueError(f"texture mapping {kind!r} is outside the restated subset")

    def _add(self, t, is_rgb):
        names = self.rgb if is_rgb else self.f
        ty, name = t.get("texture_type", ""), t.get("texture_name", "DefaultTextureName")
        one, zero = (1.0, 0.0)
        if ty == "BilerpTexture":
            v00, v01 = self._value(t, "v00", 0.0, is_rgb), self._value(t, "v01", 1.0, is_rgb)
            v10, v11 = self._value(t, "v01", 0.0, is_rgb), self._value(t, "v01", 1.0, is_rgb)
            if v00 == v01 == v10 == v11:
                names[name] = self.const(v00, is_rgb)
                return
            r, i = self._row(TEX_BILERP, is_rgb)
            for k, v in enumerate((v00, v01, v10, v11)):
                r[8 + 3 * k: 11 + 3 * k] = v if is_rgb else [v, 0.0, 0.0]
            self._mapping(r, t)
        elif ty in ("WindyTexture", "WrinkledTexture"):   # IdentityMapping3D::new(to_world) (renderprocess.rs:376-388)
            r, i = self._row(TEX_WINDY if ty == "WindyTexture" else TEX_WRINKLED, is_rgb)
            m, _ = to_world(t)
            r[28:44] = m.reshape(16)
            if ty == "WrinkledTexture":
                r[20], r[21] = float(int(t.get("octaves", 8))), float(t.get("omega", 0.5))
        elif ty == "UVTexture" and is_rgb:
            r, i = self._row(TEX_UV, True)
            self._mapping(r, t)
        elif ty == "ImageTexture" and is_rgb:   # make_tex_info / load_image (renderprocess.rs:517-566): `gamma`, `scale` unused
            r, i = self._row(TEX_IMAGE, True)
            self._mapping(r, t)
            self.images.append((i, t.get("filename", "DefaultTexture"), bool(t.get("do_trilinear", False)),
                                float(t.get("max_aniso", 8.0)), WRAP.get(t.get("wrap", "repeat"), 0)))
        elif ty == "ScaleTexture":
            c1 = self._child(names, t.get("t1", "ErrorTextureName"), one, is_rgb)
            c2 = self._child(names, t.get("t2", "ErrorTextureName"), one, is_rgb)
            r, i = self._row(TEX_SCALE, is_rgb)
            r[4], r[5] = c1, c2
        elif ty == "MixTexture":   # the amount texture is looked up under the key "t2" as well (:319,411)
            c1 = self._child(names, t.get("t1", "ErrorTextureName"), zero, is_rgb)
            c2 = self._child(names, t.get("t2", "ErrorTextureName"), one, is_rgb)
            amt = self._child(self.f, t.get("t2", "ErrorTextureName"), 0.5, False)
            r, i = self._row(TEX_MIX, is_rgb)
            r[4], r[5], r[6] = c1, c2, amt
        elif ty == "CheckerBoardTexture":
            dim = int(t.get("dimension", 2))
            if dim not in (2, 3):
                return
            c1 = self._child(names, t.get("t1", "ErrorTextureName"), one, is_rgb)
            c2 = self._child(names, t.get("t2", "ErrorTextureName"), zero, is_rgb)
            if dim == 2:
                r, i = self._row(TEX_CHECKER2D, is_rgb)
                r[3] = 0.0 if t.get("aamode", "closedform") == "none" else 1.0   # AAMethod (renderprocess.rs:357-366)
                self._mapping(r, t)
            else:
                r, i = self._row(TEX_CHECKER3D, is_rgb)
                m, _ = to_world(t)   # IdentityMapping3D::new(to_world): the matrix is used as world_to_texture
                r[28:44] = m.reshape(16)
            r[4], r[5] = c1, c2
        else:
            names[name] = -2   # known name, type outside the subset: an error only if a material uses it
            return
        names[name] = i

    def _resolve(self, names, name, is_rgb):
        i = names[name]
        if i == -2:
            raise ValueError(f"texture {name!r} has a type outside the restated subset")
        if self.rows[i][0] == TEX_CONST:   # constants stay constants in the material record
            v = self.rows[i][8:11]
            return (list(v) if is_rgb else float(v[0])), -1
        return None, i

    def fval(self, cfg, key, default):
        """-> (constant, texture id or -1)"""
        name = cfg.get(key)
        if isinstance(name, str):
            if name not in self.f:
                raise ValueError(f"float texture {name!r} does not exist (the reference panics: renderprocess.rs:621)")
            v, i = self._resolve(self.f, name, False)
            return (default if v is None else v), i
        return default, -1

    def rgbval(self, cfg, key, default):
        name = cfg.get(key)
        d = [default] * 3 if not isinstance(default, (list, tuple)) else list(default)
        if isinstance(name, str) and name in self.rgb:
            v, i = self._resolve(self.rgb, name, True)
            return (d if v is None else v), i
        return d, -1


MAT_ROW = 72
# Disney block of a material row: 40 + k the value, 54 + k its texture id (k = 10: scatter_distance, value at 50:53)
DISNEY_PARAMS = (("metallic", 0.0), ("specular_tint", 0.0), ("anisotropic", 0.0), ("sheen", 0.0), ("sheen_tint", 0.5),
                 ("clearcoat", 0.0), ("clearcoat_gloss", 1.0), ("spec_trans", 0.0), ("flatness", 0.0), ("diff_trans", 1.0))
# texture-id slots of a material row (26 + k): kd ks kr kt eta_rgb k_rgb sigma roughness u_roughness v_roughness eta
T_KD, T_KS, T_KR, T_KT, T_ETA_RGB, T_K_RGB, T_SIGMA, T_ROUGH, T_UR, T_VR, T_ETA, T_BUMP = range(12)


def material_row(cfg, tex: Textures):
    """make_materials (renderprocess.rs:664-871) -> the oracle's 72-double material record."""
    kind = MAT_KIND.get(cfg.get("material_type", ""))
    if kind is None:
        return None
    r = np.zeros(MAT_ROW)
    r[0] = kind
    r[21] = r[22] = -1.0
    r[26:38] = -1
    r[54:65] = -1

    def rgb(lo, slot, key, default):
        r[lo:lo + 3], r[26 + slot] = tex.rgbval(cfg, key, default)

    def flt(at, slot, key, default):
        r[at], r[26 + slot] = tex.fval(cfg, key, default)

    cu_n, cu_k = copper_rgb()
    if kind == 0:
        rgb(1, T_KD, "kd", 0.5)
        flt(19, T_SIGMA, "sigma", 0.0)
    elif kind == 1:
        rgb(1, T_KD, "kd", 0.25)
        rgb(4, T_KS, "ks", 0.25)
        flt(20, T_ROUGH, "roughness", 0.1)
    elif kind == 2:
        rgb(13, T_ETA_RGB, "eta", cu_n)
        rgb(16, T_K_RGB, "k", cu_k)
        flt(20, T_ROUGH, "roughness", 0.01)
        flt(21, T_UR, "u_roughness", -1.0)
        flt(22, T_VR, "v_roughness", -1.0)
    elif kind == 3:
        rgb(7, T_KR, "kr", 0.9)
    elif kind == 4:
        rgb(7, T_KR, "kr", 1.0)
        rgb(10, T_KT, "kt", 1.0)
        flt(23, T_ETA, "eta", 1.5)
        flt(21, T_UR, "u_roughness", 0.0)
        flt(22, T_VR, "v_roughness", 0.0)
    elif kind == 5:                       # renderprocess.rs:695-720
        rgb(1, T_KD, "kd", 0.25)
        rgb(4, T_KS, "ks", 0.25)
        flt(20, T_ROUGH, "roughness", 0.1)
        rgb(7, T_KR, "reflect", 0.25)
        rgb(10, T_KT, "transmit", 0.25)
    elif kind == 6:                       # renderprocess.rs:810-860
        rgb(1, T_KD, "color", 0.5)
        flt(23, T_ETA, "eta", 1.5)
        flt(20, T_ROUGH, "roughness", 0.5)
        for k, (key, default) in enumerate(DISNEY_PARAMS):
            r[40 + k], r[54 + k] = tex.fval(cfg, key, default)
        r[50:53], r[64] = tex.rgbval(cfg, "scatter_distance", 0.0)
        r[53] = 1.0 if cfg.get("thin", False) else 0.0
    if kind == 7:                         # DebugMaterial takes no parameters, not even a bump map (renderprocess.rs:861-863)
        return r
    r[24] = 1.0 if cfg.get("remap_roughness", False) else 0.0
    bump = cfg.get("bump_map")            # fetch_float_texture_opt(.., "bump_map", None) (renderprocess.rs:704-705)
    r[37] = -1
    if isinstance(bump, str):
        if bump not in tex.f:
            raise ValueError(f"float texture {bump!r} does not exist (the reference panics: renderprocess.rs:634)")
        if tex.f[bump] == -2:
            raise ValueError(f"texture {bump!r} has a type outside the restated subset")
        r[37] = tex.f[bump]
    return r


LIGHT_ROW = 80


def light_row(cfg, meshes=None):
    """make_light (renderprocess.rs:967-1053): point, distant and diffuse (area) lights.
    Row: 0 kind | 1-3 intensity / lemit | 4-6 dir | 7-22 light_to_world | 23 shape kind (0 sphere, 1 triangle) |
    24-39 sphere obj_to_world | 40-55 its inverse | 56 radius 57 z_min 58 z_max 59 phi_max (deg) |
    60-68 triangle p0 p1 p2 | 69-77 its vertex normals | 78 has normals."""
    r = np.zeros(LIGHT_ROW)
    m, _ = to_world(cfg)
    r[7:23] = m.reshape(16)
    t = cfg.get("light_type")
    if t == "point":
        r[0] = 0
        r[1:4] = _spectrum(cfg, "spectrum", 1.0)
    elif t == "distant":
        r[0] = 1
        l, sc = np.array(_spectrum(cfg, "l", 1.0)), np.array(_spectrum(cfg, "scale", 1.0))
        r[1:4] = l * sc
        r[4:7] = np.array(_xyz(cfg, "from", (0, 0, 0))) - np.array(_xyz(cfg, "to", (0, 0, 1)))
    elif t == "diffuse":
        r[0] = 2
        r[1:4] = _spectrum(cfg, "spectrum", 1.0)
        shp = cfg["light_shape"]           # "Shape Required for a DiffuseLight!" (renderprocess.rs:1015)
        if shp.get("shape_type") == "sphere":   # make_sphere (renderprocess.rs:1097-1106)
            sm, sinv = to_world(shp)
            radius = float(shp.get("radius", 1.0))
            r[23] = 0
            r[24:40] = sm.reshape(16)
            r[40:56] = sinv.reshape(16)
            r[56:60] = [radius, float(shp.get("z_min", -radius)), float(shp.get("z_max", radius)), float(shp.get("phi_max", 360.0))]
        elif shp.get("shape_type") == "triangle":  # mesh[tri_num] of a loaded obj, vertices untransformed (Q7)
            mesh = meshes[shp.get("obj_name", "")]
            k = int(shp.get("tri_num", 0))
            r[23] = 1
            r[60:69] = mesh["p"][mesh["vi"][k]].reshape(9)
            if mesh["n"] is not None and len(mesh["n"]) and mesh["ni"] is not None and len(mesh["ni"]):
                r[69:78] = mesh["n"][mesh["ni"][k]].reshape(9)
                r[78] = 1
        else:
            raise ValueError("Failed to parse a Shape (renderprocess.rs:1094)")
    elif t == "infinite":   # renderprocess.rs:1032-1046; r[23] = index of the decoded map (set by LoadedScene)
        r[0] = 3
        l, sc = np.array(_spectrum(cfg, "l", 1.0)), np.array(_spectrum(cfg, "scale", 1.0))
        r[1:4] = l * sc
        _, inv = to_world(cfg)
        r[40:56] = inv.reshape(16)
        r[23] = -1
    else:
        raise ValueError(f"light type {t!r} is outside the restated subset")
    return r


def decode_rgb8(path):
    """image::io::Reader::open(..).decode().into_rgb8(): 8-bit RGB rows, top row first (alpha dropped)."""
    from PIL import Image
    im = Image.open(path)
    if im.mode not in ("RGB", "RGBA", "L", "LA", "P"):
        raise ValueError(f"{path}: image mode {im.mode} is outside the restated subset")
    a = np.ascontiguousarray(np.array(im.convert("RGBA"))[..., :3] if im.mode in ("RGBA", "LA", "P") else np.array(im.convert("RGB")), dtype=np.uint8)
    return a


def render_params(cfg, seed=1, tile_mod=1, tile_rank=0, crop=None, want_dump=False):
    """make_film / make_camera / make_sampler / make_integrator (renderprocess.rs:1306-1499)."""
    film, cam, smp, integ = cfg["Film"], cfg["Camera"], cfg["Sampler"], cfg["Integrator"]
    p = np.zeros(48)
    p[0], p[1] = int(film.get("xres", 1280)), int(film.get("yres", 720))
    p[2] = float(film.get("diagonal", 35.0))
    flt = film["Filter"]
    ftype = flt.get("filter_type", "BoxFilter")
    kind = FILTER_KIND.get(ftype, 0)
    rad = flt.get("radius")
    default_r = (0.5, 0.5) if kind == 0 else (2.0, 2.0)
    rx, ry = (float(rad[0]), float(rad[1])) if isinstance(rad, list) and len(rad) >= 2 else default_r
    p[3], p[4], p[5], p[6] = kind, rx, ry, float(flt.get("alpha", 2.0))
    p[7] = float(film.get("scale", 1.0))
    p[8] = float(film.get("max_sample_luminance", math.inf))
    p[9:12] = _xyz(cam, "world_pos", (0, 0, 0))
    p[12:15] = _xyz(cam, "look", (1, 1, 1))
    p[15:18] = _xyz(cam, "up", (0, 0, 1))
    p[18], p[19] = float(cam.get("shutter_open", 0.0)), float(cam.get("shutter_close", 1.0))
    p[20], p[21] = float(cam.get("aperture_diameter", 1.0)), float(cam.get("focus_distance", 10.0))
    p[22] = 1.0 if cam.get("simple_weighting", True) else 0.0
    if smp.get("sampler_type") == "StratifiedSampler":      # make_sampler (renderprocess.rs:1308-1314)
        p[38], p[39] = 1, 1.0 if smp.get("jitter", True) else 0.0
        p[40], p[41], p[42] = int(smp.get("xsamp", 4)), int(smp.get("ysamp", 4)), int(smp.get("dimension", 4))
        p[23] = p[40] * p[41]
    elif smp.get("sampler_type") == "HaltonSampler":
        p[23] = int(smp.get("nsamp", 16))
        p[24] = 1.0 if smp.get("sample_at_center", False) else 0.0
    else:
        raise ValueError(f"Unsupported Sampler type (the reference panics: renderprocess.rs:1322)")
    p[25] = seed
    it = integ.get("integrator_type", "AO")
    if it == "Path":
        p[26], p[27], p[28] = 0, int(integ.get("max_depth", 5)), float(integ.get("rr_threshold", 1.0))
    elif it == "DirectLighting":
        p[26], p[27], p[28] = 1, int(integ.get("max_depth", 5)), 1.0
        p[37] = 1.0 if integ.get("light_strategy", "one") == "all" else 0.0   # renderprocess.rs:1413-1417
    elif it == "Debug":                                       # renderprocess.rs:1471-1481
        p[26], p[27], p[28] = 2, int(integ.get("max_depth", 5)), 1.0
    else:
        raise ValueError(f"integrator {it!r} is outside the restated subset")
    p[29], p[30] = tile_mod, tile_rank
    if crop is not None:
        p[31] = 1.0
        p[32:36] = crop
    p[36] = 1.0 if want_dump else 0.0
    lens = np.array(cam["lens_data"], dtype=np.float64).reshape(-1)
    return p, lens


class LoadedScene:
    def __init__(self, cfg, root: Path, tier=O.TIER_F):
        self.cfg = cfg
        self.scene = O.OracleScene(tier)
        tex = Textures(cfg)
        rows, self.mat_index = [], {}
        for m in cfg.get("materials", []) or []:
            if m.get("material_type") == "MixMaterial" and m.get("mat1", "") in self.mat_index and m.get("mat2", "") in self.mat_index:
                # renderprocess.rs:681-692 indexes scene_global.materials, which is still empty while make_materials runs (Q25)
                raise ValueError("MixMaterial over two existing materials: the reference panics while loading (Q25)")
            row = material_row(m, tex)
            if row is not None:
                self.mat_index[m.get("material_name", "DefaultMaterialName")] = len(rows)
                rows.append(row)
        self.materials = np.array(rows).reshape(-1, MAT_ROW)
        self.textures = tex.table()
        meshes = {}
        for o in cfg.get("objs", []) or []:
            meshes[o.get("obj_name", "DefaultObjName")] = parse_obj(root / o.get("filename", "DefaultObj"))
        s = self.scene
        agg = cfg["Aggregate"]
        for prim in agg.get("primitives", []) or []:
            mat = self.mat_index.get(prim.get("material_name", "DefaultMaterialName"))
            inst = prim.get("instances")
            if prim.get("primitive_type") == "sphere":
                if mat is None:
                    continue
                m, inv = to_world(prim)
                radius = float(prim.get("radius", 1.0))
                sp = s.add_sphere(m, inv, radius, float(prim.get("z_min", -radius)), float(prim.get("z_max", radius)),
                                  float(prim.get("phi_max", 360.0)))
                g = s.add_geo_sphere(sp, mat)
                if isinstance(inst, list):
                    for ic in inst:
                        im, iinv = to_world(ic)
                        s.add_prims(g, 1, s.add_xform(im, iinv))
                else:
                    s.add_prims(g, 1, -1)
            elif prim.get("primitive_type") == "triangle":
                mesh = meshes.get(prim.get("obj_name", "DefaultObjName"))
                if mesh is None or mat is None:
                    continue
                mid = s.add_mesh(mesh["p"], mesh["vi"], mesh["n"], mesh["ni"], mesh["uv"], mesh["uvi"])
                g0 = s.add_geo_triangles(mid, mat)
                nt = mesh["vi"].shape[0]
                if isinstance(inst, list):
                    for ic in inst:
                        im, iinv = to_world(ic)
                        s.add_prims(g0, nt, s.add_xform(im, iinv))
                else:
                    s.add_prims(g0, nt, -1)
        s.build(int(agg.get("max_prims_in_node", 4)))
        self.lights = np.array([light_row(l, meshes) for l in (cfg.get("lights", []) or [])]).reshape(-1, LIGHT_ROW)
        self.infinite_lights = np.array([light_row(l, meshes) for l in (cfg.get("infinite_lights", []) or [])]).reshape(-1, LIGHT_ROW)
        L = O.lib()
        L.orc_add_image.restype = C.c_int32
        L.orc_add_image.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int32, C.c_double, C.c_uint32]
        L.orc_add_env_image.restype = C.c_int32
        L.orc_add_env_image.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_set_infinite_lights.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        reached = set(int(t) for t in np.concatenate([self.materials[:, 26:38], self.materials[:, 54:65]], axis=1).reshape(-1) if t >= 0)
        for i in range(len(tex.rows) - 1, -1, -1):       # children have smaller indices
            if i in reached and tex.rows[i][0] in (TEX_SCALE, TEX_MIX, TEX_CHECKER2D, TEX_CHECKER3D):
                reached |= {int(tex.rows[i][4]), int(tex.rows[i][5])} | ({int(tex.rows[i][6])} if tex.rows[i][0] == TEX_MIX else set())
        for (row, fname, tri, aniso, wrap) in tex.images:
            if row not in reached and not (root / fname).exists():
                tex.rows[row][0] = TEX_CONST       # never evaluated (the sample scene declares such a texture)
                tex.rows[row][8:11] = 0.0
                continue
            img = decode_rgb8(root / fname)
            k = L.orc_add_image(s.h, img.shape[1], img.shape[0], img.ctypes.data, int(tri), aniso, wrap)
            if k < 0:
                raise RuntimeError(L.orc_last_error().decode())
            tex.rows[row][4] = k
        self.textures = tex.table()
        for rows, cfgs in ((self.lights, cfg.get("lights", []) or []), (self.infinite_lights, cfg.get("infinite_lights", []) or [])):
            for r, lc in zip(rows, cfgs):
                if r[0] == 3:
                    img = decode_rgb8(root / lc.get("mapname", ""))
                    r[23] = L.orc_add_env_image(s.h, img.shape[1], img.shape[0], img.ctypes.data)
        L.orc_set_materials.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_set_lights.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_set_textures.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_set_textures(s.h, self.textures.shape[0], self.textures.ctypes.data)
        L.orc_set_materials(s.h, self.materials.shape[0], self.materials.ctypes.data)
        L.orc_set_lights(s.h, self.lights.shape[0], self.lights.ctypes.data)
        L.orc_set_infinite_lights(s.h, self.infinite_lights.shape[0], self.infinite_lights.ctypes.data)

    def render(self, seed=1, nthreads=None, tile_mod=1, tile_rank=0, crop=None, want_dump=False):
        prm, lens = render_params(self.cfg, seed, tile_mod, tile_rank, crop, want_dump)
        return oracle_render(self.scene, prm, lens, nthreads, want_dump)


def oracle_render(scene, prm, lens, nthreads=None, want_dump=False):
    L = O.lib()
    L.orc_render.restype = C.c_int32
    L.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32]
    xres, yres = int(prm[0]), int(prm[1])
    npix = xres * yres
    rgb = np.zeros((yres, xres, 3))
    raw = np.zeros((yres, xres, 4))
    stats = np.zeros(16, dtype=np.uint64)
    dump_pix = npix if prm[31] == 0.0 else int(max(0, prm[34] - prm[32]) * max(0, prm[35] - prm[33]))   # a crop dumps its own pixels only
    cap = dump_pix * max(1, int(prm[23]) - 1) if want_dump else 0
    dump = np.zeros((max(cap, 1), 6))
    cnt = C.c_uint64(0)
    nthreads = O.hardware_threads() if nthreads is None else nthreads
    rc = L.orc_render(scene.h, prm.ctypes.data, lens.ctypes.data, lens.shape[0], rgb.ctypes.data, raw.ctypes.data,
                      stats.ctypes.data, dump.ctypes.data if want_dump else None, cap, C.byref(cnt), nthreads)
    if rc != 0:
        raise RuntimeError(L.orc_last_error().decode())
    names = ["camera_rays", "extension_rays", "shadow_rays", "bounces", "zero_weight", "asserts", "closest_rays",
             "closest_nodes", "closest_prims", "closest_max_stack", "any_rays", "any_nodes", "any_prims", "any_max_stack",
             "stack_overflow", "mis_probe_rays"]
    out = {"rgb": rgb, "raw": raw, "stats": {k: int(v) for k, v in zip(names, stats)}}
    if want_dump:
        out["dump"] = dump[: min(cap, cnt.value)]
    return out


def load(path, overrides=None, tier=O.TIER_F) -> LoadedScene:
    path = Path(path)
    cfg = json.loads(path.read_text())
    for k, v in (overrides or {}).items():
        cfg[k] = copy.deepcopy(v)
    return LoadedScene(cfg, path.parent, tier)
