"""SURVEY.md §8f row 3, procedural textures (texture/{bilerp,mix,scale,checkerboard,uv}.rs + mappings of
texture/mod.rs:206-347, make_textures renderprocess.rs:298-515).  CPU part: the oracle's evaluator against
hand-computed answers, the product's evaluator (csrc/texture_core.h run on the host through
rrt_texture_host_probe — the code the shade kernel runs) against the oracle bit for bit, and the two scene.json
loaders against each other.  The GPU render comparison is tests/test_gpu_render.py::test_textured_scene."""
import ctypes as C
import json
import math

import numpy as np
import pytest

import oracle_lib as O
import oracle_scene as S
from rs_ray_toy_b200 import render as R
from rs_ray_toy_b200 import capi, synth


def oracle_probe(table, uv, p, diff=None):
    L = O.lib()
    L.orc_texture_probe.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_texture_probe.restype = None
    table = np.ascontiguousarray(table, dtype=np.float64).reshape(-1, S.TEX_ROW)
    uv = np.ascontiguousarray(uv, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    d = None if diff is None else np.ascontiguousarray(diff, dtype=np.float64)
    out = np.zeros((table.shape[0], 3))
    L.orc_texture_probe(table.shape[0], table.ctypes.data, uv.ctypes.data, p.ctypes.data, None if d is None else d.ctypes.data,
                        out.ctypes.data)
    return out


def oracle_differentials(*vecs):
    L = O.lib()
    L.orc_differentials_probe.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_differentials_probe.restype = None
    a = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.float64).reshape(3) for v in vecs]))
    out = np.zeros(10)
    L.orc_differentials_probe(a.ctypes.data, out.ctypes.data)
    return out


def product_rows(table):
    """Oracle table rows (tests/oracle_scene.py layout) -> rrt_texture rows."""
    rows = []
    for r in np.asarray(table).reshape(-1, S.TEX_ROW):
        rows.append(R.texture(int(r[0]), [r[8 + 3 * k: 11 + 3 * k] for k in range(4)], mapping=int(r[2]), map8=r[20:28],
                              t1=int(r[4]), t2=int(r[5]), amount=int(r[6]), world_to_texture=r[28:44].reshape(4, 4), aa=int(r[3])))
    return rows


def _tex(cfg_float=(), cfg_rgb=()):
    return S.Textures({"float_texture": list(cfg_float), "rgb_texture": list(cfg_rgb)})


def test_known_answers():
    v = lambda r, g, b: {"values": [r, g, b]}
    t = _tex(
        cfg_float=[{"texture_name": "ramp", "texture_type": "BilerpTexture", "v00": 0.0},      # corners 0 1 0 1 (v10, v11 read "v01")
                   {"texture_name": "amt", "texture_type": "BilerpTexture", "v00": 0.25, "v01": 0.25}],
        cfg_rgb=[{"texture_name": "a", "texture_type": "BilerpTexture", "v00": v(1, 2, 3), "v01": v(1, 2, 3)},
                 {"texture_name": "b", "texture_type": "BilerpTexture", "v00": v(5, 6, 7), "v01": v(5, 6, 7)},
                 {"texture_name": "chk", "texture_type": "CheckerBoardTexture", "aamode": "none", "t1": "a", "t2": "b",
                  "mapping": {"mapping": "uv", "su": 4.0, "sv": 4.0, "du": 0.0, "dv": 0.0}},
                 {"texture_name": "chk3", "texture_type": "CheckerBoardTexture", "dimension": 3, "t1": "a", "t2": "b"},
                 {"texture_name": "amt", "texture_type": "MixTexture", "t1": "a", "t2": "amt"},   # t2 unknown as rgb -> 1; amount = float "amt"
                 {"texture_name": "sc", "texture_type": "ScaleTexture", "t1": "a", "t2": "b"},
                 {"texture_name": "uv", "texture_type": "UVTexture", "mapping": {"mapping": "uv", "su": 2.0, "sv": 3.0}},
                 {"texture_name": "pl", "texture_type": "UVTexture",
                  "mapping": {"mapping": "planar", "v1": [1, 0, 0], "v2": [0, 0, 2], "udelta": 0.5, "vdelta": 0.25}}])
    uv, p = (0.3, 0.6), (1.25, -2.5, 3.125)
    vals = oracle_probe(t.table(), uv, p)
    f, g = t.f, t.rgb
    # bilerp.rs:31-44 with corners (0, 1, 0, 1): (1-s)(1-t)*0 + (1-s)t + s(1-t)*0 + s t = t
    assert vals[f["ramp"]][0] == pytest.approx(0.6, abs=1e-15)
    # checkerboard.rs:57-64: floor(1.2) + floor(2.4) = 3, odd -> tex2
    assert vals[g["chk"]].tolist() == [5, 6, 7]
    # checkerboard.rs:121-131: floor(1.25) + floor(-2.5) + floor(3.125) = 1 - 3 + 3 = 1, odd -> tex2
    assert vals[g["chk3"]].tolist() == [5, 6, 7]
    # mix.rs:33-38: a * 0.75 + 1 * 0.25
    assert vals[g["amt"]].tolist() == [1 * 0.75 + 0.25, 2 * 0.75 + 0.25, 3 * 0.75 + 0.25]
    assert vals[g["sc"]].tolist() == [5, 12, 21]
    # uv.rs:20-27 with su 2, sv 3 and du = dv = 1 (the defaults once a mapping block exists, renderprocess.rs:582-583)
    assert vals[g["uv"]].tolist() == [2 * 0.3 + 1 - 1, 3 * 0.6 + 1 - 2, 0.0]
    # planar (texture/mod.rs:338-347): s = 0.5 + p.x, t = 0.25 + 2 p.z = 6.5
    assert vals[g["pl"]].tolist() == [0.75, 0.5, 0.0]
    # negative coordinates: floor(-0.2) = -1 -> (-1 + 0) % 2 = -1 != 0 in Rust and C alike -> tex2
    vals = oracle_probe(t.table(), (-0.05, 0.1), p)
    assert vals[g["chk"]].tolist() == [5, 6, 7]


def test_spherical_and_cylindrical_mappings():
    t = _tex(cfg_rgb=[{"texture_name": "s", "texture_type": "UVTexture", "mapping": {"mapping": "spherical"}, "world_pos": [1.0, 2.0, 3.0]},
                      {"texture_name": "c", "texture_type": "UVTexture", "mapping": {"mapping": "cylindrical"}, "world_pos": [1.0, 2.0, 3.0],
                       "scale": [2.0, 2.0, 2.0]}])
    p = np.array([1.0, 2.0, 3.0]) + np.array([0.0, 1.0, 1.0])
    vals = oracle_probe(t.table(), (0, 0), p)
    # v = normalize(0, 1, 1): theta = acos(1/sqrt 2) = pi/4 -> s = 1/4; phi = atan2(1, 0) = pi/2 -> t = 1/4
    assert vals[t.rgb["s"]][:2] == pytest.approx([0.25, 0.25], abs=1e-15)
    # cylinder: s = (pi + atan2(v.y, v.x)) / 2 pi = 3/4, t = v.z = 1/sqrt 2 (the scale drops out in the normalisation)
    assert vals[t.rgb["c"]][:2] == pytest.approx([0.75, 1 / math.sqrt(2)], abs=1e-15)


def test_host_evaluator_is_the_oracles_bit_for_bit(tmp_path):
    sc = S.load(synth.scene_textured(str(tmp_path)))
    table = sc.textures
    rows = product_rows(table)
    rng = np.random.default_rng(11)
    kinds = set(int(k) for k in table[:, 0])
    assert kinds == set(range(9)), kinds          # the scene covers every in-scope texture kind
    assert set(int(k) for k in table[:, 2]) == {0, 1, 2, 3}
    assert set(int(k) for k in table[table[:, 0] == S.TEX_CHECKER2D][:, 3]) == {0, 1}   # point-sampled and closed-form
    filtered = 0
    for _ in range(400):
        uv = rng.uniform(-2.0, 9.0, 2)
        p = rng.uniform(-8.0, 40.0, 3)
        a, b = oracle_probe(table, uv, p), R.texture_host_probe(rows, uv, p)
        assert np.array_equal(a, b), (uv, p, a, b)
        # with screen-space differentials (footprints from a hundredth of a check to several checks)
        diff = np.concatenate([rng.normal(0, 1, 6) * 10 ** rng.uniform(-3, 0.5), rng.normal(0, 1, 4) * 10 ** rng.uniform(-3, 0.5)])
        c, d = oracle_probe(table, uv, p, diff), R.texture_host_probe(rows, uv, p, diff)
        assert np.array_equal(c, d), (uv, p, diff, c, d)
        filtered += int(not np.array_equal(a, c))
    assert filtered > 100   # the closed-form filter did something


def test_closed_form_checkerboard_known_answers():
    v = lambda r, g, b: {"values": [r, g, b]}
    t = _tex(cfg_rgb=[{"texture_name": "a", "texture_type": "BilerpTexture", "v00": v(1, 1, 1), "v01": v(1, 1, 1)},
                      {"texture_name": "b", "texture_type": "BilerpTexture", "v00": v(0, 0, 0), "v01": v(0, 0, 0)},
                      {"texture_name": "chk", "texture_type": "CheckerBoardTexture", "t1": "a", "t2": "b",     # aamode: closedform
                       "mapping": {"mapping": "uv", "su": 1.0, "sv": 1.0, "du": 0.0, "dv": 0.0}}])
    i = t.rgb["chk"]
    assert t.table()[i][3] == 1.0
    z3 = [0.0] * 3

    def at(uv, dudx=0.0, dvdx=0.0, dudy=0.0, dvdy=0.0):
        return oracle_probe(t.table(), uv, (0, 0, 0), z3 + z3 + [dudx, dvdx, dudy, dvdy])[i][0]

    # a footprint inside one check: the point sample (checkerboard.rs:76-83)
    assert at((0.5, 0.5), 0.1, 0, 0, 0.1) == 1.0 and at((1.5, 0.5), 0.1, 0, 0, 0.1) == 0.0
    # footprint [0.75, 1.25] x [0.4, 0.6]: half in check (0, 0), half in check (1, 0).  bump_int(1.25) - bump_int(0.75)
    # = 0.25 - 0 -> sint = 0.25 / 0.5 = 0.5, tint = 0 -> area2 = 0.5 -> 0.5 * tex1 + 0.5 * tex2
    assert at((1.0, 0.5), 0.25, 0, 0, 0.1) == pytest.approx(0.5, abs=1e-15)
    # a quarter of the s-extent past the edge: [0.85, 1.05]: bump_int(1.05) = 0 + 2 * 0.025 -> sint = 0.05 / 0.2 = 0.25
    assert at((0.95, 0.5), 0.1, 0, 0, 0.1) == pytest.approx(0.75, abs=1e-12)
    # wider than a check in s: 50 % grey (:88-90)
    assert at((0.3, 0.5), 1.5, 0, 0, 0.1) == 0.5
    # ds = max(|dstdx|) is taken over BOTH components of dstdx (s and t derivatives along x), dt likewise (:70-71)
    assert at((0.95, 0.5), 0.0, 0.1, 0.0, 0.1) == pytest.approx(0.75, abs=1e-12)
    # a footprint that crosses an edge in s with NO extent in t divides 0 by 0 (:85-86), in the reference as here
    assert math.isnan(at((0.95, 0.5), 0.1, 0.0, 0.0, 0.0))


def test_compute_differentials():
    """interaction.rs:223-284, with the reference's y-plane slip (Q29: dot(n, ry_direction) in the numerator)."""
    rng = np.random.default_rng(5)
    p, n = np.array([0.0, 0.0, 5.0]), np.array([0.0, 0.0, -1.0])
    dpdu, dpdv = np.array([2.0, 0.0, 0.0]), np.array([0.0, 4.0, 0.0])
    o = np.zeros(3)
    rx_o, rx_d = o, np.array([0.1, 0.0, 1.0])
    ry_o, ry_d = o, np.array([0.0, 0.2, 1.0])
    out = oracle_differentials(p, n, dpdu, dpdv, rx_o, rx_d, ry_o, ry_d)
    # x: tx = -(n.rx_o - n.p) / n.rx_d = -(0 + 5) / -1 = 5 -> px = (0.5, 0, 5): dpdx = (0.5, 0, 0), dudx = 0.25
    assert out[:3].tolist() == [0.5, 0.0, 0.0] and out[6] == 0.25 and out[7] == 0.0
    # y as the reference computes it: ty = -(n.ry_d - n.p) / n.ry_d = -(-1 + 5) / -1 = 4 (5 was meant) -> py = (0, 0.8, 4)
    assert out[3:6].tolist() == [0.0, 0.8, -1.0] and out[8] == 0.0 and out[9] == pytest.approx(0.2, abs=1e-15)
    # a ray parallel to the surface: no differentials at all (:230-232)
    flat = oracle_differentials(p, n, dpdu, dpdv, rx_o, np.array([1.0, 0.0, 0.0]), ry_o, ry_d)
    assert not flat.any()
    # degenerate dpdu / dpdv: the solve fails, du / dv are zero but dpdx / dpdy stay (:276-283)
    deg = oracle_differentials(p, n, dpdu, dpdu, rx_o, rx_d, ry_o, ry_d)
    assert deg[:3].tolist() == [0.5, 0.0, 0.0] and not deg[6:].any()
    # the product's code is the oracle's, bit for bit
    for _ in range(500):
        vecs = [rng.normal(0, 3, 3) for _ in range(8)]
        if rng.uniform() < 0.2:
            vecs[1] = np.eye(3)[rng.integers(3)] * rng.choice([-1.0, 1.0])
        a, b = oracle_differentials(*vecs), R.differentials_host_probe(*vecs)
        assert np.array_equal(a, b), (vecs, a, b)


def test_loaders_agree(tmp_path):
    path = synth.scene_textured(str(tmp_path))
    sc = S.load(path)
    tex, mats, slots = R.json_texture_probe(path)
    assert len(tex) == sc.textures.shape[0] and len(mats) == sc.materials.shape[0]
    for row, t in zip(sc.textures, tex):
        assert (int(row[0]), int(row[2]), int(row[4]), int(row[5]), int(row[6])) == (t.kind, t.mapping, t.t1, t.t2, t.amount)
        assert np.array_equal(row[8:20], np.array([list(t.v[k]) for k in range(4)]).reshape(12))
        assert np.array_equal(row[20:28], np.array(list(t.map)))
        if t.kind == R.TEX_CHECKER3D or t.mapping >= R.TEXMAP_SPHERICAL:
            assert np.array_equal(row[28:44], np.array(list(t.world_to_texture)))
    assert np.array_equal(np.concatenate([sc.materials[:, 26:38], sc.materials[:, 54:65]], axis=1).astype(np.int32), slots)
    for row, m in zip(sc.materials, mats):
        assert int(row[0]) == m.kind
        assert np.array_equal(row[1:4], list(m.kd)) and np.array_equal(row[4:7], list(m.ks))
        assert row[19] == m.sigma and row[20] == m.roughness


def test_loaders_agree_on_translucent_disney_debug_and_mix(tmp_path):
    """make_materials' remaining arms (renderprocess.rs:678-720, 810-866): the product's loader and the oracle's build the
    same records — defaults, texture slots, the Disney block — skip a MixMaterial that names an unknown material and
    refuse the one the reference panics on (Q25)."""
    path = synth.scene_more_materials(str(tmp_path))
    sc = S.load(path)
    tex, mats, slots = R.json_texture_probe(path)
    assert len(mats) == sc.materials.shape[0] == 8          # m_mix has no entry
    assert [m.kind for m in mats] == [5, 5, 6, 6, 6, 6, 6, 7]
    assert np.array_equal(np.concatenate([sc.materials[:, 26:38], sc.materials[:, 54:65]], axis=1).astype(np.int32), slots)
    names = [n for n, _ in S.DISNEY_PARAMS]
    for row, m in zip(sc.materials, mats):
        assert int(row[0]) == m.kind
        assert np.array_equal(row[1:4], list(m.kd)) and np.array_equal(row[4:7], list(m.ks))
        assert np.array_equal(row[7:10], list(m.kr)) and np.array_equal(row[10:13], list(m.kt))
        assert row[20] == m.roughness and row[24] == m.remap_roughness
        if m.kind == 6:
            assert row[23] == m.eta and row[53] == m.thin
            assert [row[40 + k] for k in range(10)] == [getattr(m, n) for n in names]
            assert np.array_equal(row[50:53], list(m.scatter_distance))
    assert slots[6, R.SLOT_KD] >= 0 and slots[6, R.SLOT_METALLIC] >= 0 and slots[6, R.SLOT_ROUGHNESS] >= 0
    cfg = json.loads(open(path).read())
    cfg["materials"][-1]["mat2"] = "m_disney"
    p2 = tmp_path / "mix.json"
    p2.write_text(json.dumps(cfg))
    with pytest.raises(ValueError, match="Q25"):
        S.load(str(p2))
    with pytest.raises(capi.RrtError, match="Q25"):
        R.json_texture_probe(str(p2))


def test_out_of_scope_textures_are_refused(tmp_path):
    path = synth.scene_textured(str(tmp_path))
    cfg = json.loads(open(path).read())
    mapped = dict(cfg)
    mapped["rgb_texture"] = [dict(t) for t in cfg["rgb_texture"]]
    mapped["rgb_texture"][3]["mapping"] = {"mapping": "conical"}      # the reference panics (renderprocess.rs:601-606)
    p2 = tmp_path / "mapped.json"
    p2.write_text(json.dumps(mapped))
    with pytest.raises(ValueError):
        S.load(str(p2))
    with pytest.raises(capi.RrtError):
        R.json_texture_probe(str(p2))
    # a child that is not defined earlier is not a table the evaluator accepts
    bad = [R.texture(R.TEX_SCALE, t1=0, t2=1), R.texture(R.TEX_CONSTANT, [1.0])]
    with pytest.raises(capi.RrtError):
        R.texture_host_probe(bad, (0, 0), (0, 0, 0))
    # a float parameter naming a texture that does not exist panics in the reference (renderprocess.rs:621)
    missing = dict(cfg)
    missing["materials"] = [dict(cfg["materials"][1], sigma="no_such_texture")]
    p3 = tmp_path / "missing.json"
    p3.write_text(json.dumps(missing))
    with pytest.raises(ValueError):
        S.load(str(p3))
    with pytest.raises(capi.RrtError):
        R.json_texture_probe(str(p3))


def test_noise_textures_known_answers():
    """WindyTexture / WrinkledTexture over Perlin noise, fBm and turbulence (texture/mod.rs:75-188, windy.rs, wrinkled.rs)."""
    t = _tex(cfg_float=[{"texture_name": "w0", "texture_type": "WrinkledTexture", "octaves": 0},
                        {"texture_name": "w3", "texture_type": "WrinkledTexture", "octaves": 3, "omega": 0.5},
                        {"texture_name": "wind", "texture_type": "WindyTexture"},
                        {"texture_name": "w8", "texture_type": "WrinkledTexture"}])
    f = t.f
    z3 = [0.0] * 3
    p = (1.37, -2.61, 0.42)
    # no octaves at all: the partial-octave term alone, lerp(smooth_step(0) = 0, 0.2, |noise|) = 0.2 (:178-184)
    assert oracle_probe(t.table(), (0, 0), p)[f["w0"]][0] == 0.2
    # a footprint of several units clamps the octave count to 0 (:165): 0.2 + 0.2 (1 + 1/2 + 1/4) for three octaves,
    # and fBm's only term is weighted by smooth_step(0) = 0: the wind texture is 0
    wide = [5.0, 0.0, 0.0, 0.0, 5.0, 0.0, 0, 0, 0, 0]
    v = oracle_probe(t.table(), (0, 0), p, wide)
    assert v[f["w3"]][0] == pytest.approx(0.55, abs=1e-15) and v[f["wind"]][0] == 0.0
    # lattice points: Perlin noise is 0 there, so the first octave adds nothing; with a footprint of 0.25 units
    # n = -1 - 0.5 log2(1/16) = 1: one full octave (0 at a lattice point), then the partial term with smooth_step(0) = 0
    # -> 0.2 * omega, then the clamped octaves 1..2: 0.2 (1/2 + 1/4)
    quarter = [0.25, 0.0, 0.0, 0.0, 0.25, 0.0, 0, 0, 0, 0]
    v = oracle_probe(t.table(), (0, 0), (3.0, -2.0, 7.0), quarter)
    assert v[f["w3"]][0] == pytest.approx(0.5 * 0.2 + 0.2 * (0.5 + 0.25), abs=1e-15)
    # without differentials every octave is summed (log2(0) = -inf): |noise| <= 1 bounds turbulence by the geometric sum
    vals = np.array([oracle_probe(t.table(), (0, 0), q)[f["w8"]][0] for q in np.random.default_rng(2).uniform(-9, 9, (200, 3))])
    assert 0.0 < vals.min() and vals.max() < 2.0 and vals.std() > 0.02
    # a float texture's value reaches a material scalar through component 0; the rgb flavour is grey
    t2 = _tex(cfg_rgb=[{"texture_name": "w", "texture_type": "WrinkledTexture", "octaves": 4}])
    g = oracle_probe(t2.table(), (0, 0), p)[t2.rgb["w"]]
    assert g[0] == g[1] == g[2] > 0
