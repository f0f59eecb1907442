"""CPU-only checks of the oracle's renderer restatement and of the product's host-side loader."""
import numpy as np

import oracle_lib as O
import oracle_scene as S
from rs_ray_toy_b200 import render, synth


def test_radical_inverse_known_answers():
    L = O.lib()
    import ctypes as C
    L.orc_radical_inverse.restype = C.c_double
    L.orc_radical_inverse.argtypes = [C.c_int32, C.c_uint64]
    # base 2: bit reversal; base 3: digit reversal (lowdiscrepancy.rs:188-236)
    assert L.orc_radical_inverse(0, 1) == 0.5 and L.orc_radical_inverse(0, 2) == 0.25 and L.orc_radical_inverse(0, 3) == 0.75
    assert abs(L.orc_radical_inverse(1, 1) - 1 / 3) < 1e-15 and abs(L.orc_radical_inverse(1, 5) - (2 / 3 + 1 / 9)) < 1e-15
    assert abs(L.orc_radical_inverse(2, 7) - (2 / 5 + 1 / 25)) < 1e-15
    L.orc_scrambled_radical_inverse.restype = C.c_double
    L.orc_scrambled_radical_inverse.argtypes = [C.c_int32, C.c_uint64, C.c_uint64]
    # identity permutations (seed 0): scrambled == plain for digits with perm[0] = 0
    for a in (1, 17, 123456):
        assert abs(L.orc_scrambled_radical_inverse(3, a, 0) - L.orc_radical_inverse(3, a)) < 1e-15


def test_halton_pixel_index_quirks():
    """halton.rs:23-105 at 640x360: scales 128 = 2^7 and 243 = 3^5, stride 31104; the pixel offset
    of the base-2 term is reversed over base_exponents[1] = 5 digits (Q13), so two pixels that
    differ only above bit 4 of x share their Halton indices."""
    import ctypes as C
    L = O.lib()
    L.orc_halton_index.restype = C.c_uint64
    L.orc_halton_index.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.POINTER(C.c_uint64)]
    stride = C.c_uint64()
    i0 = L.orc_halton_index(640, 360, 0, 0, 1, C.byref(stride))
    assert stride.value == 128 * 243 and i0 == stride.value
    a = L.orc_halton_index(640, 360, 3, 7, 2, None)
    b = L.orc_halton_index(640, 360, 3 + 32, 7, 2, None)
    assert a == b            # Q13
    assert L.orc_halton_index(640, 360, 3 + 128, 7, 2, None) == a   # kMaxResolution wrap (halton.rs:82-85)
    assert L.orc_halton_index(640, 360, 4, 7, 2, None) != a


def test_seeded_permutations_are_permutations():
    import ctypes as C
    L = O.lib()
    L.orc_halton_perms.argtypes = [C.c_uint64, C.c_void_p, C.c_uint32]
    primes = [2, 3, 5, 7, 11, 13]
    n = sum(primes)
    for seed in (0, 1, 99):
        buf = np.zeros(n, dtype=np.uint16)
        L.orc_halton_perms(seed, buf.ctypes.data, n)
        off = 0
        for p in primes:
            assert sorted(buf[off:off + p].tolist()) == list(range(p))
            if seed == 0:
                assert buf[off:off + p].tolist() == list(range(p))
            off += p


def _probe(mat_row, wo, wi, u, allow=True):
    import ctypes as C
    L = O.lib()
    pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.orc_bsdf_probe.argtypes = [pd, pd, pd, pd, C.c_int32, pd]
    out = np.zeros(12)
    L.orc_bsdf_probe(np.ascontiguousarray(mat_row, dtype=np.float64), np.array(wo, dtype=np.float64),
                     np.array(wi, dtype=np.float64), np.array(u, dtype=np.float64), int(allow), out)
    return out


def test_bsdf_known_answers():
    tex = S.Textures({})
    wo = np.array([0.3, 0.2, np.sqrt(1 - 0.13)])
    wi = np.array([-0.5, 0.1, np.sqrt(1 - 0.26)])
    # Matte kd=0.5: f = kd / pi, pdf = cos / pi (reflection.rs:807-840, :480-486)
    out = _probe(S.material_row({"material_type": "MatteMaterial"}, tex), wo, wi, (0.3, 0.6))
    assert np.allclose(out[:3], 0.5 / np.pi) and np.isclose(out[3], wi[2] / np.pi)
    assert np.isclose(out[10], out[9] / np.pi) and out[11] == 5       # sampled: DIFFUSE | REFLECTION
    # Mirror: specular reflection of wo about +z, f = kr / |cos|, pdf 1 (reflection.rs:638-649)
    out = _probe(S.material_row({"material_type": "MirrorMaterial"}, tex), wo, wi, (0.3, 0.6))
    assert np.allclose(out[:4], 0.0) and np.allclose(out[7:10], [-wo[0], -wo[1], wo[2]]) and out[10] == 1.0
    assert np.allclose(out[4:7], 0.9 / wo[2]) and out[11] == 17
    # Plastic (Q15): two reflective lobes -> pdf of the chosen lobe / 2; u0 < 0.5 picks the Lambertian
    out = _probe(S.material_row({"material_type": "PlasticMaterial"}, tex), wo, wi, (0.25, 0.6))
    assert np.isclose(out[10], out[9] / np.pi / 2.0) and np.allclose(out[4:7], 0.25 / np.pi)
    out2 = _probe(S.material_row({"material_type": "PlasticMaterial"}, tex), wo, wi, (0.75, 0.6))
    assert out2[11] == 9 and out2[10] > 0                             # GLOSSY | REFLECTION
    # Glass with multiple lobes allowed: FresnelSpecular, type SPECULAR|ALL; f(wo,wi) == 0
    out = _probe(S.material_row({"material_type": "GlassMaterial"}, tex), wo, wi, (0.9, 0.6))
    assert np.allclose(out[:4], 0.0) and out[11] == 18 and out[9] < 0  # transmitted below the surface
    # energy: Lambert sample weight f * cos / pdf == kd
    out = _probe(S.material_row({"material_type": "MatteMaterial"}, tex), wo, wi, (0.11, 0.83))
    assert np.allclose(out[4:7] * out[9] / out[10], 0.5)


def _const_tex(values):
    """A Textures table of constant float / rgb textures named after their position: (name -> value)."""
    ft = [{"texture_name": n, "texture_type": "BilerpTexture", "v00": v, "v01": v} for n, v in values.items() if not isinstance(v, tuple)]
    ct = [{"texture_name": n, "texture_type": "BilerpTexture", "v00": {"values": list(v)}, "v01": {"values": list(v)}}
          for n, v in values.items() if isinstance(v, tuple)]
    return S.Textures({"float_texture": ft, "rgb_texture": ct})


def test_translucent_disney_debug_known_answers():
    """The lobes of translucent.rs / disney.rs / debug_material.rs through Bsdf::f / pdf / sample_f, against the formulas
    written out independently in numpy — including what the reference does differently from pbrt: the clearcoat lobe is
    typed neither reflection nor transmission, so f() never contains it (Q37), and gtr1 divides by log10 (Q36)."""
    wo = np.array([0.3, 0.2, np.sqrt(1 - 0.13)])
    wi = np.array([-0.5, 0.1, np.sqrt(1 - 0.26)])
    wt = wi * np.array([1, 1, -1.0])
    sw = lambda c: np.clip(1 - c, 0, 1) ** 5
    # ---- Translucent without its glossy pair: Lambertian reflection r * kd and Lambertian transmission t * kd
    tex = _const_tex({"kd": (0.3, 0.5, 0.4), "black": (0.0, 0.0, 0.0), "r": (0.5, 0.6, 0.7), "t": (0.6, 0.5, 0.2)})
    row = S.material_row({"material_type": "TranslucentMaterial", "kd": "kd", "ks": "black", "reflect": "r", "transmit": "t"}, tex)
    out = _probe(row, wo, wi, (0.25, 0.6))
    assert np.allclose(out[:3], np.array([0.5, 0.6, 0.7]) * [0.3, 0.5, 0.4] / np.pi, rtol=1e-14)
    assert np.isclose(out[3], wi[2] / np.pi / 2, rtol=1e-14) and out[11] == 5 and out[9] > 0
    out = _probe(row, wo, wt, (0.75, 0.6))       # the second lobe: sampled below the surface, DIFFUSE | TRANSMISSION
    assert np.allclose(out[:3], np.array([0.6, 0.5, 0.2]) * [0.3, 0.5, 0.4] / np.pi, rtol=1e-14)
    assert np.isclose(out[3], wi[2] / np.pi / 2, rtol=1e-14) and out[11] == 6 and out[9] < 0
    assert np.isclose(out[10], -out[9] / np.pi / 2, rtol=1e-14)     # Q15: the other lobe's pdf (0 there) is added, then / 2
    # all four lobes; a black reflect AND transmit -> no Bsdf at all (f = 0, nothing sampled)
    row4 = S.material_row({"material_type": "TranslucentMaterial", "kd": "kd", "reflect": "r", "transmit": "t"}, tex)
    kinds = [int(_probe(row4, wo, wi, ((k + 0.5) / 4, 0.4))[11]) for k in range(4)]
    assert kinds == [5, 6, 9, 10]
    none = _probe(S.material_row({"material_type": "TranslucentMaterial", "reflect": "black", "transmit": "black"}, tex), wo, wi, (0.3, 0.3))
    assert not none.any()

    # ---- Disney, solid: DisneyDiffuse + DisneyRetro + MicrofacetReflection(DisneyFresnel, separable G)
    col, rough, eta, metallic, tint, aniso = np.array([0.6, 0.3, 0.2]), 0.4, 1.5, 0.3, 0.4, 0.5
    tex = _const_tex({"col": tuple(col), "rough": rough, "metallic": metallic, "tint": tint, "aniso": aniso, "sheen": 0.6, "stint": 0.3,
                      "cc": 0.8, "ccg": 0.7, "strans": 0.5, "flat": 0.4, "dt": 0.6})
    base = {"material_type": "DisneyMaterial", "color": "col", "roughness": "rough", "metallic": "metallic", "specular_tint": "tint",
            "anisotropic": "aniso"}
    row = S.material_row(base, tex)
    out = _probe(row, wo, wi, (0.1, 0.6))
    dw = 1 - metallic
    fo, fi = sw(wo[2]), sw(wi[2])
    wh = (wo + wi) / np.linalg.norm(wo + wi)
    cd = wi @ wh
    diffuse = col * dw / np.pi * (1 - fo / 2) * (1 - fi / 2)
    rr = 2 * rough * cd * cd
    retro = col * dw / np.pi * rr * (fo + fi + fo * fi * (rr - 1))
    aspect = np.sqrt(1 - aniso * 0.9)
    ax, ay = max(rough ** 2 / aspect, 1e-3), max(rough ** 2 * aspect, 1e-3)

    def D(h):
        c2 = h[2] ** 2
        s2 = 1 - c2
        return 1 / (np.pi * ax * ay * c2 * c2 * (1 + (s2 / c2) * (h[0] ** 2 / s2 / ax ** 2 + h[1] ** 2 / s2 / ay ** 2)) ** 2)

    def lam(w):
        s2 = 1 - w[2] ** 2
        a2 = (w[0] ** 2 * ax ** 2 + w[1] ** 2 * ay ** 2) / s2
        return (-1 + np.sqrt(1 + a2 * s2 / w[2] ** 2)) / 2

    def fr_diel(c, ei, et):
        st = ei / et * np.sqrt(max(0, 1 - c * c))
        ct = np.sqrt(max(0, 1 - st * st))
        rl = (et * c - ei * ct) / (et * c + ei * ct)
        rp = (ei * c - et * ct) / (ei * c + et * ct)
        return (rl * rl + rp * rp) / 2

    lum = col @ [0.212671, 0.715160, 0.072169]
    ctint = col / lum
    r0 = ((eta - 1) / (eta + 1)) ** 2
    cspec0 = (1 - metallic) * ((1 - tint) * np.ones(3) + tint * ctint) * r0 + metallic * col
    ch = wi @ wh
    fres = (1 - metallic) * fr_diel(ch, 1.0, eta) + metallic * ((1 - sw(ch)) * cspec0 + sw(ch))
    G = 1 / (1 + lam(wo)) / (1 + lam(wi))          # separable, not 1 / (1 + lam + lam)
    spec = D(wh) * G * fres / (4 * wo[2] * wi[2])
    assert np.allclose(out[:3], diffuse + retro + spec, rtol=1e-12)
    pdf_spec = D(wh) / (1 + lam(wo)) * abs(wo @ wh) / wo[2] / (4 * (wo @ wh))
    assert np.isclose(out[3], (2 * wi[2] / np.pi + pdf_spec) / 3, rtol=1e-12)
    # sheen adds c_sheen * sheen * diffuse_weight * schlick(cos_d)
    with_sheen = _probe(S.material_row(dict(base, sheen="sheen", sheen_tint="stint"), tex), wo, wi, (0.1, 0.6))
    csheen = (1 - 0.3) * np.ones(3) + 0.3 * ctint
    assert np.allclose(with_sheen[:3] - out[:3], csheen * 0.6 * dw * sw(cd), rtol=1e-9)
    # clearcoat: in the lobe count and the pdf, never in f (Q37); pdf through gtr1 with log10 (Q36)
    cc = _probe(S.material_row(dict(base, clearcoat="cc", clearcoat_gloss="ccg"), tex), wo, wi, (0.9, 0.6))
    assert np.array_equal(cc[:3], out[:3])
    gloss = 0.1 * (1 - 0.7) + 0.001 * 0.7
    a2 = gloss * gloss
    gtr1 = (a2 - 1) / (np.pi * np.log10(a2) * (1 + (a2 - 1) * wh[2] ** 2))
    pdf_cc = gtr1 * wh[2] / (4 * (wo @ wh))
    assert np.isclose(cc[3], (2 * wi[2] / np.pi + pdf_spec + pdf_cc) / 4, rtol=1e-12)
    assert cc[11] == 12                                     # u0 = 0.9 of four lobes: the clearcoat, DIFFUSE | GLOSSY
    swi = cc[7:10]
    swh = (wo + swi) / np.linalg.norm(wo + swi)
    gs = lambda c: 1 / (c + np.sqrt(0.0625 + c * c - 0.0625 * c * c))
    f_cc = 0.8 * gs(wo[2]) * gs(swi[2]) * (0.04 * (1 - sw(wo @ swh)) + sw(wo @ swh)) * \
        ((a2 - 1) / (np.pi * np.log10(a2) * (1 + (a2 - 1) * swh[2] ** 2))) / 4
    assert np.allclose(cc[4:7], f_cc, rtol=1e-9)            # sample_f hands back the chosen lobe's f alone (Q15)
    # ---- Disney, thin, everything on: the eight lobes in the order disney.rs adds them
    thin = dict(base, thin=True, sheen="sheen", clearcoat="cc", spec_trans="strans", flatness="flat", diff_trans="dt")
    rowt = S.material_row(thin, tex)
    kinds = [int(_probe(rowt, wo, wi, ((k + 0.5) / 8, 0.4))[11]) for k in range(8)]
    assert kinds == [5, 5, 5, 5, 9, 12, 10, 6]
    below = _probe(rowt, wo, wt, (0.99, 0.4))
    dwt = (1 - metallic) * (1 - 0.5)
    assert np.isclose(below[4] , col[0] * 0.6 / np.pi, rtol=1e-14)   # LambertianTransmission(c * diff_trans)
    flat_f = _probe(rowt, wo, wi, (0.01, 0.4))
    assert np.allclose(flat_f[4:7] / (col * dwt * (1 - 0.4) * (1 - 0.6) / np.pi),
                       (1 - sw(wo[2]) / 2) * (1 - sw(flat_f[9]) / 2), rtol=1e-12)
    # ---- Debug: two constant lobes, the "specular" one cosine-sampled like the diffuse one
    row = S.material_row({"material_type": "Debug"}, S.Textures({}))
    out = _probe(row, wo, wi, (0.75, 0.6))
    assert out[:3].tolist() == [0.0, 1.0, 1.0] and np.isclose(out[3], wi[2] / np.pi)
    assert out[4:7].tolist() == [0.0, 0.0, 1.0] and out[11] == 17 and np.isclose(out[10], out[9] / np.pi / 2)


def test_rough_glass_known_answers():
    """GlassMaterial with non-zero roughness (glass.rs:77-108): MicrofacetReflection + MicrofacetTransmission
    (reflection.rs:1028-1146) over an anisotropic Trowbridge-Reitz distribution, against the textbook formulas
    written out independently in numpy."""
    tex = S.Textures({"float_texture": [{"texture_name": "ur", "texture_type": "BilerpTexture", "v00": 0.2, "v01": 0.2},
                                        {"texture_name": "vr", "texture_type": "BilerpTexture", "v00": 0.35, "v01": 0.35}],
                      "rgb_texture": [{"texture_name": "black", "texture_type": "BilerpTexture", "v00": {"values": [0, 0, 0]},
                                       "v01": {"values": [0, 0, 0]}},
                                      {"texture_name": "kt", "texture_type": "BilerpTexture", "v00": {"values": [0.9, 0.8, 0.7]},
                                       "v01": {"values": [0.9, 0.8, 0.7]}}]})
    row = S.material_row({"material_type": "GlassMaterial", "kr": "black", "kt": "kt", "u_roughness": "ur", "v_roughness": "vr"}, tex)
    ax, ay, eta_b = 0.2, 0.35, 1.5
    wo = np.array([0.3, 0.2, np.sqrt(1 - 0.13)])
    wi = np.array([-0.25, 0.1, -np.sqrt(1 - 0.0725)])

    def D(wh):
        c2 = wh[2] ** 2
        t2 = (1 - c2) / c2
        s2 = max(1 - c2, 0.0)
        cp2, sp2 = (wh[0] ** 2 / s2, wh[1] ** 2 / s2) if s2 > 0 else (1.0, 0.0)
        return 1.0 / (np.pi * ax * ay * c2 * c2 * (1 + t2 * (cp2 / ax ** 2 + sp2 / ay ** 2)) ** 2)

    def lam(w):
        c2 = w[2] ** 2
        s2 = max(1 - c2, 0.0)
        t = np.sqrt(s2 / c2)
        cp2, sp2 = (w[0] ** 2 / s2, w[1] ** 2 / s2) if s2 > 0 else (1.0, 0.0)
        a = np.sqrt(cp2 * ax ** 2 + sp2 * ay ** 2)
        return (-1 + np.sqrt(1 + (a * t) ** 2)) / 2

    def fresnel(c, ei, et):
        c = np.clip(c, -1, 1)
        if c <= 0:
            ei, et, c = et, ei, abs(c)
        st = ei / et * np.sqrt(max(0, 1 - c * c))
        if st >= 1:
            return 1.0
        ct = np.sqrt(max(0, 1 - st * st))
        rl = (et * c - ei * ct) / (et * c + ei * ct)
        rp = (ei * c - et * ct) / (ei * c + et * ct)
        return (rl * rl + rp * rp) / 2

    eta = eta_b / 1.0                                   # wo above the surface
    wh = wo + wi * eta
    wh = wh / np.linalg.norm(wh)
    if wh[2] < 0:
        wh = -wh
    F = fresnel(wo @ wh, 1.0, eta_b)
    denom = wo @ wh + eta * (wi @ wh)
    G = 1 / (1 + lam(wo) + lam(wi))
    f_expect = (1 - F) * np.array([0.9, 0.8, 0.7]) * abs(D(wh) * G * eta * eta * abs(wi @ wh) * abs(wo @ wh) * (1 / eta) ** 2
                                                         / (wi[2] * wo[2] * denom * denom))
    whp = (wo + wi * eta) / np.linalg.norm(wo + wi * eta)   # pdf uses the unflipped half vector (reflection.rs:1138)
    pdf_expect = D(whp) * (1 / (1 + lam(wo))) * abs(wo @ whp) / abs(wo[2]) * abs(eta * eta * (wi @ whp) / denom ** 2)
    out = _probe(row, wo, wi, (0.37, 0.61), allow=False)
    assert np.allclose(out[:3], f_expect, rtol=1e-12) and f_expect.min() > 0
    assert np.isclose(out[3], pdf_expect, rtol=1e-12)
    # the sampled direction is the refraction of wo about a visible normal: below the surface, and the generalised half
    # vector of (wo, wi) is that normal again, so f / pdf are the values above evaluated at the sample
    swi = out[7:10]
    assert out[11] == 10 and swi[2] < 0 and np.isclose(np.linalg.norm(swi), 1.0)      # GLOSSY | TRANSMISSION
    again = _probe(row, wo, swi, (0.37, 0.61), allow=False)
    assert np.allclose(again[:3], out[4:7], rtol=1e-12) and np.isclose(again[3], out[10], rtol=1e-12)
    # with kr too there are two lobes: the glossy reflection half is MicrofacetReflection with a dielectric Fresnel term
    row2 = S.material_row({"material_type": "GlassMaterial", "u_roughness": "ur", "v_roughness": "vr"}, tex)
    wr = np.array([-0.1, 0.25, np.sqrt(1 - 0.0725)])
    o2 = _probe(row2, wo, wr, (0.2, 0.6), allow=True)                                 # allow_multiple_lobes is ignored when rough
    h = (wo + wr) / np.linalg.norm(wo + wr)
    fr = fresnel(wr @ h, 1.0, 1.5) * D(h) / (1 + lam(wo) + lam(wr)) / (4 * wr[2] * wo[2])
    assert np.allclose(o2[:3], fr, rtol=1e-12)
    assert o2[11] in (9, 10)


def test_film_weights_and_q10_q14(tmp_path):
    """nsamp = N renders N-1 samples (Q10); every sample, vignetted or not, adds its filter weight
    three times (Q14); a box filter of radius 0.5 puts each sample in its own pixel."""
    path = synth.scene_c1(str(tmp_path / "c1"), xres=64, yres=36, nsamp=4)
    r = S.load(path).render(seed=1, want_dump=True)
    assert r["dump"].shape[0] == 64 * 36 * 3
    assert np.array_equal(r["raw"][..., 3], np.full((36, 64), 9.0))
    assert r["stats"]["camera_rays"] + r["stats"]["zero_weight"] == 64 * 36 * 3
    r1 = S.load(path, {"Sampler": {"sampler_type": "HaltonSampler", "nsamp": 1}}).render(seed=1)
    assert r1["raw"].sum() == 0.0                                       # nsamp = 1 renders nothing


def test_oracle_render_is_deterministic_and_thread_independent(tmp_path):
    path = synth.scene_c4(str(tmp_path / "c4"), n_spheres=300, xres=64, yres=36, nsamp=5, extent=6.0)
    a = S.load(path).render(seed=5, nthreads=1)
    b = S.load(path).render(seed=5, nthreads=4)
    assert np.array_equal(a["rgb"], b["rgb"]) and a["stats"] == b["stats"]
    halves = [S.load(path).render(seed=5, tile_mod=2, tile_rank=r)["raw"] for r in (0, 1)]
    assert np.array_equal(halves[0] + halves[1], a["raw"])


def test_product_loader_matches_oracle_loader(tmp_path):
    """The C++ loader (rrt_scene_json_probe, host only) and the Python restatement read the same
    counts and the same render description from the same files."""
    for path in (synth.scene_c1(str(tmp_path / "c1"), nsamp=7),
                 synth.scene_c2(str(tmp_path / "c2"), n_instances=50, xres=80, yres=60),
                 synth.scene_c4(str(tmp_path / "c4"), n_spheres=200, xres=80, yres=60, nsamp=3, extra_materials=True)):
        info, d = render.json_probe(path)
        ls = S.load(path)
        prm, lens = S.render_params(ls.cfg)
        assert info["prims"] == ls.scene.num_prims
        assert info["materials"] == ls.materials.shape[0] and info["lights"] == ls.lights.shape[0]
        assert info["lens_values"] == lens.shape[0]
        assert (d.xres, d.yres, d.diagonal_mm, d.filter_kind) == (prm[0], prm[1], prm[2], prm[3])
        assert list(d.filter_radius) == [prm[4], prm[5]] and d.scale == prm[7]
        assert list(d.cam_pos) == list(prm[9:12]) and list(d.cam_look) == list(prm[12:15]) and list(d.cam_up) == list(prm[15:18])
        assert (d.aperture_diameter, d.focus_distance, d.nsamp) == (prm[20], prm[21], prm[23])
        assert (d.integrator_kind, d.max_depth, d.rr_threshold) == (prm[26], prm[27], prm[28])
    # the sample scene as shipped: Debug integrator (always "all" lights), StratifiedSampler defaults, the Debug material
    # (declared, named by no primitive) and the unreferenced ImageTexture skipped by both loaders
    shipped = synth.scene_c1_as_shipped(str(tmp_path / "c1s"))
    info, d = render.json_probe(shipped)
    prm, _ = S.render_params(S.load(shipped).cfg)
    assert (d.sampler_kind, d.strat_xsamp, d.strat_ysamp, d.strat_dimension, d.strat_jitter) == (1, 4, 4, 4, 1) == \
        (prm[38], prm[40], prm[41], prm[42], prm[39])
    assert (d.integrator_kind, d.max_depth, d.light_strategy, d.nsamp) == (2, 5, 1, 16) and prm[26] == 2 and prm[23] == 16
    assert info["materials"] == 4 == S.load(shipped).materials.shape[0] and info["lights"] == 3
    _, d = render.json_probe(shipped, {"Sampler": {"sampler_type": "StratifiedSampler", "xsamp": 3, "ysamp": 5, "dimension": 2, "jitter": False}})
    assert (d.strat_xsamp, d.strat_ysamp, d.strat_dimension, d.strat_jitter, d.nsamp) == (3, 5, 2, 0, 15)
    ov = {"Integrator": {"integrator_type": "Path", "max_depth": 3, "rr_threshold": 0.5}}
    _, d = render.json_probe(synth.scene_c2(str(tmp_path / "c2b"), n_instances=10, xres=32, yres=32), ov)
    assert (d.integrator_kind, d.max_depth, d.rr_threshold) == (0, 3, 0.5)


def test_write_image_quantisation_and_png(tmp_path):
    """renderprocess.rs:1501-1530: sRGB gamma (misc.rs:46-52), clamp(255 g + 0.5) as u8, alpha 255 —
    the host code of the product against a numpy restatement, and the PNG it writes decoded back."""
    import struct
    import zlib
    rng = np.random.default_rng(4)
    img = rng.uniform(-0.1, 1.3, (37, 53, 3))
    img[0, 0] = (0.0031308, 0.0, 1.0)
    img[0, 1] = (np.nan, 0.5, 2.0)
    g = np.where(img <= 0.0031308, 12.92 * img, 1.055 * np.power(np.maximum(img, 0), 1.0 / 2.4) - 0.055)
    ref = np.clip(255.0 * g + 0.5, 0.0, 255.0)
    ref = np.where(np.isnan(ref), 0.0, ref).astype(np.uint8)
    path = tmp_path / "out.png"
    got = render.rgb_to_png(img, path)
    assert np.array_equal(got[..., :3], ref) and (got[..., 3] == 255).all()
    data = path.read_bytes()
    assert data[:8] == bytes([0x89, 0x50, 0x4E, 0x47, 0x0D, 0x0A, 0x1A, 0x0A])
    pos, idat, ihdr = 8, b"", None
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == (zlib.crc32(typ + body) & 0xFFFFFFFF)
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        if typ == b"IDAT":
            idat += body
        pos += 12 + n
    assert ihdr == (53, 37, 8, 6, 0, 0, 0)
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(37, 1 + 53 * 4)
    assert (raw[:, 0] == 0).all() and np.array_equal(raw[:, 1:].reshape(37, 53, 4), got)


def test_area_light_sampling_known_answers():
    """DiffuseAreaLight::sample_li over Shape::sample_ref (diffuse.rs:62-79, shape/mod.rs:33-48), by hand:
    sphere r=1 at (0,0,5), u=(0.5,0): uniform_sample_sphere -> (1,0,0); p=(1,0,5), n=(1,0,0); from the origin
    w=(1,0,5), |w|^2=26, pdf = 26 / |cos| = 26*sqrt(26) (Q28: no 1/area), and the sample faces away -> L = 0."""
    import ctypes as C
    L = O.lib()
    L.orc_area_light_probe.argtypes = [C.c_void_p] * 4
    row = S.light_row({"light_type": "diffuse", "spectrum": {"values": [2, 3, 4]},
                       "light_shape": {"shape_type": "sphere", "radius": 1.0, "world_pos": [0, 0, 5]}})
    out = np.zeros(10)
    ref, u = np.zeros(3), np.array([0.5, 0.0])
    L.orc_area_light_probe(row.ctypes.data, ref.ctypes.data, u.ctypes.data, out.ctypes.data)
    assert np.allclose(out[4:7], [1, 0, 5], atol=1e-15)
    assert np.isclose(out[3], 26 * np.sqrt(26), rtol=1e-14)
    assert np.allclose(out[0:3], np.array([1, 0, 5]) / np.sqrt(26), rtol=1e-15)
    assert (out[7:10] == 0).all()
    # u=(0.5, 0.5): phi = pi -> p_obj = (-1, 0, 0), seen from (-4,0,5) it faces the reference point -> lemit
    ref = np.array([-4.0, 0.0, 5.0])
    u = np.array([0.5, 0.5])
    L.orc_area_light_probe(row.ctypes.data, ref.ctypes.data, u.ctypes.data, out.ctypes.data)
    assert np.allclose(out[4:7], [-1, 0, 5], atol=1e-15) and np.isclose(out[3], 9.0, rtol=1e-14)
    assert np.allclose(out[7:10], [2, 3, 4])
    # triangle, Q20: the "barycentrics" are a point of the unit sphere: u=(0,0) -> b=(0,0,1) -> p = p2
    tri = {"p": np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0.0]]), "vi": np.array([[0, 1, 2]]), "n": None, "ni": None}
    row = S.light_row({"light_type": "diffuse", "light_shape": {"shape_type": "triangle", "obj_name": "m", "tri_num": 0}}, {"m": tri})
    ref = np.array([0.0, 1.0, 2.0])
    u = np.array([0.0, 0.0])
    L.orc_area_light_probe(row.ctypes.data, ref.ctypes.data, u.ctypes.data, out.ctypes.data)
    assert np.allclose(out[4:7], [0, 1, 0]) and np.isclose(out[3], 4.0 / 1.0) and np.allclose(out[7:10], [1, 1, 1])


def test_area_light_scene_renders_and_the_bsdf_half_is_dead(tmp_path):
    """Two DiffuseAreaLights + a point light: the image is lit by all three, and estimate_direct's BSDF-sampling
    half traces rays (counted) that add nothing — with it switched off by hand the film would be identical; here
    we check the counters and that radiance arrives from the area lights alone."""
    path = synth.scene_area_lights(str(tmp_path / "a"), xres=96, yres=54, nsamp=5)
    r = S.load(path).render(seed=1)
    st = r["stats"]
    assert st["mis_probe_rays"] > 0 and st["shadow_rays"] > 0
    assert r["rgb"].max() > 0
    import json
    cfg = json.loads(open(path).read())
    only_area = {"lights": cfg["lights"][:2]}
    r2 = S.load(path, only_area).render(seed=1)
    assert r2["rgb"].max() > 0 and not np.array_equal(r2["rgb"], r["rgb"])
    # Tier L renders too (Q9 shadow rays towards the sampled point)
    r3 = S.load(path, tier=O.TIER_L).render(seed=1)
    assert np.isfinite(r3["rgb"]).all()


def test_halton_tables_are_the_digit_loops():
    """The device draws Halton samples through per-dimension tables (1 / base, the scrambled tail term, an exact
    multiply-shift division) and per-pixel index terms (csrc/halton.cuh).  On the host the same code must give the
    generic digit loops' bits for every dimension, with 32- and 64-bit indices, on the films of configs 1, 4 and 5."""
    import ctypes as C
    from rs_ray_toy_b200 import capi
    L = capi.lib()
    L.rrt_halton_host_probe.restype = C.c_int
    L.rrt_halton_host_probe.argtypes = [C.c_int64, C.c_int64, C.c_uint64, C.c_int, C.c_uint64] + [C.c_void_p] * 6
    rng = np.random.default_rng(3)
    for (xres, yres, spp) in ((640, 360, 17), (1920, 1080, 65), (3840, 2160, 257)):
        n = 20000
        px = rng.integers(-2, xres + 2, n).astype(np.int64)
        py = rng.integers(-2, yres + 2, n).astype(np.int64)
        sm = rng.integers(0, spp, n).astype(np.uint64)
        sm[:50] = 1 << 40                       # indices beyond 32 bits take the 64-bit digit loop first
        dim = rng.integers(0, 128, n).astype(np.uint32)
        dim[:128] = np.arange(128)
        outs = []
        for tables in (0, 1):
            idx, val = np.zeros(n, dtype=np.uint64), np.zeros(n)
            capi.check(L.rrt_halton_host_probe(xres, yres, 7, tables, n, px.ctypes.data, py.ctypes.data, sm.ctypes.data,
                                               dim.ctypes.data, idx.ctypes.data, val.ctypes.data))
            outs.append((idx, val))
        assert np.array_equal(outs[0][0], outs[1][0])
        assert np.array_equal(outs[0][1], outs[1][1])
        assert 0.0 <= outs[1][1].min() and outs[1][1].max() < 1.0 and len(np.unique(outs[1][1])) > n // 2


def test_sample_all_lights_is_one_estimate_per_light(tmp_path):
    """DirectLighting with light_strategy "all" (directlighting.rs:102-110, uniform_sample_all_lights
    integrator/mod.rs:304-355).  The sample arrays are requested on a throwaway sampler (Q30), so every light gets one
    estimate from two get_2d draws and the estimates are summed: the image is an unbiased estimate of the same
    integral as strategy "one" (which picks one light and multiplies by their number), with one shadow ray per light
    and lit hit instead of one per hit."""
    path = synth.scene_c1(str(tmp_path), xres=160, yres=90, nsamp=17, integrator="DirectLighting", max_depth=1)
    one = S.load(path).render(seed=1)
    every = S.load(path, {"Integrator": {"integrator_type": "DirectLighting", "max_depth": 1, "light_strategy": "all"}}).render(seed=1)
    assert every["stats"]["extension_rays"] == one["stats"]["extension_rays"]
    assert every["stats"]["shadow_rays"] > 2 * one["stats"]["shadow_rays"]        # three lights
    m1, m2 = one["rgb"].mean(axis=(0, 1)), every["rgb"].mean(axis=(0, 1))
    assert np.allclose(m1, m2, rtol=0.05), (m1, m2)
    assert not np.allclose(one["rgb"], every["rgb"])


def test_bump_map_known_cases(tmp_path):
    """Material::bump (material/mod.rs:22-65) on the sample scene's cubes (per-face vertex normals: dndu = dndv = 0).
    A constant displacement leaves the shading tangents as they were — dpdu + n (d - d) / du + dndu d — and the
    normal becomes normalize(ss x ts) instead of the face normal the triangle code had put there (interaction.rs:186-202,
    authoritative orientation): the same direction to rounding, so the image moves by rounding only; a ramp tilts every
    normal and changes it."""
    path = synth.scene_c1(str(tmp_path), xres=96, yres=54, nsamp=5)
    mats = lambda extra: [{"material_type": "MetalMaterial", "material_name": "mat_metal"},
                          {"material_type": "PlasticMaterial", "material_name": "mat_plastic"},
                          dict({"material_type": "MatteMaterial", "material_name": "mat_matte"}, **extra)]
    plain = S.load(path, {"materials": mats({})}).render(seed=1)
    ftex = [{"texture_name": "flat", "texture_type": "BilerpTexture", "v00": 0.3, "v01": 0.3},
            {"texture_name": "ramp", "texture_type": "BilerpTexture", "v00": 0.0},          # corners 0 1 0 1: d = t
            {"texture_name": "tilt", "texture_type": "ScaleTexture", "t1": "ramp", "t2": "flat"}]
    flat = S.load(path, {"float_texture": ftex, "materials": mats({"bump_map": "flat"})}).render(seed=1)
    rel = np.sqrt(np.mean((plain["rgb"] - flat["rgb"]) ** 2)) / np.sqrt(np.mean(plain["rgb"] ** 2))
    assert rel < 1e-9, rel
    tilt = S.load(path, {"float_texture": ftex, "materials": mats({"bump_map": "tilt"})}).render(seed=1)
    assert not np.allclose(plain["rgb"], tilt["rgb"])
    assert tilt["stats"]["camera_rays"] == plain["stats"]["camera_rays"]
