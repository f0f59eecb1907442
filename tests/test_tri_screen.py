"""The closest-hit kernel's fp32 triangle screen (csrc/tri_screen.h) run on the host through rrt_tri_screen_host_probe:
a screened-out candidate must be one the f64 Moller-Trumbore test (triangle.rs:233-265, restated below in numpy f64 with
the reference's operation order) turns down as well — or one whose hit lies beyond best_t.  The screen may abstain as
often as it likes; it must never reject a candidate the deciding arithmetic would keep."""
import ctypes as C

import numpy as np

from rs_ray_toy_b200 import capi


def screen(o, d, best_t, verts):
    L = capi.lib()
    n = len(o)
    o = np.ascontiguousarray(o, dtype=np.float64)
    d = np.ascontiguousarray(d, dtype=np.float64)
    bt = np.ascontiguousarray(best_t, dtype=np.float64)
    v = np.ascontiguousarray(verts, dtype=np.float32).reshape(n, 9)
    out = np.zeros(n, dtype=np.uint8)
    capi.check(L.rrt_tri_screen_host_probe(n, o.ctypes.data, d.ctypes.data, bt.ctypes.data, v.ctypes.data, out.ctypes.data))
    return out.astype(bool)


def mt_f64(o, d, p0, p1, p2):
    """tri_test of csrc/aggregate.cu == triangle.rs:233-265: accepted?, t."""
    with np.errstate(all="ignore"):
        e1, e2 = p1 - p0, p2 - p0
        P = np.cross(d, e2)
        a = (e1[:, 0] * P[:, 0] + e1[:, 1] * P[:, 1]) + e1[:, 2] * P[:, 2]
        ok = ~((a > -1e-7) & (a < 1e-7))
        f = 1.0 / a
        T = o - p0
        u = f * ((T[:, 0] * P[:, 0] + T[:, 1] * P[:, 1]) + T[:, 2] * P[:, 2])
        ok &= ~((u < 0.0) | (u > 1.0))
        Q = np.cross(T, e1)
        v = f * ((d[:, 0] * Q[:, 0] + d[:, 1] * Q[:, 1]) + d[:, 2] * Q[:, 2])
        ok &= ~((v < 0.0) | (u + v > 1.0))
        t = f * ((e2[:, 0] * Q[:, 0] + e2[:, 1] * Q[:, 1]) + e2[:, 2] * Q[:, 2])
        ok &= ~(t < 1e-7)
        ok &= ~np.isnan(t)
    return ok, t


def check(o, d, best_t, verts, min_reject=None):
    verts = np.asarray(verts, dtype=np.float32).reshape(-1, 3, 3)
    v64 = verts.astype(np.float64)
    rej = screen(o, d, best_t, verts)
    ok, t = mt_f64(np.asarray(o, dtype=np.float64), np.asarray(d, dtype=np.float64), v64[:, 0], v64[:, 1], v64[:, 2])
    kept_by_f64 = ok & ~(t > np.asarray(best_t))
    bad = rej & kept_by_f64
    assert not bad.any(), (int(bad.sum()), np.flatnonzero(bad)[:5])
    if min_reject is not None:
        assert rej.mean() >= min_reject, rej.mean()
    return rej, kept_by_f64


def soup(n, rng, edge=0.01, scale=1.0, offset=0.0):
    v0 = rng.uniform(0, 1, (n, 3)) * scale + offset
    v1 = v0 + rng.uniform(-edge, edge, (n, 3)) * scale
    v2 = v0 + rng.uniform(-edge, edge, (n, 3)) * scale
    return np.stack([v0, v1, v2], axis=1).astype(np.float32)


def test_random_rays_against_random_triangles():
    rng = np.random.default_rng(11)
    n = 400000
    tri = soup(n, rng)
    o = rng.uniform(0, 1, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rej, kept = check(o, d, np.full(n, np.inf), tri, min_reject=0.95)
    assert kept.sum() < 0.05 * n


def test_rays_aimed_at_triangles_edges_and_vertices():
    """Rays through points ON the triangle (interior, edges, vertices, a hair outside): the decisions the bands exist
    for.  Most interior / edge cases must abstain; none may be wrongly rejected."""
    rng = np.random.default_rng(12)
    n = 300000
    for scale, offset in ((1.0, 0.0), (1e-3, 0.0), (50.0, 200.0), (1.0, 4096.0)):
        tri = soup(n, rng, scale=scale, offset=offset).astype(np.float64)   # fp32-exact vertices, as PrimRec48 holds them
        b = rng.uniform(0, 1, (n, 2))
        kind = rng.integers(0, 5, n)
        b[kind == 1, 1] = 0.0                                   # on edge v = 0
        b[kind == 2] = (0.0, 0.0)                               # vertex p0
        flip = (b.sum(1) > 1) & (kind != 3)
        b[flip] = 1 - b[flip]
        b[kind == 3, 0] = 1.0 - b[kind == 3, 1]                 # on edge u + v = 1
        b[kind == 4] += rng.choice([-1, 1], (int((kind == 4).sum()), 2)) * 10.0 ** rng.uniform(-9, -3, (int((kind == 4).sum()), 2))
        target = tri[:, 0] + b[:, :1] * (tri[:, 1] - tri[:, 0]) + b[:, 1:] * (tri[:, 2] - tri[:, 0])
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        dist = 10.0 ** rng.uniform(-6, 1, (n, 1)) * scale
        o = target - d * dist
        # best_t: inf, exactly the hit distance (a tie must stay in), slightly nearer / farther
        bt = np.full(n, np.inf)
        sel = rng.integers(0, 4, n)
        bt[sel == 1] = dist[sel == 1, 0]
        bt[sel == 2] = dist[sel == 2, 0] * (1 - 10.0 ** rng.uniform(-12, -2, int((sel == 2).sum())))
        bt[sel == 3] = dist[sel == 3, 0] * (1 + 10.0 ** rng.uniform(-12, -2, int((sel == 3).sum())))
        rej, kept = check(o, d, bt, tri.astype(np.float32))
        if scale >= 1.0:   # (at scale 1e-3 the determinant is below the reference's absolute 1e-7: everything is rejected)
            assert kept.sum() > 0.3 * n, (scale, offset, kept.sum())    # hits by construction, up to the perturbations


def test_degenerate_and_extreme_inputs_abstain_or_agree():
    rng = np.random.default_rng(13)
    n = 50000
    tri = soup(n, rng)
    o = rng.uniform(0, 1, (n, 3))
    d = rng.normal(size=(n, 3))
    # grazing rays: direction in the triangle's plane (determinant ~ 0)
    e1 = (tri[:, 1] - tri[:, 0]).astype(np.float64)
    check(tri[:, 0] - 3 * e1 + 1e-9 * rng.normal(size=(n, 3)), e1 + 1e-9 * rng.normal(size=(n, 3)), np.full(n, np.inf), tri)
    # zero-area triangles, tiny and huge coordinates, huge / tiny directions, NaN / inf in the ray
    z = tri.copy()
    z[:, 2] = z[:, 1]
    check(o, d, np.full(n, np.inf), z)
    check(o * 1e-30, d, np.full(n, np.inf), (tri * np.float32(1e-30)))
    check(o * 1e20, d, np.full(n, np.inf), (tri * np.float32(1e20)))
    check(o, d * 1e25, np.full(n, np.inf), tri)
    check(o, d * 1e-25, np.full(n, np.inf), tri)
    bad = d.copy()
    bad[::3, 0] = np.nan
    bad[1::3, 1] = np.inf
    rej = screen(o, bad, np.full(n, np.inf), tri)
    assert not rej[::3].any()                 # a NaN never decides
    # origin exactly on a vertex, best_t = 0 and negative
    check(tri[:, 0].astype(np.float64), d, np.zeros(n), tri)
    check(o, d, np.full(n, -1.0), tri)


def test_bounce_rays_of_the_bench_generator_screen_well():
    """Config 3's own rays: nearly everything that reaches a leaf and misses is screened out in fp32."""
    from rs_ray_toy_b200 import synth
    p, idx = synth.soup_triangles(20000)
    rays = synth.bounce_rays(p, idx, 60000, seed=4)
    rng = np.random.default_rng(5)
    pick = rng.integers(0, len(idx), len(rays))
    tri = p[idx[pick]].astype(np.float32)
    o = rays[:, 0:3] if rays.ndim == 2 else np.stack([rays["o"]], 0)
    d = rays[:, 3:6]
    rej, kept = check(o, d, np.full(len(o), np.inf), tri, min_reject=0.98)
