"""Synthetic workloads of BASELINE.json's configs (seeded, numpy only, no device code).

Scenes are returned as plain arrays in the reference's own terms — mesh vertices `p`
(f64), 0-based triangle vertex indices, instance transforms as (m, m_inv) 4x4 pairs —
so the same arrays feed both the CPU oracle (tests only) and the CUDA library.

Vertices are drawn in float32 and widened to float64: the reference computes in f64
(`src/geometry.rs:12-20`), and an fp32-exact vertex set lets the device keep 12-byte
vertices without changing a single input bit.  Rays stay full f64.
"""
from __future__ import annotations

import numpy as np

SEED_C2_INSTANCES = 2
SEED_C3_SOUP = 3
SEED_C3_RAYS = 4
SEED_C4_SPHERES = 5
SEED_C5_SOUP = 6


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def soup_triangles(n_tris: int, edge: float = 0.01, seed: int = SEED_C3_SOUP):
    """Config 3/5 'random-soup' mesh: v0 ~ U([0,1]^3), v1,v2 = v0 + U([-edge,edge]^3).

    Returns (p[3n,3] f64, idx[n,3] u32): one private vertex triple per triangle, the
    layout `create_triangle_mesh` (`src/shape/triangle.rs:131-165`) would get from an
    .obj without shared vertices.
    """
    rng = _rng(seed)
    v0 = rng.random((n_tris, 3), dtype=np.float32)
    d1 = (rng.random((n_tris, 3), dtype=np.float32) * 2 - 1) * np.float32(edge)
    d2 = (rng.random((n_tris, 3), dtype=np.float32) * 2 - 1) * np.float32(edge)
    p = np.empty((n_tris, 3, 3), dtype=np.float32)
    p[:, 0] = v0
    p[:, 1] = v0 + d1
    p[:, 2] = v0 + d2
    idx = np.arange(3 * n_tris, dtype=np.uint32).reshape(n_tris, 3)
    return p.reshape(-1, 3).astype(np.float64), idx


def _concentric_disk(u: np.ndarray) -> np.ndarray:
    """`concentric_sample_disk`, `src/sampling.rs:277-298` (vectorised)."""
    uo = 2.0 * u - 1.0
    x, y = uo[:, 0], uo[:, 1]
    use_x = np.abs(x) > np.abs(y)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(use_x, x, y)
        theta = np.where(use_x, (np.pi / 4) * (y / x), (np.pi / 2) - (np.pi / 4) * (x / y))
    theta = np.where((x == 0) & (y == 0), 0.0, theta)
    r = np.where((x == 0) & (y == 0), 0.0, r)
    return np.stack([r * np.cos(theta), r * np.sin(theta)], axis=1)


def bounce_rays(p: np.ndarray, idx: np.ndarray, n_rays: int, seed: int = SEED_C3_RAYS, offset: float = 1e-4):
    """Config 3 'incoherent diffuse bounce' rays.

    Pick a triangle uniformly, a uniform barycentric point on it, lift the origin by
    `offset` along the unit normal and draw a cosine-weighted direction about that normal
    (`cosine_sample_hemisphere`, `src/sampling.rs:265-274`).  t_max = +inf.
    Returns rays[n,7] f64 = (o, d, t_max) with |d| = 1 to f64 rounding.
    """
    rng = _rng(seed)
    tri = rng.integers(0, idx.shape[0], size=n_rays)
    v = p[idx[tri]]  # [n,3,3]
    b = rng.random((n_rays, 2))
    su = np.sqrt(b[:, 0])
    b0 = 1.0 - su
    b1 = b[:, 1] * su
    pt = b0[:, None] * v[:, 0] + b1[:, None] * v[:, 1] + (1.0 - b0 - b1)[:, None] * v[:, 2]
    n = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    ln[ln == 0] = 1.0
    n = n / ln
    flip = rng.random(n_rays) < 0.5
    n[flip] = -n[flip]
    d2 = _concentric_disk(rng.random((n_rays, 2)))
    z = np.sqrt(np.maximum(0.0, 1.0 - d2[:, 0] ** 2 - d2[:, 1] ** 2))
    # local frame about n (coordinate_system, src/geometry.rs:1146-1161)
    big_x = np.abs(n[:, 0]) > np.abs(n[:, 1])
    s = np.where(big_x[:, None],
                 np.stack([-n[:, 2], np.zeros(n_rays), n[:, 0]], axis=1),
                 np.stack([np.zeros(n_rays), n[:, 2], -n[:, 1]], axis=1))
    ls = np.linalg.norm(s, axis=1, keepdims=True)
    ls[ls == 0] = 1.0
    s = s / ls
    t = np.cross(n, s)
    d = d2[:, 0:1] * s + d2[:, 1:2] * t + z[:, None] * n
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.empty((n_rays, 7), dtype=np.float64)
    rays[:, 0:3] = pt + offset * n
    rays[:, 3:6] = d
    rays[:, 6] = np.inf
    return rays


def shadow_rays_from(rays: np.ndarray, light_pos, seed: int = 11):
    """Any-hit workload: from each ray origin toward a point light, t_max just short of it."""
    o = rays[:, 0:3]
    seg = np.asarray(light_pos, dtype=np.float64)[None, :] - o
    dist = np.linalg.norm(seg, axis=1, keepdims=True)
    out = np.empty_like(rays)
    out[:, 0:3] = o
    out[:, 3:6] = seg / dist
    out[:, 6] = dist[:, 0] * (1.0 - 1e-4)
    return out


def random_unit_vectors(n: int, rng: np.random.Generator) -> np.ndarray:
    z = 1.0 - 2.0 * rng.random(n)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = 2.0 * np.pi * rng.random(n)
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)


def instance_params(n: int, extent: float, seed: int, rotate: bool = True):
    """Config 2/4 instance blocks: world_pos ~ U([-extent,extent]^3), random axis/angle, scale 1.

    Positions are float32-exact.  Returns dict(world_pos[n,3], axis[n,3], angle[n]).
    """
    rng = _rng(seed)
    pos = ((rng.random((n, 3), dtype=np.float32) * 2 - 1) * np.float32(extent)).astype(np.float64)
    if rotate:
        axis = random_unit_vectors(n, rng)
        angle = rng.random(n) * 360.0
    else:
        axis = np.zeros((n, 3))
        angle = np.zeros(n)
    return {"world_pos": pos, "axis": axis, "angle": angle}


def camera_like_rays(n: int, eye, extent: float, seed: int = 12):
    """Rays from one eye point through uniformly random points of the cube [-extent,extent]^3."""
    rng = _rng(seed)
    tgt = (rng.random((n, 3)) * 2 - 1) * extent
    eye = np.asarray(eye, dtype=np.float64)
    d = tgt - eye[None, :]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.empty((n, 7))
    rays[:, 0:3] = eye
    rays[:, 3:6] = d
    rays[:, 6] = np.inf
    return rays


CUBE_P = np.array(
    [[1, 1, -1], [1, -1, -1], [1, 1, 1], [1, -1, 1], [-1, 1, -1], [-1, -1, -1], [-1, 1, 1], [-1, -1, 1]],
    dtype=np.float64,
)
# samples/cube.obj faces (1-based in the file; 0-based here), `v//vn`
CUBE_VI = np.array(
    [[4, 2, 0], [2, 7, 3], [6, 5, 7], [1, 7, 5], [0, 3, 1], [4, 1, 5],
     [4, 6, 2], [2, 6, 7], [6, 4, 5], [1, 3, 7], [0, 2, 3], [4, 0, 1]],
    dtype=np.uint32,
)
CUBE_N = np.array([[0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [1, 0, 0], [0, 0, -1]], dtype=np.float64)
CUBE_NI = np.repeat(np.array([0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5], dtype=np.uint32)[:, None], 3, axis=1)
