"""Host-side mirror of `Transform` (src/transform.rs:180-351) for assembling instance blocks.

Each constructor returns `(m, m_inv)` exactly as the reference's `Transform{m, m_inv}` carries
them: analytic inverses for translate / scale / rotate, `m_inv = b.inv * a.inv` for products
(transform.rs:441-449), and sums in the reference's left-to-right order so the matrices that
reach the GPU aggregate are the same f64 bits the reference would hold.
"""
from __future__ import annotations

import math

import numpy as np


def identity():
    return np.eye(4), np.eye(4)


def translate(delta):
    """transform.rs:254-265"""
    m, inv = np.eye(4), np.eye(4)
    m[0:3, 3] = delta
    inv[0:3, 3] = -np.asarray(delta, dtype=np.float64)
    return m, inv


def scale(x, y, z):
    """transform.rs:266-290"""
    return np.diag([x, y, z, 1.0]).astype(np.float64), np.diag([1.0 / x, 1.0 / y, 1.0 / z, 1.0])


def _normalize(v):
    """Vector3f::normalize (geometry.rs:925-931): the zero vector is returned unchanged."""
    v = np.asarray(v, dtype=np.float64)
    l = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    return v if l == 0.0 else v / l


def rotate(theta_deg, axis):
    """transform.rs:327-351 — rotation by theta (degrees) about `axis`."""
    a = _normalize(axis)
    rad = (math.pi / 180.0) * theta_deg  # misc.rs:56-58
    s, c = math.sin(rad), math.cos(rad)
    ax, ay, az = float(a[0]), float(a[1]), float(a[2])
    m = np.eye(4)
    m[0, 0] = ax * ax + (1.0 - ax * ax) * c
    m[0, 1] = ax * ay * (1.0 - c) - az * s
    m[0, 2] = ax * az * (1.0 - c) + ay * s
    m[1, 0] = ax * ay * (1.0 - c) + az * s
    m[1, 1] = ay * ay + (1.0 - ay * ay) * c
    m[1, 2] = ay * az * (1.0 - c) - ax * s
    m[2, 0] = ax * az * (1.0 - c) - ay * s
    m[2, 1] = ay * az * (1.0 - c) + ax * s
    m[2, 2] = az * az + (1.0 - az * az) * c
    return m, m.T.copy()


def _mul44(a, b):
    """Matrix4x4::mul (transform.rs:138-177): r[i][j] = a[i][0]*b[0][j] + ... in index order."""
    r = np.empty((4, 4))
    for i in range(4):
        for j in range(4):
            r[i, j] = a[i, 0] * b[0, j] + a[i, 1] * b[1, j] + a[i, 2] * b[2, j] + a[i, 3] * b[3, j]
    return r


def mul(t1, t2):
    """Transform * Transform (transform.rs:441-449)."""
    return _mul44(t1[0], t2[0]), _mul44(t2[1], t1[1])


def make_to_world(world_pos=(0.0, 0.0, 0.0), rotation_axis=(0.0, 0.0, 0.0), rotation_angle=0.0, scale_xyz=(1.0, 1.0, 1.0)):
    """`make_to_world` (src/renderprocess.rs:242-252): translate * rotate(angle, axis) * scale."""
    # the loader normalises the axis (renderprocess.rs:238-239) and Transform::rotate normalises it again
    return mul(mul(translate(world_pos), rotate(rotation_angle, _normalize(rotation_axis))), scale(*scale_xyz))


def make_to_world_batch(world_pos, rotation_axis, rotation_angle):
    """Vectorised `make_to_world` with unit scale for large instance blocks (configs 2 and 4).

    Returns (m[n,4,4], m_inv[n,4,4]); same arithmetic order as the scalar path.
    """
    pos = np.asarray(world_pos, dtype=np.float64)
    axis = np.asarray(rotation_axis, dtype=np.float64)
    ang = np.asarray(rotation_angle, dtype=np.float64)
    n = pos.shape[0]
    a = axis
    for _ in range(2):  # normalised by the loader and again by Transform::rotate
        l = np.sqrt(a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1] + a[:, 2] * a[:, 2])
        safe = np.where(l == 0.0, 1.0, l)
        a = np.where((l == 0.0)[:, None], a, a / safe[:, None])
    rad = (math.pi / 180.0) * ang
    s, c = np.sin(rad), np.cos(rad)
    ax, ay, az = a[:, 0], a[:, 1], a[:, 2]
    R = np.zeros((n, 4, 4))
    R[:, 3, 3] = 1.0
    R[:, 0, 0] = ax * ax + (1.0 - ax * ax) * c
    R[:, 0, 1] = ax * ay * (1.0 - c) - az * s
    R[:, 0, 2] = ax * az * (1.0 - c) + ay * s
    R[:, 1, 0] = ax * ay * (1.0 - c) + az * s
    R[:, 1, 1] = ay * ay + (1.0 - ay * ay) * c
    R[:, 1, 2] = ay * az * (1.0 - c) - ax * s
    R[:, 2, 0] = ax * az * (1.0 - c) - ay * s
    R[:, 2, 1] = ay * az * (1.0 - c) + ax * s
    R[:, 2, 2] = az * az + (1.0 - az * az) * c
    Rinv = np.transpose(R, (0, 2, 1)).copy()
    T = np.zeros((n, 4, 4))
    Tinv = np.zeros((n, 4, 4))
    for k in range(4):
        T[:, k, k] = 1.0
        Tinv[:, k, k] = 1.0
    T[:, 0:3, 3] = pos
    Tinv[:, 0:3, 3] = -pos
    S = np.broadcast_to(np.eye(4), (n, 4, 4))

    def mm(a_, b_):
        r = np.empty((n, 4, 4))
        for i in range(4):
            for j in range(4):
                r[:, i, j] = (a_[:, i, 0] * b_[:, 0, j] + a_[:, i, 1] * b_[:, 1, j] + a_[:, i, 2] * b_[:, 2, j]
                              + a_[:, i, 3] * b_[:, 3, j])
        return r

    m = mm(mm(T, R), S)
    inv = mm(S, mm(Rinv, Tinv))
    return m, inv
