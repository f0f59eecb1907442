"""CPU checks of the device LBVH builder's per-element code (csrc/lbvh_core.h), run on the host through
rrt_lbvh_host_probe: the emitted tree must be a valid BVH over exactly the input primitives.  The kernels in
bvh_lbvh.cu call the same functions; their parity with the oracle is tested on the GPU (test_gpu_parity.py)."""
import numpy as np
import pytest

from rs_ray_toy_b200 import synth
from rs_ray_toy_b200.aggregate import lbvh_host_probe


def tri_bounds(n, edge=0.01, seed=11, scale=1.0, offset=0.0):
    p, idx = synth.soup_triangles(n, edge, seed)
    v = p[idx] * scale + offset          # [n, 3, 3]
    return np.concatenate([v.min(axis=1), v.max(axis=1)], axis=1)


def walk(words, order, bounds, max_leaf):
    """Returns (leaf runs, max depth); asserts containment on the way."""
    planes = words[:, :12].view(np.float32)
    child = words[:, 12:14].view(np.int32)
    n = bounds.shape[0]
    seen = np.zeros(n, dtype=np.int64)
    runs, max_depth = [], 0
    stack = [(0, 1, None)]
    while stack:
        node, depth, parent_box = stack.pop()
        max_depth = max(max_depth, depth)
        pl = planes[node]
        boxes = [np.array([pl[0], pl[2], pl[8], pl[1], pl[3], pl[9]], dtype=np.float64),
                 np.array([pl[4], pl[6], pl[10], pl[5], pl[7], pl[11]], dtype=np.float64)]
        for c in range(2):
            box = boxes[c]
            assert (box[:3] <= box[3:]).all()
            if parent_box is not None:   # children lie inside the box their parent was given
                assert (box[:3] >= parent_box[:3] - 1e-12).all() and (box[3:] <= parent_box[3:] + 1e-12).all()
            ref = int(child[node, c])
            if ref >= 0:
                assert ref > node or ref < len(words)
                stack.append((ref, depth + 1, box))
            else:
                r = ~ref & 0xFFFFFFFF
                first, cnt = r >> 3, (r & 7) + 1
                assert cnt <= max_leaf
                runs.append((first, cnt))
                ids = order[first:first + cnt]
                seen[ids] += 1
                b = bounds[ids]
                assert (b[:, :3] >= box[:3]).all() and (b[:, 3:] <= box[3:]).all()
    return runs, max_depth, seen


@pytest.mark.parametrize("n,max_leaf", [(17, 4), (1000, 4), (1000, 1), (5000, 8), (20000, 4)])
def test_radix_tree_is_a_valid_bvh(n, max_leaf):
    bounds = tri_bounds(n)
    words, order, info = lbvh_host_probe(bounds, max_leaf)
    assert sorted(order.tolist()) == list(range(n))
    assert info["nodes"] == len(words)
    runs, depth, seen = walk(words, order, bounds, max_leaf)
    assert (seen == 1).all()                       # every primitive in exactly one leaf
    runs.sort()
    pos = 0
    for first, cnt in runs:                        # leaves are consecutive runs of the sorted order
        assert first == pos
        pos += cnt
    assert pos == n
    assert len(runs) == info["leaves"]
    assert depth == info["max_depth"]
    assert len(words) == len(runs) - 1             # binary tree


def test_duplicate_centroids_and_flat_scenes():
    """Equal Morton keys (identical boxes) are split by position; a scene flat in one axis still builds."""
    b = np.tile(np.array([[0.1, 0.2, 0.3, 0.4, 0.5, 0.6]]), (300, 1))
    words, order, info = lbvh_host_probe(b, 4)
    runs, depth, seen = walk(words, order, b, 4)
    assert (seen == 1).all() and depth <= 16
    flat = tri_bounds(2000)
    flat[:, 2] = 0.0
    flat[:, 5] = 0.0
    words, order, info = lbvh_host_probe(flat, 4)
    runs, depth, seen = walk(words, order, flat, 4)
    assert (seen == 1).all()


def test_morton_order_keeps_neighbours_together():
    """Sanity of the 63-bit Morton key: leaves of a uniform soup are spatially small."""
    bounds = tri_bounds(20000)
    words, order, info = lbvh_host_probe(bounds, 4)
    c = 0.5 * (bounds[:, :3] + bounds[:, 3:])[order]
    d = np.linalg.norm(c[1:] - c[:-1], axis=1)
    assert np.median(d) < 0.05                     # a random order would give ~0.66


def test_too_few_primitives_is_refused():
    from rs_ray_toy_b200 import capi
    with pytest.raises(capi.RrtError):
        lbvh_host_probe(tri_bounds(3), 4)
