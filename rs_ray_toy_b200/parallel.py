"""Multi-GPU rendering: the reference's tile parallelism (rayon over 16x16 sample tiles,
src/integrator/mod.rs:55-71) dealt to one process per GPU.

No data moves while rendering: rank r owns the tiles t with t % world == r (tile index =
tile_y * n_tiles_x + tile_x over the film's SAMPLE bounds), every rank holds the whole scene, and the
Halton index of a sample depends only on (pixel, sample number).  The single exchange is the
`merge_film_tile` of src/film.rs:248-263 across ranks.  With the default box filter (radius 0.5) a tile's
samples land in the tile's own pixels, so the frame is a GATHER: every rank packs the pixels of its tiles
(rrt_render_pack_owned, 1 / G of the film) and rank 0 unpacks them (`gather_film`, NCCL gather over NVLink).
Wider filters splat across tile borders: there the 4-f64-per-pixel accumulation films are summed onto rank 0
(`reduce_film`).  One rank builds the tree and the others replicate it (`commit_replicated`).
"""
from __future__ import annotations

import numpy as np


def sample_bounds(xres: int, yres: int, rx: float = 0.5, ry: float = 0.5):
    """Film::get_sample_bounds (film.rs:188-199) for the full-frame crop window."""
    import math
    x0, y0 = math.floor(0 + 0.5 - rx), math.floor(0 + 0.5 - ry)
    x1, y1 = math.ceil(xres - 0.5 + rx), math.ceil(yres - 0.5 + ry)
    return int(x0), int(y0), int(x1), int(y1)


def tiles_for_rank(xres: int, yres: int, world: int, rank: int, rx: float = 0.5, ry: float = 0.5, tile: int = 16):
    """Tile ids (and their sample rectangles) that `rank` renders."""
    x0, y0, x1, y1 = sample_bounds(xres, yres, rx, ry)
    ntx, nty = (x1 - x0 + tile - 1) // tile, (y1 - y0 + tile - 1) // tile
    out = []
    for t in range(ntx * nty):
        if t % world == rank:
            tx, ty = t % ntx, t // ntx
            out.append((t, (x0 + tx * tile, y0 + ty * tile, min(x0 + (tx + 1) * tile, x1), min(y0 + (ty + 1) * tile, y1))))
    return out


def reduce_sums(local, dst: int = 0):
    """Sum of the ranks' accumulation films on `dst` (torch tensor or numpy array, any backend)."""
    import torch
    import torch.distributed as dist
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    return t


def reduce_film(render, dst: int = 0):
    """Rank `dst` ends up holding the whole frame in its renderer's film; returns the film tensor.

    The film is staged through a torch CUDA tensor only because torch.distributed owns the NCCL
    communicator; the copies are device-to-device."""
    import torch
    import torch.distributed as dist
    _, n = render.film_device()
    buf = torch.empty(n, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    render.film_copy(buf.data_ptr(), False, stream)
    reduce_sums(buf, dst)
    if not dist.is_initialized() or dist.get_rank() == dst:
        render.film_copy(buf.data_ptr(), True, stream)
    torch.cuda.current_stream().synchronize()
    return buf


def gather_film(render, world: int, rank: int, dst: int = 0):
    """Rank `dst` ends up holding the whole frame: every other rank ships the pixels of its own tiles only.
    Falls back to the full-film sum when the filter is wider than a pixel (tiles do not own their pixels then)."""
    import torch
    import torch.distributed as dist
    from . import capi
    if world == 1:
        return
    try:
        sizes = [render.owned_doubles(world, r) for r in range(world)]
        n = max(sizes)
        buf = torch.zeros(n, dtype=torch.float64, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        if rank != dst:
            render.pack_owned(world, rank, buf.data_ptr(), n, stream)
    except capi.RrtError as e:
        if e.status != capi.RRT_ERR_UNSUPPORTED:
            raise
        reduce_film(render, dst)
        return
    parts = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(world)] if rank == dst else None
    dist.gather(buf, parts, dst=dst)
    if rank == dst:
        for r in range(world):
            if r != dst and sizes[r]:
                render.unpack_owned(world, r, parts[r].data_ptr(), n, stream)
    torch.cuda.current_stream().synchronize()


def commit_replicated(agg, max_prims_in_node: int = 4, src: int = 0):
    """BVHAccel::new on ONE rank: `src` builds the tree, every rank ends up with a committed aggregate holding the same
    nodes and records (rrt_scene_export_tree -> broadcast -> rrt_scene_commit_from_tree).  All ranks must have made the
    same rrt_scene_add_* calls."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return agg.commit(max_prims_in_node)
    rank = dist.get_rank()
    cuda = dist.get_backend() == "nccl"
    dev = "cuda" if cuda else "cpu"
    if rank == src:
        agg.commit(max_prims_in_node)
        blob = torch.from_numpy(agg.export_tree())
        size = torch.tensor([blob.numel()], dtype=torch.int64, device=dev)
        dist.broadcast(size, src=src)
        t = blob.to(dev)
        dist.broadcast(t, src=src)
        return agg
    size = torch.zeros(1, dtype=torch.int64, device=dev)
    dist.broadcast(size, src=src)
    t = torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return agg.commit_from_tree(t.cpu().numpy())
