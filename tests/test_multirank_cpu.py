"""world_size-2 gloo test of the multi-rank render path on CPU: the tile deal and the film reduce are
host logic; the per-rank renderer here is the CPU oracle (no GPU in this suite)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, scene_path, out_dir, port):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import oracle_scene as S
    from rs_ray_toy_b200 import parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = S.load(scene_path).render(seed=2, nthreads=2, tile_mod=world, tile_rank=rank)
    mine = parallel.tiles_for_rank(96, 64, world, rank)
    # every pixel this rank sampled lies in one of its tiles
    mask = np.zeros((64, 96), dtype=bool)
    for _, (x0, y0, x1, y1) in mine:
        mask[max(y0, 0):y1, max(x0, 0):x1] = True
    assert (r["raw"][..., 3][~mask] == 0).all() and (r["raw"][..., 3][mask] > 0).all()
    total = parallel.reduce_sums(r["raw"].copy(), dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), total.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tile_deal_and_film_reduce(tmp_path):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_scene as S
    from rs_ray_toy_b200 import parallel, synth
    path = synth.scene_c4(str(tmp_path / "c4"), n_spheres=300, xres=96, yres=64, nsamp=4, extent=6.0)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, path, str(tmp_path), port), nprocs=2, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    full = S.load(path).render(seed=2)["raw"]
    assert np.array_equal(reduced, full)   # box filter 0.5: one owner per pixel -> bit-identical
    # the deal covers every tile exactly once
    ids = sorted(t for r in range(2) for t, _ in parallel.tiles_for_rank(96, 64, 2, r))
    assert ids == list(range(len(ids))) and len(ids) == 6 * 4
