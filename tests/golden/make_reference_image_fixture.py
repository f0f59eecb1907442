"""Writes tests/golden/reference_scene_png.npz from /root/reference/samples/scene.png — the one image the reference
repository holds that the reference itself rendered (640 x 360 RGBA, 8 bit).  Run in the build container (the
reference is not present on the GPU box); the fixture is what travels.

Kept: the coloured-pixel mask (bit-packed), the mean colour of every 8 x 8 block (uint8), and the summary numbers
`write_image` (renderprocess.rs:1501-1530) would have printed for it.  tests/test_reference_image.py reads it."""
from pathlib import Path

import numpy as np
from PIL import Image

SRC = Path("/root/reference/samples/scene.png")
OUT = Path(__file__).resolve().parent / "reference_scene_png.npz"


def main():
    im = np.array(Image.open(SRC))
    assert im.shape == (360, 640, 4) and (im[..., 3] == 255).all()
    rgb = im[..., :3]
    mask = rgb.astype(int).sum(-1) > 0
    blocks = rgb.reshape(45, 8, 80, 8, 3).astype(np.float64).mean(axis=(1, 3))
    np.savez_compressed(OUT, mask=np.packbits(mask), block_mean=np.round(blocks).astype(np.uint8),
                        coloured=np.int64(mask.sum()), distinct=np.int64(len(np.unique(rgb.reshape(-1, 3), axis=0))),
                        mean_rgb_coloured=rgb[mask].astype(np.float64).mean(0),
                        frac_blue_below_red=np.float64((rgb[mask][:, 2] < rgb[mask][:, 0]).mean()))
    print(OUT, OUT.stat().st_size, "bytes; coloured", int(mask.sum()))


if __name__ == "__main__":
    main()
