"""One pass of every pre-processing kernel at the reference's stack size (for ncu): hot pixels, bleed-through +
illumination over three channels, chromatic + drift warp of one channel."""
import sys

import numpy as np

sys.path.insert(0, ".")
from imageanalysis3_b200 import _lib  # noqa: E402

rng = np.random.default_rng(0)
shape = (30, 2048, 2048)
ims = [rng.integers(200, 3000, size=shape, dtype=np.uint16) for _ in range(3)]
ims[0][:, 100, 200] = 60000
_lib.init()
stacks = [_lib.Stack(im) for im in ims]
print("hot columns", stacks[0].remove_hot_pixels())
illum = (0.7 + 0.6 * rng.random(shape[1:])).astype(np.float32)
bleed = (np.eye(3)[0][:, None, None] + 0.05 * rng.random((3,) + shape[1:])).astype(np.float32)
mixed = _lib.Stack.mix(stacks, bleed=bleed, illum=illum)
chrom = (rng.standard_normal((3, 1) + shape[1:]) * 0.5).astype(np.float32)
out = mixed.warp(drift=[0.4, -1.3, 2.2], chroma=chrom)
print("checksum", int(out.fetch().astype(np.uint64).sum()))
