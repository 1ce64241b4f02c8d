"""Times the device corrections at the reference's stack size (30 x 2048 x 2048, three channels) and the numpy / scipy
path (the oracle, test infrastructure) on one channel beside it.  python tools/corr_time.py [--cpu]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from imageanalysis3_b200 import _lib  # noqa: E402
from imageanalysis3_b200.io_tools import load  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    shape = (30, 2048, 2048)
    chs = ['750', '647', '561']
    ims = [rng.integers(200, 3000, size=shape, dtype=np.uint16) for _ in chs]
    for im in ims:
        im[:, 100, 200] = 60000
    illum = {ch: (0.7 + 0.6 * rng.random(shape[1:])).astype(np.float32) for ch in chs}
    bleed = (np.eye(3)[:, :, None, None] + 0.05 * rng.random((3, 3) + shape[1:])).astype(np.float32)
    # what correction_tools/chromatic.py saves: float64, one plane per z (3 GB per channel at this size)
    chrom = {ch: (rng.standard_normal((3,) + shape) * 0.5) if ch != '647' else None for ch in chs}
    drift = [0.4, -1.3, 2.2]
    _lib.init()
    for rep in range(3):
        t0 = time.perf_counter()
        stacks = [_lib.Stack(im) for im in ims]
        t1 = time.perf_counter()
        for s in stacks:
            s.remove_hot_pixels()
        t2 = time.perf_counter()
        mixed = [_lib.Stack.mix(stacks, bleed=load.resident_profile(bleed), bleed_row=i, illum=load.resident_profile(illum[ch])) for i, ch in enumerate(chs)]
        t3 = time.perf_counter()
        warped = [s.warp(drift=drift, chroma=load.resident_profile(chrom[ch])) for s, ch in zip(mixed, chs)]
        t4 = time.perf_counter()
        outs = [s.fetch() for s in warped]
        t5 = time.perf_counter()
        print(f"rep {rep}: upload {1e3*(t1-t0):.1f} ms, hot pixels {1e3*(t2-t1):.1f}, bleed+illumination {1e3*(t3-t2):.1f}, "
              f"warp {1e3*(t4-t3):.1f}, fetch {1e3*(t5-t4):.1f}; three channels total {1e3*(t5-t0):.1f} ms", flush=True)
        del stacks, mixed, warped
    t0 = time.perf_counter()
    got = load.correct_image_stacks(ims, chs, chs, chs, drift=drift, illumination_profile=illum, bleed_profile=bleed, chromatic_profile=chrom)
    print(f"correct_image_stacks (3 channels, host arrays in and out): {1e3*(time.perf_counter()-t0):.1f} ms")
    if "--cpu" in sys.argv:
        from oracle import correct_oracle
        t0 = time.perf_counter()
        want = correct_oracle.correct_stacks(ims, chs, ['750'], chs, drift=drift, illumination_profile=illum, bleed_profile=bleed, chromatic_profile=chrom)
        dt = time.perf_counter() - t0
        d = np.abs(want[0].astype(np.int64) - got[0].astype(np.int64))
        print(f"numpy/scipy path, ONE selected channel (bleed-through over three): {dt:.1f} s; device vs it: max diff {d.max()}, differing voxels {(d>0).sum()} of {d.size}")


if __name__ == "__main__":
    main()
