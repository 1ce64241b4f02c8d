"""Times drift estimation in its bead-fitting mode (correction_tools.alignment.align_image(use_autocorr=False)) at the
reference's stack size on the device, and the same crops through the CPU oracle (the reference's scipy path) beside it.
python tools/align_time.py [--cpu]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from imageanalysis3_b200 import _lib  # noqa: E402
from imageanalysis3_b200.correction_tools import alignment  # noqa: E402
from imageanalysis3_b200.synth import bead_pair  # noqa: E402


def main():
    shape = (30, 2048, 2048)
    drift = (0.6, -3.2, 4.7)
    ref, src, _ = bead_pair(shape, 4000, drift, 5)
    _lib.init()
    kw = dict(use_autocorr=False, correction_args=dict(single_im_size=list(shape)), verbose=False)
    for rep in range(3):
        t0 = time.perf_counter()
        d, flag = alignment.align_image(src, ref, **kw)
        print(f"rep {rep}: align_image on the device path {1e3 * (time.perf_counter() - t0):.1f} ms -> drift {np.round(d, 4)} flag {flag} (planted {drift})", flush=True)
    if "--cpu" in sys.argv:
        from oracle import fit_oracle
        import imageanalysis3_b200.spot_tools.fitting as fitting
        crops = alignment.generate_drift_crops(list(shape))
        t0 = time.perf_counter()
        drifts = []
        for crop in crops[:3]:
            s = tuple(slice(*r) for r in crop)
            cts = []
            for im in (src[s], ref[s]):
                spots, _ = fit_oracle.fit_fov_image_oracle(np.ascontiguousarray(im), **alignment._default_align_fitting_args)
                cts.append(fitting.select_sparse_centers(spots[:, 1:4], 2.))
            dft, _, _ = alignment.align_beads(cts[0], cts[1], src[s], ref[s], verbose=False)
            drifts.append(-dft)
        dt = time.perf_counter() - t0
        print(f"CPU path (oracle fit_fov_image + the same host pairing), three crops: {dt:.1f} s -> drift {np.round(np.mean(drifts, axis=0), 4)}; "
              f"device result differs by {np.abs(np.mean(drifts, axis=0) - d).max():.2e} px")


if __name__ == "__main__":
    main()
