"""GPU: run one BASELINE.json config end to end and print what happened (sizes, sweeps, time, memory).
    python tools/run_config.py C4      # 60x2048x2048, 50 000 planted spots
    python tools/run_config.py C1      # 30x512x512, 500 planted spots
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from imageanalysis3_b200 import _lib
from imageanalysis3_b200.spot_tools import fitting
from imageanalysis3_b200.synth import synth_torch

CFG = {"C1": ((30, 512, 512), 500, 0, (600., 4000.)), "C2": ((50, 2048, 2048), 5000, 1, (600., 4000.)),
       "C4": ((60, 2048, 2048), 50000, 4, (400., 3000.))}
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
shape, n, seed, hr = CFG[name]
_lib.init(0)
d = synth_torch(shape, n, seed, torch.device("cuda", 0), h_range=hr)
h = torch.empty(shape, dtype=torch.int16, pin_memory=True)
h.copy_(d)
del d
torch.cuda.empty_cache()
im = h.numpy().view(np.uint16)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    spots = fitting.fit_fov_image(im, '647', th_seed=300, max_num_seeds=None, verbose=False)
    dt = time.perf_counter() - t0
    print(f"{name} rep {rep}: {len(spots)} spots in {1e3 * dt:.1f} ms ({len(spots) / dt:.0f} spots/s), dtype {spots.dtype}, "
          f"finite {np.isfinite(spots).all()}, mem {torch.cuda.mem_get_info()[0] / 2**30:.1f} GiB free")
print("height quantiles", np.percentile(spots[:, 0], [5, 50, 95]), "sigma_x median", np.median(spots[:, 6]))
