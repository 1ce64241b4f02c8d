"""GPU: per-sweep device time and MINPACK effort of the fit stage on one synthetic stack.
    python tools/profile_fit.py [Z X Y n_planted]
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from imageanalysis3_b200 import _lib
from imageanalysis3_b200.External import Fitting_v4
from imageanalysis3_b200.spot_tools import fitting
from imageanalysis3_b200.synth import synth_torch

shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (50, 2048, 2048)
n_planted = int(sys.argv[4]) if len(sys.argv) >= 5 else 5000
_lib.init(0)
d = synth_torch(shape, n_planted, 1, torch.device("cuda", 0))
host = d.cpu().numpy().view(np.uint16)
st = _lib.Stack(device_ptr=d.data_ptr(), shape=shape, dtype=np.uint16)
t0 = time.perf_counter()
seeds = fitting.get_seeds(host, max_num_seeds=None, th_seed=300.0, _stack=st)
print(f"seeds {len(seeds)} in {1e3*(time.perf_counter()-t0):.1f} ms (wall)")
f = Fitting_v4.iter_fit_seed_points(host, seeds.T, _stack=st)
t0 = time.perf_counter()
f.firstfit()
print(f"firstfit wall {1e3*(time.perf_counter()-t0):.1f} ms device {f._h.last_ms:.2f} ms levels {f._h.num_levels} ties {f.n_tie_voxels}")
print("  nfev pct 50/90/99/max", np.percentile(f.nfev, [50, 90, 99, 100]), "info", np.bincount(f.info))
h = f._h
n = len(seeds)
conv = np.zeros(n, bool)
ps_old = h.ps.copy()
for it in range(12):
    t0 = time.perf_counter()
    h.repeat_sweep(2.5, ~conv)
    wall = 1e3 * (time.perf_counter() - t0)
    act = ~conv
    nf = h.nfev[act]
    dist = ((h.ps[:, 1:4] - ps_old[:, 1:4]) ** 2).sum(1)
    conv = dist < 0.01
    ps_old = h.ps.copy()
    print(f"sweep {it}: active {act.sum()} wall {wall:.1f} ms device {h.last_ms:.2f} ms nfev 50/90/99/max "
          f"{np.percentile(nf, [50, 90, 99, 100])} info {np.bincount(h.info[act])} -> unconverged {(~conv).sum()}")
    if conv.all():
        break
