"""Developer trace of the fit engine on one stack: per-round work (one-warp tasks, team tasks, device clock),
the seeds that sit on the critical path, and -- with the profiling build (IA3_LIB=.../libia3b200_prof.so,
built by `python -m imageanalysis3_b200.build --variant=prof -DIA3_FIT_PROF`) -- the cycles the team kernel
spends per function evaluation in each phase of lmder.

    python tools/engine_trace.py [c1|c2|c4crop] [v4|v3]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageanalysis3_b200 import _lib                                   # noqa: E402
from imageanalysis3_b200.External import Fitting_v3, Fitting_v4       # noqa: E402
from imageanalysis3_b200.spot_tools import fitting                     # noqa: E402
from imageanalysis3_b200.synth import synth                            # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
ver = sys.argv[2] if len(sys.argv) > 2 else "v4"
shape, n, seed, kw = {"c1": ((30, 512, 512), 500, 0, {}), "c2": ((50, 2048, 2048), 5000, 1, {}),
                      "c4crop": ((60, 512, 512), 3125, 4, dict(h_range=(400.0, 3000.0)))}[which]
_lib.init(0)
im = synth(shape, n, seed, **kw)
st = _lib.Stack(im)
seeds = fitting.get_seeds(im, max_num_seeds=None, th_seed=300.0, _stack=st)
mod = Fitting_v4 if ver == "v4" else Fitting_v3
for rep in range(2):
    f = mod.iter_fit_seed_points(im, seeds.T, _stack=st)
    t0 = time.perf_counter()
    f._fit_all()
    wall = time.perf_counter() - t0
s = f._h.engine_stats(trace=True)
tr = s.pop("trace")
print(f"{which} {ver}: {len(seeds)} seeds, wall {1e3 * wall:.1f} ms, device {f._h.last_ms:.2f} ms, n_iter {f.n_iter}")
print(s)
print("round: one-warp tasks, team tasks, t [ms]")
for r, (nb, nt, t) in enumerate(tr):
    print(f"  {r:3d}: {nb:6d} {nt:5d} {t:9.3f}")
h = f._h
order = np.argsort(-h.n_visits.astype(np.int64) * 100000 - h.nfev)[:25]
from scipy.spatial import cKDTree
tree = cKDTree(seeds)
print("seeds with the most visits / longest last run: index, visits, nfev of last run, info, seeds within 10 px, height")
for i in order:
    nb = len(tree.query_ball_point(seeds[i], 10.0)) - 1
    print(f"  {i:6d} {h.n_visits[i]:3d} {h.nfev[i]:5d} {h.info[i]:2d} {nb:3d} {h.ps[i, 0]:9.1f}")
print("nfev histogram of the last runs:", np.histogram(h.nfev, bins=[0, 8, 12, 24, 48, 100, 300, 999, 2000])[0])
