"""GPU: per-row deviations of the device fit from the oracle on a synthetic stack, with the oracle's
conditioning / MINPACK effort of the same rows (to tell numerical noise from real differences)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from imageanalysis3_b200 import _lib
from imageanalysis3_b200.External import Fitting_v4
from imageanalysis3_b200.synth import synth
from oracle import fit_oracle, seed_oracle

_lib.init(0)
im = synth((60, 256, 256), 780, 4, h_range=(400.0, 3000.0))
seeds = seed_oracle.get_seeds_oracle(im, th_seed=300.0, backend="c")
o = fit_oracle.iter_fit(im, seeds.T, version=4)
f = Fitting_v4.iter_fit_seed_points(im, seeds.T)
f.firstfit()
first = f._ps.copy()
nf_first = f.nfev.copy()
f.repeatfit()
got = np.array([np.asarray(r, float) for r in f.ps])
want = np.array([np.asarray(r, float) for r in o["ps"]])
w1 = np.array([np.asarray(r, float) for r in o["first_ps"]])
rel = np.abs(got[:, 5:8] - want[:, 5:8]) / np.abs(want[:, 5:8])
dc = np.abs(got[:, 1:4] - want[:, 1:4]).max(1)
rel1 = (np.abs(first[:, 5:8] - w1[:, 5:8]) / np.abs(w1[:, 5:8])).max(1)
bad = np.nonzero((rel.max(1) > 5e-5) | (dc > 3e-4))[0]
print("n", len(seeds), "n_iter dev/oracle", f.n_iter, o["n_iter"], "comparable", o["comparable"].mean())
for i in bad:
    print(f"seed {i:4d} dc {dc[i]:.2e} sig rel {rel[i].max():.2e} | firstfit rel {rel1[i]:.2e} nfev first dev/ref {nf_first[i]}/{o['nfev_first'][i]} "
          f"| cond_max {o['cond_max'][i]:.3g} nfev_max {o['nfev_max'][i]} comparable {o['comparable'][i]} sigmas {want[i,5:8]}")
