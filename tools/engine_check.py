"""Developer check of the fit engine on a GPU box: the device-resident run (firstfit + repeatfit in one go,
per-seed dataflow, memo, speculation) against the host-driven sweep loop on the same seeds, engine
counters, and wall-clock latency of one stack alone.

    python tools/engine_check.py [small|c1|c2]
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

which = sys.argv[1] if len(sys.argv) > 1 else "small"
shape, n, seed, kw = {"small": ((30, 128, 128), 400, 31, dict(h_range=(500.0, 3000.0))),
                      "c1": ((30, 512, 512), 500, 0, {}),
                      "c2": ((50, 2048, 2048), 5000, 1, {})}[which]
_lib.init(0)
im = synth(shape, n, seed, **kw)
seeds = fitting.get_seeds(im, max_num_seeds=None, th_seed=300.0 if which != "small" else 200.0)
print(f"{which}: {len(seeds)} seeds", flush=True)


def same(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


for name, mod in (("v4", Fitting_v4), ("v3", Fitting_v3)):
    st = _lib.Stack(im)
    t0 = time.perf_counter()
    fa = mod.iter_fit_seed_points(im, seeds.T, _stack=st)
    fa._fit_all()
    t_all = time.perf_counter() - t0
    print(f"{name} fit_all: {1e3 * t_all:.1f} ms wall, device {fa._h.last_ms:.2f} ms, n_iter {fa.n_iter}, stats {fa._h.engine_stats()}", flush=True)

    fb = mod.iter_fit_seed_points(im, seeds.T, _stack=st)
    t0 = time.perf_counter()
    fb.firstfit()
    t_first = time.perf_counter() - t0
    first_ps = fb._ps_array()
    t0 = time.perf_counter()
    fb.repeatfit()
    t_rep = time.perf_counter() - t0
    print(f"{name} firstfit {1e3 * t_first:.1f} ms + repeatfit {1e3 * t_rep:.1f} ms, n_iter {fb.n_iter}, stats {fb._h.engine_stats()}", flush=True)

    fc = mod.iter_fit_seed_points(im, seeds.T, _stack=st)
    fc.firstfit()
    t0 = time.perf_counter()
    fc._repeatfit_host_loop()
    t_loop = time.perf_counter() - t0
    print(f"{name} host-driven sweeps {1e3 * t_loop:.1f} ms, n_iter {fc.n_iter}, stats {fc._h.engine_stats()}", flush=True)

    ok = True
    for tag, f in (("two calls", fb), ("host loop", fc)):
        for attr in ("_ps", "_succ", "converged", "dists", "nfev", "info", "success_old", "centers_fit_old"):
            a, b = np.asarray(getattr(fa, attr)), np.asarray(getattr(f, attr))
            if not same(a, b):
                ok = False
                bad = np.nonzero(~np.isclose(a.astype(float), b.astype(float), rtol=0, atol=0, equal_nan=True).reshape(len(a), -1).all(1))[0] if a.shape == b.shape else []
                print(f"  MISMATCH {name} fit_all vs {tag}: {attr} rows {bad[:8]} ({len(bad)} of {len(a)})")
        if fa.n_iter != f.n_iter:
            ok = False
            print(f"  MISMATCH {name} n_iter {fa.n_iter} vs {tag} {f.n_iter}")
    print(f"{name}: fit_all == two calls == host loop: {ok}", flush=True)
    if name == "v4" and which != "c2":
        from oracle import fit_oracle
        t0 = time.perf_counter()
        o = fit_oracle.iter_fit(im, seeds.T, version=4)
        cmp_ok = o["comparable"]
        got = np.asarray(fa._ps_array(), dtype=np.float64)
        want = np.asarray([np.asarray(r, dtype=np.float64) for r in o["ps"]])
        dc = np.abs(got[cmp_ok, 1:4] - want[cmp_ok, 1:4]).max()
        rel = (np.abs(got[cmp_ok][:, [0, 5, 6, 7]] - want[cmp_ok][:, [0, 5, 6, 7]]) / np.abs(want[cmp_ok][:, [0, 5, 6, 7]])).max()
        print(f"  vs oracle ({time.perf_counter() - t0:.1f} s): comparable {cmp_ok.sum()}/{len(cmp_ok)}, max centre dev {dc:.2e}, max rel dev {rel:.2e}, "
              f"n_iter {fa.n_iter} vs {o['n_iter']}, converged equal {np.array_equal(fa.converged[cmp_ok], o['converged'][cmp_ok])}", flush=True)
    del fa, fb, fc, st

# latency of one stack alone through the public call
for rep in range(3):
    t0 = time.perf_counter()
    spots = fitting.fit_fov_image(im, '647', th_seed=300.0 if which != "small" else 200.0, max_num_seeds=None, verbose=False)
    print(f"fit_fov_image alone: {1e3 * (time.perf_counter() - t0):.1f} ms, {len(spots)} spots", flush=True)
