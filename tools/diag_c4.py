"""Developer diagnostic (GPU box): full-size C4 against the parallel oracle; which comparable rows leave the
tolerance, their reference conditioning, and whether the library variant matters (IA3_LIB).
    python tools/diag_c4.py oracle   -> gpurun_out/c4_oracle.npz
    IA3_LIB=... python tools/diag_c4.py device <tag>  -> gpurun_out/c4_dev_<tag>.npz
    python tools/diag_c4.py compare <tag> [<tag2>]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageanalysis3_b200.synth import synth            # noqa: E402

mode = sys.argv[1]
OUT = "gpurun_out"
shape, n, seed, hr = (60, 2048, 2048), 50000, 4, (400.0, 3000.0)
if len(sys.argv) > 3 and sys.argv[3] == "small":
    shape, n = (60, 1024, 1024), 12500
if mode == "oracle":
    from oracle import seed_oracle
    im = synth(shape, n, seed, h_range=hr)
    seeds = seed_oracle.get_seeds_oracle(im, th_seed=300.0, backend="c")
    t0 = time.perf_counter()
    from oracle import parallel
    res = parallel.iter_fit_parallel(im, seeds.T, procs=min(32, os.cpu_count()))
    print(f"oracle: {len(seeds)} seeds, {time.perf_counter() - t0:.1f} s, n_iter {res['n_iter']}")
    np.savez(os.path.join(OUT, "c4_oracle.npz"), seeds=seeds, ps=np.array([np.asarray(r, dtype=np.float64) for r in res["ps"]]),
             comparable=res["comparable"], cond_max=res["cond_max"], nfev_max=res["nfev_max"], well=res["well_posed"], unstable=res["unstable"],
             converged=res["converged"])
elif mode == "device":
    from imageanalysis3_b200 import _lib
    from imageanalysis3_b200.External import Fitting_v4
    _lib.init(0)
    im = synth(shape, n, seed, h_range=hr)
    o = np.load(os.path.join(OUT, "c4_oracle.npz"))
    f = Fitting_v4.iter_fit_seed_points(im, o["seeds"].T)
    t0 = time.perf_counter()
    f._fit_all()
    print(f"device {sys.argv[2]}: {1e3 * (time.perf_counter() - t0):.0f} ms, n_iter {f.n_iter}, stats {f._h.engine_stats()}")
    np.savez(os.path.join(OUT, f"c4_dev_{sys.argv[2]}.npz"), ps=f._ps_array().astype(np.float64), nfev=f.nfev, n_visits=f._h.n_visits, converged=f.converged)
else:
    o = np.load(os.path.join(OUT, "c4_oracle.npz"))
    w, ok = o["ps"], o["comparable"]
    from scipy.spatial import cKDTree
    tree = cKDTree(o["seeds"])
    for tag in sys.argv[2:]:
        if tag == "small":
            continue
        d = np.load(os.path.join(OUT, f"c4_dev_{tag}.npz"))
        g = d["ps"]
        dc = np.abs(g[:, 1:4] - w[:, 1:4]).max(1)
        rel = (np.abs(g[:, [0, 5, 6, 7]] - w[:, [0, 5, 6, 7]]) / np.abs(w[:, [0, 5, 6, 7]])).max(1)
        bad = ok & ((dc > 1e-3) | (rel > 1e-4))
        print(f"{tag}: comparable {ok.sum()}, out of tolerance {bad.sum()}; max centre dev {dc[ok].max():.2e}, max rel {rel[ok].max():.2e}; "
              f"rows with any difference at all among comparable: {(ok & ((dc > 0) | (rel > 1e-6))).sum()}")
        for i in np.nonzero(bad)[0][:40]:
            nb = [j for j in tree.query_ball_point(o["seeds"][i], 12.0) if j != i]
            print(f"  row {i}: dc {dc[i]:.2e} rel {rel[i]:.2e} | ref cond {o['cond_max'][i]:.1f} nfev_max {o['nfev_max'][i]} height {w[i, 0]:.0f} sig {w[i, 5:8].round(2)} "
                  f"| dev nfev {d['nfev'][i]} visits {d['n_visits'][i]} | neighbours {nb[:6]} comparable {[bool(ok[j]) for j in nb[:6]]} "
                  f"nb cond {[round(float(o['cond_max'][j]), 1) for j in nb[:6]]} nb unstable {[bool(o['unstable'][j]) for j in nb[:6]]}")
    if len(sys.argv) > 3 and sys.argv[3] != "small":
        a, b = np.load(os.path.join(OUT, f"c4_dev_{sys.argv[2]}.npz"))["ps"], np.load(os.path.join(OUT, f"c4_dev_{sys.argv[3]}.npz"))["ps"]
        print("variants differ on rows:", (np.abs(a - b).max(1) > 0).sum(), "max centre diff", np.nanmax(np.abs(a[:, 1:4] - b[:, 1:4])))
