"""ORACLE (test infrastructure, CPU) -- the v4 fit oracle over a process pool.

Not part of the product (see oracle/fit_oracle.py).  The reference's iter_fit_seed_points visits the seeds
one after the other (External/Fitting_v4.py:606-683), but two seeds influence each other only if their
windows can interact: through the nearest-seed rule of firstfit (seeds within 2 (r + sqrt 3)) or through
im_subtr / im_add (overlapping windows, a subset of that).  The connected components of that relation
are independent problems; inside a component the reference's order is kept.  Each component is fitted by
fit_oracle.iter_fit on a crop of the stack, in the coordinates of the full stack and with the
nearest-seed tree of ALL seeds, so every number is the one the sequential loop produces (asserted by
tests/test_oracle_pinned.py::test_parallel_oracle_equals_sequential).  n_iter of the whole problem is the
largest n_iter of a component (converged seeds are frozen, so the global loop just keeps visiting the rest).
"""
import multiprocessing as mp
import os

import numpy as np
from scipy.spatial import cKDTree

from . import fit_oracle

_G = {}


def components(centers_nx3, radius_fit=5):
    """labels of the connected components of "centres closer than 2 (r + sqrt 3)" (int(c) decides the window, the
    float decides the distance: use the larger reach for both)"""
    cen = np.asarray(centers_nx3, dtype=np.float64).reshape(-1, 3)
    n = len(cen)
    parent = np.arange(n)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    reach = 2.0 * (radius_fit + np.sqrt(3.0)) + 1.0 + 1e-6        # + 1: windows sit at int(c)
    for a, b in cKDTree(cen).query_pairs(reach):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    return np.array([find(i) for i in range(n)])


def _fit_group(idx_lists):
    im, cen, tree, kw = _G["im"], _G["cen"], _G["tree"], _G["kw"]
    r = kw.get("radius_fit", 5)
    out = []
    for idx in idx_lists:
        idx = np.asarray(idx)
        c = cen[idx]
        lo = np.maximum(np.floor(c.min(0)).astype(np.int64) - r - 1, 0)
        hi = np.minimum(np.floor(c.max(0)).astype(np.int64) + r + 2, np.array(im.shape))
        crop = im[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
        res = fit_oracle.iter_fit(crop, c.T, version=4, origin=lo, full_shape=im.shape, tree=tree, global_index=idx, **kw)
        out.append((idx, {k: res[k] for k in ("ps", "success", "converged", "n_iter", "dists", "first_ps", "first_success", "nfev_first",
                                             "nfev_max", "cond_max", "well_posed", "unstable", "comparable")}))
    return out


def iter_fit_parallel(im, centers_3xn, procs=None, **kw):
    """fit_oracle.iter_fit(im, centers_3xn, version=4, **kw) over a process pool -> the same dict (without the
    float64 volumes im_subtr / im_add)."""
    cen = np.asarray(centers_3xn, dtype=np.float64).T
    n = len(cen)
    if n == 0:
        return fit_oracle.iter_fit(im, centers_3xn, version=4, **kw)
    procs = procs or min(32, os.cpu_count() or 1)
    lab = components(cen, kw.get("radius_fit", 5))
    order = np.argsort(lab, kind="stable")
    groups = np.split(order, np.nonzero(np.diff(lab[order]))[0] + 1)          # ascending seed index inside a group
    # balance: big components first, round robin over ~8 chunks per process
    groups.sort(key=len, reverse=True)
    n_chunks = max(1, min(len(groups), procs * 8))
    chunks = [groups[i::n_chunks] for i in range(n_chunks)]
    _G.update(im=im, cen=cen, tree=cKDTree(cen), kw=kw)
    try:
        if procs <= 1:
            parts = [_fit_group(ch) for ch in chunks]
        else:
            with mp.get_context("fork").Pool(procs) as pool:
                parts = pool.map(_fit_group, chunks, chunksize=1)
    finally:
        _G.clear()
    ps, first_ps = [None] * n, [None] * n
    arr = {k: np.zeros(n, dtype=t) for k, t in (("success", bool), ("converged", bool), ("dists", float), ("first_success", bool),
                                                ("nfev_first", np.int64), ("nfev_max", np.int64), ("cond_max", float),
                                                ("well_posed", bool), ("unstable", bool), ("comparable", bool))}
    n_iter = 0
    for part in parts:
        for idx, res in part:
            n_iter = max(n_iter, res["n_iter"])
            for j, i in enumerate(idx):
                ps[i], first_ps[i] = res["ps"][j], res["first_ps"][j]
            for k in arr:
                arr[k][idx] = np.asarray(res[k])
    # the reference recomputes dists for every seed after each sweep: a group that was done before the last global
    # sweep has old == new there, i.e. 0
    for part in parts:
        for idx, res in part:
            if res["n_iter"] < n_iter:
                arr["dists"][idx] = 0.0
    out = dict(ps=ps, first_ps=first_ps, n_iter=n_iter, n_components=len(groups))
    out.update(arr)
    out["success"] = list(out["success"])
    return out
