"""Pairing of fitted centres between two images (reference spot_tools/matching.py:148-287): small host bookkeeping on a
few hundred points, kept in numpy / scipy like the reference."""
import numpy as np


def find_paired_centers(tar_cts, ref_cts, drift=None, cutoff=2, dimension=3,
                        return_paired_cts=True, return_kept_inds=False, verbose=False):
    """Centres of ``tar_cts`` and ``ref_cts`` (+ rough drift) that are each other's ONLY partner within ``cutoff``;
    returns the mean of their differences [+ the paired centres] [+ their indices]"""
    from scipy.spatial.distance import cdist
    dim = int(dimension)
    tar, ref = np.array(tar_cts), np.array(ref_cts)
    if np.shape(tar)[1] > 3:
        tar = tar[:, 1:1 + dim]
    if np.shape(ref)[1] > 3:
        ref = ref[:, 1:1 + dim]
    shift = np.zeros(np.shape(tar)[1]) if drift is None else np.array(drift, dtype=float)[:dim]
    if verbose:
        print(f"-- aligning {len(tar)} centers to {len(ref)} ref_centers, given drift:{np.round(shift, 2)}", end=', ')
    near = cdist(tar, ref + shift) <= cutoff
    it, ir = np.where(near)
    one_t, one_r = near.sum(axis=1) == 1, near.sum(axis=0) == 1
    pairs = [[t, r] for t, r in zip(it, ir) if one_t[t] and one_r[r]]
    p_tar = np.array([tar[t] for t, _ in pairs])
    p_ref = np.array([ref[r] for _, r in pairs])
    new_drift = np.nanmean(p_tar - p_ref, axis=0)
    if verbose:
        print(f"{len(p_tar)} pairs found, updated_drift:{np.round(new_drift, 2)}")
    ret = [new_drift]
    if return_paired_cts:
        ret += [p_tar, p_ref]
    if return_kept_inds:
        idx = np.array(pairs, dtype=int)
        ret += [idx[:, 0], idx[:, 1]]
    return tuple(ret)


def check_paired_centers(paired_tar_cts, paired_ref_cts, outlier_sigma=1.5, return_paired_cts=True, verbose=False):
    """Drops pairs whose shift differs from the inverse-distance weighted shift of their Delaunay neighbours by more
    than mean + outlier_sigma * std of those differences; returns the mean shift of the kept pairs [+ the pairs]"""
    from scipy.spatial import Delaunay
    tar, ref = np.array(paired_tar_cts, dtype=float), np.array(paired_ref_cts, dtype=float)
    shifts = tar - ref
    if verbose:
        print(f"-- check {len(tar)} pairs of centers", end=', ')
    simplices = Delaunay(ref).simplices
    expected = []
    for i, rc in enumerate(ref):
        nb = np.unique(simplices[(simplices == i).any(axis=1)])
        nb = nb[(nb != i) & (nb != -1)]
        w = 1 / np.linalg.norm(ref[nb] - rc, axis=1)
        expected.append(np.dot(shifts[nb].T, w) / np.sum(w))
    diffs = np.linalg.norm(np.array(expected) - shifts, axis=1)
    keep = np.array(diffs < np.mean(diffs) + np.std(diffs) * outlier_sigma)
    new_drift = np.nanmean(tar[keep] - ref[keep], axis=0)
    if verbose:
        print(f"{int(keep.sum())} pairs kept. new drift:{np.round(new_drift, 2)}")
    ret = [new_drift]
    if return_paired_cts:
        ret += [tar[keep], ref[keep]]
    return tuple(ret)
