"""Drop-in for the fitting part of the reference's ``External/Fitting_v4.py``
(GaussianFit :165-396, in_dim :399-401, closest_faster :422-424, iter_fit_seed_points :559-683).
Same class names, constructor arguments, methods and attributes; the fits run on the B200."""
import numpy as np

from ._iterfit import GaussianFitBase, IterFitBase, in_dim, window_offsets  # noqa: F401


class GaussianFit(GaussianFitBase):
    _personality = 4

    def __init__(self, im, X, center=None, n_aprox=10, min_w=0.5, max_w=4., delta_center=3.,
                 init_w=1.5):
        self._setup(im, X, center, n_aprox, min_w, max_w, delta_center, init_w, 0.0)


def closest_faster(xyz, ic, tree, rsearch=6):
    """voxels of xyz (m,3) whose nearest seed in `tree` is seed ic -> (3,k)"""
    dists_, nns_ = tree.query(xyz, distance_upper_bound=rsearch)
    return xyz[nns_ == ic].T


class iter_fit_seed_points(IterFitBase):
    _personality = 4

    def __init__(self, im, centers, radius_fit=5, min_delta_center=1., max_delta_center=2.5,
                 n_max_iter=10, max_dist_th=0.1,
                 min_w=0.5, max_w=4, init_w=1.5, _stack=None):
        """Given seeds <centers> (3,N) in a 3d image <im>, iteratively fit 3d gaussians around
        the seeds (in order of brightness) and subtract the gaussian signal."""
        self._setup(im, centers, radius_fit, min_delta_center, max_delta_center, n_max_iter, max_dist_th,
                    min_w, max_w, init_w, 0.0, _stack)

    def firstfit(self):
        """First fit with the gaussian constrained close to the seed (delta = min_delta_center);
        every seed only sees the voxels of its window that are closest to it."""
        if len(self.centers) > 0:
            self._firstfit_device()
        else:
            # the reference reads self.im_subtr here, which firstfit never created (:639)
            raise AttributeError("'iter_fit_seed_points' object has no attribute 'im_subtr'")

    @property
    def gparms(self):
        """[im_, X, center] per seed, rebuilt on demand (host) from the same cKDTree rule."""
        from scipy.spatial import cKDTree
        tree = cKDTree(self.centers)
        out = []
        for ic, (zc, xc, yc) in enumerate(self.centers):
            z, x, y = in_dim(int(zc) + self.zb, int(xc) + self.xb, int(yc) + self.yb, self.sz, self.sx, self.sy)
            X = closest_faster(np.array([z, x, y], dtype=int).T, ic, tree, rsearch=self.radius_fit * 2)
            out.append([self.im[X[0], X[1], X[2]], X, [zc, xc, yc]])
        return out


def inv_sigma(sigma):
    """closed-form inverse of a symmetric 3x3 matrix (External/Fitting_v4.py:425-432)"""
    return np.linalg.inv(np.asarray(sigma, dtype=np.float64))


def fast_fit_big_image(im, centers_zxy, radius_fit=4, avoid_neigbors=True, recenter=False, verbose=True,
                       better_fit=False, troubleshoot=False):
    """Moment-based ("fast") fit of every seed (External/Fitting_v4.py:494-556).  better_fit=False:
    gfit_fast (:433-447) for all seeds in one kernel -> (N, 12) float64
    [h, z, x, y, background, cov_zz, cov_xx, cov_yy, cov_zx, cov_zy, cov_xy, nan]; better_fit=True: a
    GaussianFit (delta_center=2.5) on the same voxels, one device batch -> (N, 11)."""
    from .. import _lib
    from ..spot_tools.fitting import _device_image
    if troubleshoot:
        raise NotImplementedError("troubleshoot=True only adds matplotlib figures in the reference")
    centers_zxy = np.asarray(centers_zxy)
    if len(centers_zxy) == 0:
        return np.array([])
    arr = np.asarray(im)
    if not better_fit:
        dev = arr if arr.dtype in (np.uint16, np.float32, np.float64) else _device_image(arr)
        st = _lib.Stack(dev)
        out = st.moment_fit(centers_zxy[:, :3], radius_fit, avoid_neigbors, recenter, 0.1)
        st.close()
        return out
    # better_fit: windows exactly as the reference builds them (host bookkeeping), all fits in one batch
    from scipy.spatial import cKDTree
    from scipy.spatial.distance import cdist
    zb, xb, yb = window_offsets(radius_fit)
    X_c = np.array([zb, xb, yb]).T
    sz, sx, sy = arr.shape
    if avoid_neigbors:
        tree = cKDTree(centers_zxy)
        inters = tree.query_ball_tree(tree, radius_fit * 2)
    vals, coords, cens, empty = [], [], [], []
    for ic, (zc, xc, yc) in enumerate(centers_zxy):
        if avoid_neigbors:
            common = inters[ic]
            rel = centers_zxy[common] - [zc, xc, yc]
            keep = np.argmin(cdist(rel, X_c), 0) == common.index(ic)
            zb_, xb_, yb_ = X_c[keep].T
        else:
            zb_, xb_, yb_ = zb, xb, yb
        z, x, y = in_dim(int(zc) + zb_, int(xc) + xb_, int(yc) + yb_, sz, sx, sy)
        if recenter and len(z) > 0:
            k = np.argmax(arr[z, x, y])
            zc, xc, yc = z[k], x[k], y[k]
            z, x, y = in_dim(int(zc) + zb_, int(xc) + xb_, int(yc) + yb_, sz, sx, sy)
        v = arr[z, x, y]
        empty.append(len(v) == 0)
        if len(v) == 0:
            continue
        X = np.array([z, x, y])
        vals.append(np.asarray(v, dtype=np.float64))
        coords.append(X.astype(np.float32))
        cens.append(X[:, np.argmax(v)].astype(np.float64))
    cfg = _lib.make_fit_cfg(4, 5, 0.5, 4., 1.5)
    rows = iter(_lib.gaussfit_batch(cfg, 2.5, vals, coords, np.array(cens).reshape(-1, 3))[0]) if vals else iter(())
    ps = [np.array([np.nan] * 11) if e else next(rows) for e in empty]
    return np.array(ps)
