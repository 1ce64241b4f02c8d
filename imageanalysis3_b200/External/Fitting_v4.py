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
                 min_w=0.5, max_w=4, init_w=1.5, _stack=None, eval_fp32=False):
        """Given seeds <centers> (3,N) in a 3d image <im>, iteratively fit 3d gaussians around
        the seeds (in order of brightness) and subtract the gaussian signal."""
        self._setup(im, centers, radius_fit, min_delta_center, max_delta_center, n_max_iter, max_dist_th,
                    min_w, max_w, init_w, 0.0, _stack, eval_fp32)

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
