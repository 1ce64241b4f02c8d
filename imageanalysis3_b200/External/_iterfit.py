"""Shared host-side logic of the two iter_fit_seed_points / GaussianFit mirrors
(External/Fitting_v4.py:165-396,559-683 and External/Fitting_v3.py:50-257,312-421 of the
reference).  All fits run in libia3b200.so; this file only keeps the bookkeeping the reference
does in Python between fits (convergence test of repeatfit, list-valued attributes)."""
import numpy as np

from .. import _lib


def window_offsets(radius_fit):
    """np.indices([2r]*3) - r, kept where d^2 <= r^2 (C order): offsets -r..r-1."""
    zb, xb, yb = np.reshape(np.indices([radius_fit * 2] * 3) - radius_fit, [3, -1])
    keep = zb * zb + xb * xb + yb * yb <= radius_fit ** 2
    return zb[keep], xb[keep], yb[keep]


def in_dim(x, y, z, xmax, ymax, zmax):
    keep = ((x >= 0) & (x < xmax) & (y >= 0) & (y < ymax) & (z >= 0) & (z < zmax)) > 0
    return x[keep], y[keep], z[keep]


class GaussianFitBase:
    """One constrained 10-parameter 3D Gaussian fit (device batch of one problem)."""
    _personality = 4

    def _setup(self, im, X, center, n_aprox, min_w, max_w, delta_center, init_w, weight_sigma):
        if n_aprox != 10:
            raise NotImplementedError("the device fit uses n_aprox=10 (the only value used by the reference)")
        self._min_w, self._max_w = float(min_w), float(max_w)
        self.min_w = min_w * min_w
        self.max_w = max_w * max_w
        self.delta_center = delta_center
        self.weight_sigma = weight_sigma
        self._values = np.asarray(im)
        self.im = np.array(im, dtype=np.float32)
        self.x, self.y, self.z = np.array(X, dtype=np.float32)
        if center is None:
            order = np.argsort(im)
            center = np.median(np.asarray(X)[:, order][:, -n_aprox:], -1)
        self.center_est = center
        self._init_w = init_w
        self.success = False
        self.p = None
        self.p_ = None

    def _cfg(self):
        return _lib.make_fit_cfg(self._personality, 5, self._min_w, self._max_w, self._init_w,
                                 weight_sigma=self.weight_sigma, maxfev=0)

    def fit(self, eps_frac=10E-3, eps_dist=10E-3, eps_angle=10E-3):
        """Levenberg-Marquardt (MINPACK lmder semantics) on the device.
        [height,x,y,z,background,width_1,width_2,width_3,sin_theta,sin_phi,error] = self.p"""
        self.eps_frac, self.eps_dist, self.eps_angle = eps_frac, eps_dist, eps_angle
        if len(self.im) < 10:
            self.success = False
            return
        coords = np.array([self.x, self.y, self.z], dtype=np.float32)
        ps, praw, succ, nfev, info, _ = _lib.gaussfit_batch(
            self._cfg(), self.delta_center, [np.asarray(self._values, dtype=np.float64)], [coords],
            np.asarray(self.center_est, dtype=np.float64)[None, :])
        self.p_ = praw[0]
        self.p = ps[0]
        self.nfev, self.ier = int(nfev[0]), int(info[0])
        self.center = self.p[1:4]
        self.success = True

    def get_im(self):
        """Gaussian part of the fitted model (no background) on the current self.x, self.y, self.z."""
        coords = np.array([self.x, self.y, self.z], dtype=np.float64).astype(np.float32)
        self.f0 = _lib.gauss_eval(self._cfg(), self.delta_center, self.p_, self.center_est, coords)
        return self.f0


class IterFitBase:
    _personality = 4

    # -- construction ----------------------------------------------------------------------
    def _setup(self, im, centers, radius_fit, min_delta_center, max_delta_center, n_max_iter, max_dist_th,
               min_w, max_w, init_w, weight_sigma, _stack):
        self.im = im
        self.radius_fit = radius_fit
        self.n_max_iter = n_max_iter
        self.max_dist_th = max_dist_th
        self.min_delta_center = min_delta_center
        self.max_delta_center = max_delta_center
        self.centers = centers.T
        self.z, self.x, self.y = centers
        self.zb, self.xb, self.yb = window_offsets(radius_fit)
        self.zxyb = np.array([self.zb, self.xb, self.yb]).T
        self.sz, self.sx, self.sy = im.shape
        self.min_w = min_w
        self.max_w = max_w
        self.init_w = init_w
        self.weight_sigma = weight_sigma
        self._stack = _stack
        self._h = None

    def _make_handle(self):
        if self._stack is None:
            self._stack = _lib.Stack(np.asarray(self.im))
        cfg = _lib.make_fit_cfg(self._personality, self.radius_fit, self.min_w, self.max_w, self.init_w,
                                weight_sigma=self.weight_sigma, maxfev=0)
        self._h = _lib.FitHandle(self._stack, np.asarray(self.centers, dtype=np.float64), cfg)
        return self._h

    # -- firstfit ----------------------------------------------------------------------------
    def _prepare_device(self):
        """handle + Voronoi membership of the window voxels; ties go through scipy's own tree"""
        h = self._make_handle()
        n_ties = h.first_prepare()
        self.n_tie_voxels = n_ties
        if n_ties and self._personality == 4:
            # cKDTree's choice among equidistant seeds is an implementation detail of scipy:
            # ask the very same tree about the (few) tied voxels   (Fitting_v4.py:422-424)
            from scipy.spatial import cKDTree
            spot, zxy = h.first_ties(n_ties)
            tree = cKDTree(self.centers)
            _, nn = tree.query(zxy, distance_upper_bound=self.radius_fit * 2)
            h.first_resolve(nn == spot)
        return h

    def _take_first(self):
        h = self._h
        self._ps = h.ps.copy()
        self._succ = h.success.astype(bool)
        self._has_fit = self._succ.copy()
        self.nfev = h.nfev.copy()
        self.info = h.info.copy()

    def _firstfit_device(self):
        h = self._prepare_device()
        h.run(1, self.min_delta_center, self.max_delta_center, self.max_dist_th ** 2, self.n_max_iter)
        self._take_first()

    def _fit_all(self):
        """firstfit() followed by repeatfit() as ONE device run (no host round trip in between, the first
        repeat visit of isolated seeds overlaps their firstfit).  Leaves the object as the two calls do."""
        if len(self.centers) == 0:
            return self.firstfit()             # raises exactly as the reference does on an empty seed list
        h = self._prepare_device()
        h.run(3, self.min_delta_center, self.max_delta_center, self.max_dist_th ** 2, self.n_max_iter)
        self._take_repeat()

    # -- list-valued attributes the callers read --------------------------------------------
    @property
    def ps(self):
        return [row if ok else [np.nan] * 11 for row, ok in zip(self._ps, self._ever_ok())]

    def _ever_ok(self):
        # rows that hold a fit (anything else is the NaN placeholder of a failed firstfit)
        return self._has_fit

    @property
    def success(self):
        return [bool(v) for v in self._succ]

    @property
    def centers_fit(self):
        return [row[1:4] if ok else [np.nan] * 3 for row, ok in zip(self._ps, self._ever_ok())]

    @property
    def ims_rec(self):
        return [self._h.rec(i)[0] if ok else np.nan for i, ok in enumerate(self._ever_ok())]

    @property
    def im_subtr(self):
        if self._h is None:
            raise AttributeError(f"'{type(self).__name__}' object has no attribute 'im_subtr'")
        return self._h.volume(0)

    @property
    def im_add(self):
        if self._h is None:
            raise AttributeError(f"'{type(self).__name__}' object has no attribute 'im_add'")
        return self._h.volume(1)

    def _ps_array(self):
        """np.array(self.ps) without building 10^4 python objects: float32 (n, 11) if every row holds a
        fit, float64 with NaN rows otherwise (the reference's dtype trap)."""
        ok = self._ever_ok()
        if ok.all():
            return self._ps.copy()
        out = self._ps.astype(np.float64)
        out[~ok] = np.nan
        return out

    def _centers_fit_array(self):
        """np.array(self.centers_fit): float32 if every row is a fit, float64 otherwise."""
        ok = self._ever_ok()
        if ok.all():
            return np.ascontiguousarray(self._ps[:, 1:4])
        out = self._ps[:, 1:4].astype(np.float64)
        out[~ok] = np.nan
        return out

    # -- repeatfit ---------------------------------------------------------------------------
    def _take_repeat(self):
        """attributes of the reference after repeatfit() from the per-seed results of the device run"""
        h = self._h
        n = len(self.centers)
        self._ps = h.ps.copy()
        self._succ = h.success.astype(bool)
        self._has_fit = self._succ.copy()        # every seed is visited in sweep 1: it holds a fit iff its last visit succeeded
        self.nfev = h.nfev.copy()
        self.info = h.info.copy()
        self.n_iter = int(h.n_visits.max()) if n else 0
        self.converged = h.converged.astype(bool)
        # state at the start of the last sweep: seeds visited in it held success_old / centers_old
        last = h.n_visits == self.n_iter
        # the reference recomputes dists for every seed after each sweep: a seed that was not visited in the
        # last one has old == new there, i.e. 0
        self.dists = np.where(last, h.dists, 0.0)
        self.success_old = np.where(last, h.success_old.astype(bool), self._succ)
        had_fit = np.where(last & (self.n_iter == 1), h.success_old.astype(bool), self._has_fit)
        old = np.where(last[:, None], h.centers_old, self._ps[:, 1:4])
        if had_fit.all():
            self.centers_fit_old = np.ascontiguousarray(old, dtype=np.float32)
        else:
            old = old.astype(np.float64)
            old[~had_fit] = np.nan
            self.centers_fit_old = old

    def repeatfit(self):
        n = len(self.centers)
        if n == 0:
            self.n_iter = 0
            self.converged = np.zeros(0, dtype=bool)
            self.dists = np.zeros(0) + np.inf
            return
        self._h.run(2, self.min_delta_center, self.max_delta_center, self.max_dist_th ** 2, self.n_max_iter)
        self._take_repeat()

    def _repeatfit_host_loop(self):
        """The reference's loop run on the host, one device sweep per iteration (Fitting_v4.py:641-683):
        kept as a cross-check of the device-resident rule above (tests)."""
        n = len(self.centers)
        self.n_iter = 0
        self.converged = np.zeros(n, dtype=bool)
        self.dists = np.zeros(n) + np.inf
        converged = np.all(self.converged)
        while not converged:
            self.success_old, self.centers_fit_old = self._succ.copy(), self._centers_fit_array().copy()
            self._h.repeat_sweep(self.max_delta_center, ~self.converged)
            self._ps = self._h.ps.copy()
            self._succ = self._h.success.astype(bool)
            self._has_fit |= self._succ
            self.nfev = self._h.nfev.copy()
            self.info = self._h.info.copy()
            keep = (self._succ & self.success_old) > 0
            self.dists[~keep] = 0
            new = self._centers_fit_array()
            old = self.centers_fit_old
            if old.dtype != new.dtype:
                old, new = old.astype(np.float64), new.astype(np.float64)
            self.dists[keep] = np.sum((old[keep] - new[keep]) ** 2, axis=-1)
            self.converged = self.dists < self.max_dist_th ** 2
            converged = np.all(self.converged)
            self.n_iter += 1
            converged = converged or (self.n_iter > self.n_max_iter)
