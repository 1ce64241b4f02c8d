"""Drop-in for the fitting part of the reference's ``External/Fitting_v3.py``
(closest :40-47, GaussianFit :50-257, in_dim :308-310, iter_fit_seed_points :312-421): the legacy
"personality" -- per-axis initial widths, the to_center quirk of line 86, no overflow guards,
optional width prior ``weight_sigma`` and brute-force Voronoi membership (lowest index wins)."""
import numpy as np

from . import _sigma_zxy
from ._iterfit import GaussianFitBase, IterFitBase, in_dim, window_offsets  # noqa: F401


class GaussianFit(GaussianFitBase):
    _personality = 3

    def __init__(self, im, X, center=None, n_aprox=10, min_w=0.5, max_w=4., delta_center=3.,
                 init_w=_sigma_zxy, weight_sigma=0):
        self._setup(im, X, center, n_aprox, min_w, max_w, delta_center, np.array(init_w[:3], dtype=float), weight_sigma)


class iter_fit_seed_points(IterFitBase):
    _personality = 3

    def __init__(self, im, centers, radius_fit=5, min_delta_center=1., max_delta_center=2.5,
                 n_max_iter=10, max_dist_th=0.1, init_w=_sigma_zxy, weight_sigma=0, _stack=None, eval_fp32=False):
        self._setup(im, centers, radius_fit, min_delta_center, max_delta_center, n_max_iter, max_dist_th,
                    0.5, 4., np.array(init_w[:3], dtype=float), weight_sigma, _stack, eval_fp32)

    def firstfit(self):
        if len(self.centers) > 0:
            self._firstfit_device()
        else:
            raise ValueError(f"{len(self.centers)} points have been seeded, exit.")
