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
                 n_max_iter=10, max_dist_th=0.1, init_w=_sigma_zxy, weight_sigma=0, _stack=None):
        self._setup(im, centers, radius_fit, min_delta_center, max_delta_center, n_max_iter, max_dist_th,
                    0.5, 4., np.array(init_w[:3], dtype=float), weight_sigma, _stack)

    def firstfit(self):
        if len(self.centers) > 0:
            self._firstfit_device()
        else:
            raise ValueError(f"{len(self.centers)} points have been seeded, exit.")


def get_seed_points_base(im, gfilt_size_min=1, gfilt_size_max=3, filt_size=3, th_seed=0., max_num=None,
                         use_snr=False, hot_pix_th=0, return_h=False):
    """Legacy seeder of Fitting_v3 (External/Fitting_v3.py:261-306): local maxima of the raw image that are
    not local minima, height = blur(gfilt_size_min) - blur(gfilt_size_max) at those voxels, in the image's
    own dtype (on a uint16 image the subtraction wraps around, exactly as in the reference).  The rank
    filters and both Gaussian blurs run on the device; the bookkeeping below is the reference's numpy."""
    from .. import _lib
    from ..spot_tools.fitting import _device_image, _gauss_half_kernel
    im_plt = np.array(im)
    st = _lib.Stack(_device_image(im_plt))
    zxy, _, _ = st.seed_candidates(None, None, int(filt_size), 0, 0.0, -1e300)       # (max == im) & (min != im)
    z, x, y = (zxy[:, a].astype(np.int64) for a in range(3))
    st.seed_candidates(_gauss_half_kernel(gfilt_size_min), _gauss_half_kernel(gfilt_size_max), int(filt_size), 0, 0.0, 1e300)
    # the two blurs are read at the candidate voxels only (a gather on the device, a few kB back)
    flat = (z * im_plt.shape[1] + x) * im_plt.shape[2] + y
    g_sm = st.seed_volume_at(0, flat).astype(im_plt.dtype, copy=False)
    g_bg = st.seed_volume_at(1, flat).astype(im_plt.dtype, copy=False)
    st.close()
    with np.errstate(all='ignore'):
        h = g_sm - g_bg
        snr = 1. * g_sm / g_bg
    keep = snr > th_seed if use_snr else h > th_seed
    x, y, z = x[keep], y[keep], z[keep]
    h, snr = h[keep], snr[keep]
    if hot_pix_th > 0 and len(x) > 0:
        xy = y * np.max(x) + x
        xy_, cts_ = np.unique(xy, return_counts=True)
        bad_xy = xy_[cts_ > hot_pix_th]
        keep = ~np.isin(xy, bad_xy)
        x, y, z = x[keep], y[keep], z[keep]
        snr = snr[keep]
        h = h[keep]
    ind = np.argsort(snr)[::-1] if use_snr else np.argsort(h)[::-1]
    centers = np.array([z[ind], x[ind], y[ind]])
    if return_h:
        centers = np.array([z, x, y, h])
    if max_num is not None:
        centers = centers[:, :max_num]
    return centers
