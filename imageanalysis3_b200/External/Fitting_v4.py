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


# ---- the remaining public names of the reference module -----------------------------
def _cov_matrices(t, p, w_1, w_2, w_3):
    """inverse covariance of the rotated Gaussian with sin(theta) = t, sin(phi) = p and widths w_i"""
    t2, p2 = t * t, p * p
    ct2, cp2 = 1 - t2, 1 - p2
    ct, cp = np.sqrt(ct2), np.sqrt(cp2)
    s1, s2, s3 = 1. / (w_1 * w_1), 1. / (w_2 * w_2), 1. / (w_3 * w_3)
    a = cp2 * ct2 * s1 + t2 * s2 + p2 * ct2 * s3
    b = cp2 * t2 * s1 + ct2 * s2 + p2 * t2 * s3
    c = p2 * s1 + cp2 * s3
    d = ct * t * (cp2 * s1 - s2 + p2 * s3)
    e = p * cp * ct * (s3 - s1)
    f = p * cp * t * (s3 - s1)
    return a, b, c, d, e, f


def _adjugate_over_det(a, b, c, d, e, f):
    det = a * b * c - c * d ** 2 - b * e ** 2 + 2 * d * e * f - a * f ** 2
    adj = np.array([[b * c - f ** 2, -(c * d) + e * f, -(b * e) + d * f],
                    [-(c * d) + e * f, a * c - e ** 2, d * e - a * f],
                    [-(b * e) + d * f, d * e - a * f, a * b - d ** 2]])
    return adj, det


def to_sigmas(t, p, w_1, w_2, w_3):
    """covariance matrix and its inverse from t = sin(theta), p = sin(phi) and the three widths
    (External/Fitting_v4.py:685-704) -> (sigma, sigma_inv)"""
    a, b, c, d, e, f = _cov_matrices(t, p, w_1, w_2, w_3)
    adj, det = _adjugate_over_det(a, b, c, d, e, f)
    return adj / det, np.array([[a, d, e], [d, b, f], [e, f, c]])


def to_sigmas_abc(a, b, c, d, e, f):
    """(sigma, sigma_inv, det) of the symmetric matrix [[a, d, e], [d, b, f], [e, f, c]] (:705-710)"""
    adj, det = _adjugate_over_det(a, b, c, d, e, f)
    adj = np.array([[b * c - f * f, -(c * d) + e * f, -(b * e) + d * f], [-(c * d) + e * f, a * c - e * e, d * e - a * f],
                    [-(b * e) + d * f, d * e - a * f, a * b - d * d]])
    return np.array([[a, d, e], [d, b, f], [e, f, c]]), adj / det, det


def gfit_fast(im_, X_, bk_f=0.1, reconstruct=False, plt_val=False, compare_with_fitting=False):
    """Moment estimate of ONE spot from its voxel values im_ (m,) and coordinates X_ (3, m) (:433-491, without the
    matplotlib branches) -> [h, z, x, y, background, cov_zz, cov_xx, cov_yy, cov_zx, cov_zy, cov_xy, eps].
    A few hundred numbers: plain numpy, like the reference; whole seed lists go through fast_fit_big_image (device)."""
    if plt_val:
        raise NotImplementedError("plt_val=True only draws matplotlib figures in the reference")
    out = np.array([np.nan] * 12)
    if len(im_) == 0:
        return out
    X_ = np.asarray(X_)
    bk = np.sort(im_)[int(len(im_) * bk_f)]
    wts = im_ - bk
    wts[wts < 0] = 0
    h = np.max(wts)
    wts = wts / np.sum(wts)
    mu = np.sum(X_ * wts, -1)
    dev = X_.T - mu
    cov = np.sum(np.array([[dev[:, i] * dev[:, j] for i in range(3)] for j in range(3)]) * wts, -1)
    [[a, d, e], [d, b, f], [e, f, c]] = cov
    eps = np.nan
    if reconstruct:
        icov = inv_sigma(cov)
        model = h * np.exp(-np.sum(np.dot(dev, icov) * dev, -1) * 0.5) + bk
        eps = np.mean(np.abs(im_ - model))
    return np.array([h, mu[0], mu[1], mu[2], bk, a, b, c, d, e, f, eps])


def gker(gaus=[3, 3, 3], exp=8):
    """normalised float32 outer product of scipy.signal.windows.gaussian(int(g * exp), g) (:11-16)"""
    from scipy.signal.windows import gaussian
    sz = [int(g * exp) for g in gaus]
    k = np.outer(np.outer(gaussian(sz[0], gaus[0]), gaussian(sz[1], gaus[1])), gaussian(sz[2], gaus[2])).reshape(sz)
    return (k / np.sum(k)).astype(np.float32)


def fft_gaussian_fast(arr, gaus=[3, 3, 3], exp=8):
    """Gaussian blur of the reference's FFT path (:66-70: reflect() padding, 'valid' convolution with gker) evaluated
    on the device as three direct FP64 passes.  The reference multiplies single-precision FFTs, so the two agree to
    ~1e-6 relative, not bit for bit."""
    from .. import _lib
    st = _lib.Stack(np.asarray(arr))
    out = st.fft_gaussian(gaus, exp)
    st.close()
    return out


def _sorted_centers(shape, idx, h, max_num):
    z, x, y = np.unravel_index(idx, shape)
    order = np.argsort(h)[::-1]
    cen = np.array([z[order], x[order], y[order], h[order]])
    return cen[:, :max_num] if max_num is not None else cen


def get_seed_points_base(im_sm, gfilt_size=5, filt_size=3, th_seed=3., max_num=None):
    """Seeds of the log-ratio image log(im) - log(fft_gaussian_fast(im, [gfilt_size] * 3)): local maxima within
    filt_size that stand th_seed standard deviations out (:72-92) -> ((4, N) [z, x, y, h] brightest first, std)."""
    from .. import _lib
    st = _lib.Stack(np.asarray(im_sm))
    idx, h, std = st.seed_logratio(gfilt_size, filt_size, th_seed)
    st.close()
    return _sorted_centers(np.shape(im_sm), idx, h, max_num), std


def get_seed_points_base_v2(im_sm, gfilt_size=5, filt_size=3, th_seed=3., max_num=None):
    """Seeds of im - cv2.blur(im, (gfilt_size, gfilt_size)) per slice: voxels above th_seed standard deviations that
    are >= their filt_size^3 neighbours, neighbours taken modulo the shape (:95-126) -> ((4, N) float32 [z, x, y, h]
    brightest first, std).  The blur is bit-identical to cv2's; std is reduced in FP64 on the device (numpy: float32
    pairwise), which moves the cutoff by ~1e-7 relative."""
    from .. import _lib
    if gfilt_size == 0:
        raise NotImplementedError("gfilt_size=0 (no normalisation) is not available on the device")
    st = _lib.Stack(np.asarray(im_sm))
    idx, h, std = st.seed_v2(gfilt_size, filt_size, th_seed)
    st.close()
    z, x, y = np.unravel_index(idx, np.shape(im_sm))
    order = np.argsort(h)[::-1]
    cen = np.array([z[order], x[order], y[order], h[order]])       # int64 and float32 rows -> float64, like the reference
    return (cen[:, :max_num] if max_num is not None else cen), std
