"""Drop-in for the reference's ``spot_tools/fitting.py`` (get_seeds :20-154, remove_edge_points
:156-165, fit_fov_image :169-262, get_centers :268-334, select_sparse_centers :338-363).

Same names, arguments, defaults, return arrays and exceptions.  The heavy lifting -- both
Gaussian blurs, the rank filters, the candidate mask, the ordered compaction and all
Levenberg-Marquardt fits -- runs in libia3b200.so on the B200; this module keeps only the
order- and dtype-sensitive numpy bookkeeping that operates on the (small) candidate list:
the dynamic-threshold descent, the hot-pixel filter, the unstable argsort and the top-N cut
(SURVEY.md App. A, steps 6-10).
"""
import os
import threading
import time

import numpy as np

from .. import _lib
from ..External import Fitting_v4


def _gauss_half_kernel(sigma, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d (order 0) -- centre..edge half, float64."""
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    x = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[lw:])


def _effective_threshold(th, h_dtype):
    """The number numpy really compares a ``h_dtype`` array against in ``h >= th``
    (python scalars are weak under NEP 50 and are cast to the array dtype)."""
    if isinstance(th, np.generic):
        return float(np.asarray(th).astype(np.result_type(h_dtype, th.dtype)))
    return float(np.asarray(th, dtype=h_dtype))


def _score_from_counts(counts, per):
    """scipy.stats.scoreatpercentile(a, per) for an integer array given as the counts of its values: the two order
    statistics around per / 100 * (n - 1) and scipy's own interpolation arithmetic."""
    if not (0 <= per <= 100):
        raise ValueError("percentile must be in the range [0, 100]")
    n = int(counts.sum())
    idx = per / 100. * (n - 1)
    i = int(idx)
    cum = np.cumsum(counts.astype(np.int64))
    a_i = int(np.searchsorted(cum, i, side='right'))
    if i == idx:
        return np.float64(a_i) / 1.0
    a_j = int(np.searchsorted(cum, i + 1, side='right'))
    w = np.array([(i + 1 - idx), (idx - i)], float)
    return np.add.reduce(np.array([a_i, a_j], dtype=np.uint16) * w) / w.sum()


def _percentile_span(im, stack, hi, lo):
    """scoreatpercentile(im, hi) - scoreatpercentile(im, lo): from the device histogram when the stack is uint16 and
    resident, else with scipy on the host"""
    if stack is not None and stack.dtype == np.uint16 and im.dtype == np.uint16 and tuple(stack.shape) == tuple(im.shape):
        counts = stack.histogram()
        return _score_from_counts(counts, hi) - _score_from_counts(counts, lo)
    from scipy.stats import scoreatpercentile
    return scoreatpercentile(im, hi) - scoreatpercentile(im, lo)


def _device_image(im):
    """dtype handling of the seed stage: uint16 / float32 / float64 are native.  Other integer types whose values fit
    uint16 are narrowed: scipy filters in double and stores in the input's integer type by truncation, so the
    filtered integers (and hence every seed) are the same in either type."""
    if im.dtype in (np.uint16, np.float32, np.float64):
        return im
    if im.dtype.kind in "uib":
        if im.size == 0 or (im.min() >= 0 and im.max() <= 65535):
            return im.astype(np.uint16)
        raise NotImplementedError(f"integer stacks on the GPU must hold values in [0, 65535], got {im.dtype} "
                                  f"with range [{im.min()}, {im.max()}]")
    raise NotImplementedError(f"seed stage on the GPU supports integer, float32 and float64 stacks, got {im.dtype}")


def get_seeds(im, max_num_seeds=None, th_seed=150,
              th_seed_per=95, use_percentile=False,
              sel_center=None, seed_radius=30,
              gfilt_size=0.75, background_gfilt_size=7.5,
              filt_size=3, min_edge_distance=2,
              use_dynamic_th=True, dynamic_niters=10, min_dynamic_seeds=1,
              remove_hot_pixel=True, hot_pixel_th=3,
              return_h=False, verbose=False, _stack=None):
    """Seeding pixels of a 3D image: local maxima of the sigma=gfilt_size blur that are not
    local minima of the background blur and stand at least th_seed above it."""
    if not isinstance(im, np.ndarray):
        raise TypeError(f"image given should be a numpy.ndarray, but {type(im)} is given.")
    if th_seed_per >= 100 or th_seed_per <= 50:
        use_percentile = False
        print(f"th_seed_per should be a percentile > 50, invalid value given ({th_seed_per}), so not use percentile here.")
    ndim = len(np.shape(im))
    if sel_center is not None:
        if len(sel_center) != ndim:
            raise IndexError(f"num of dimensions should match for selected center and image given.")
        ctr = np.array(sel_center, dtype=int)
        lo = np.max([np.zeros(ndim), ctr - seed_radius], axis=0)
        hi = np.min([np.array(im.shape), ctr + seed_radius], axis=0)
        lims = np.array(np.transpose(np.stack([lo, hi])), dtype=int)
        sub = im[tuple(slice(a, b) for a, b in lims)]
        local_edges = lo
        stack = None
    else:
        local_edges = np.zeros(ndim)
        sub = im
        stack = _stack

    if use_percentile:
        if _stack is None and sel_center is None and im.ndim == 3 and im.dtype == np.uint16:
            _stack = stack = _lib.Stack(im)              # the percentile is taken over the FULL image (fitting.py:76)
        th0 = _percentile_span(im, _stack, th_seed_per, (100 - th_seed_per) / 2)
    else:
        th0 = th_seed
    if verbose:
        t0 = time.time()
        if not use_dynamic_th:
            print(f"-- start seeding image with threshold: {th0:.2f}", end='; ')
        else:
            print(f"-- start seeding image, th={th0:.2f}", end='')
    niters = int(dynamic_niters) if use_dynamic_th else 1
    ths = [th0 * (1 - it / niters) for it in range(niters)]

    if sub.ndim not in (2, 3):
        raise NotImplementedError("the GPU seed stage needs a 2D (X, Y) image or a 3D (Z, X, Y) stack")
    two_d = sub.ndim == 2
    if stack is None:
        dev = _device_image(sub)
        stack = _lib.Stack(dev[np.newaxis] if two_d else dev)      # a 2D image is held as a one-plane stack
    h_dtype = np.float32
    floor = min(_effective_threshold(t, h_dtype) for t in ths)
    zxy, hs_all, timing = stack.seed_candidates(
        _gauss_half_kernel(gfilt_size) if gfilt_size else None,
        _gauss_half_kernel(background_gfilt_size) if background_gfilt_size else None,
        int(filt_size), 0, float(min_edge_distance) if min_edge_distance > 0 else 0.0, floor, two_d=two_d)
    if two_d:
        zxy = zxy[:, 1:]

    # dynamic threshold descent on the candidate list (fitting.py:113-125)
    for th in ths:
        sel = hs_all >= th
        if np.count_nonzero(sel) >= min_dynamic_seeds:
            break
    if verbose and use_dynamic_th:
        print(f"->{th:.2f}", end=', ')
    coords = tuple(zxy[sel, a].astype(np.int64) for a in range(ndim))
    hs = hs_all[sel]
    # hot pixels: (x, y) columns holding >= hot_pixel_th seeds (fitting.py:131-138)
    if remove_hot_pixel:
        if two_d:
            raise IndexError("tuple index out of range")     # the reference indexes _coords[2] here: 3D stacks only
        key = coords[1] * (int(sub.shape[2]) + 1) + coords[2]
        _, inv, cts = np.unique(key, return_inverse=True, return_counts=True)
        keep = cts[inv] < hot_pixel_th if len(key) else np.zeros(0, dtype=bool)
        coords = tuple(c[keep] for c in coords)
        hs = hs[keep]
    final = np.array(coords) + local_edges[:, np.newaxis]
    if return_h:
        final = np.concatenate([final, hs[np.newaxis, :]])
    final = np.transpose(final)[np.flipud(np.argsort(hs))]
    if verbose:
        print(f"found {len(final)} seeds in {time.time()-t0:.2f}s")
    if max_num_seeds is not None and max_num_seeds > 0 and max_num_seeds <= len(final):
        final = final[:int(max_num_seeds)]
        if verbose:
            print(f"--- {max_num_seeds} seeds are kept.")
    return final


def remove_edge_points(im, T_seeds, distance=2):
    """keep flags for seeds with distance <= coord <= size - distance on every axis (inclusive)"""
    size = np.array(np.shape(im))
    pts = np.array(T_seeds)[:len(size), :].transpose()
    if len(pts) == 0:
        return np.array([], dtype=bool)
    return np.array(((pts >= distance) & (pts <= size - distance)).all(axis=1), dtype=bool)


def _find_image_background(im, dtype=np.uint16, bin_size=10, max_iter=10):
    """io_tools/load.py:642-686 -- mode of the intensity histogram (host)."""
    import scipy.signal
    if dtype is None:
        dtype = im.dtype
    cts, bins = np.histogram(im, bins=np.arange(np.iinfo(dtype).min, np.iinfo(dtype).max, bin_size))
    peaks, height, it = [], np.size(im) / 50, 0
    while len(peaks) == 0:
        height = height / 2
        peaks, params = scipy.signal.find_peaks(cts, height=height)
        it += 1
        if it > max_iter:
            break
    if it > max_iter:
        return np.nanmedian(im)
    sel = peaks[np.argmax(params['peak_heights'])]
    return (bins[sel] + bins[sel + 1]) / 2


def _neighbor_boxes(coords, crop_size, shape):
    """generate_neighboring_crop (io_tools/crop.py:59-88, sub_pixel_precision=False) for all spots at once:
    (M, 3) centres -> (M, 6) int32 [z0, z1, x0, x1, y0, y1) (np.round = half to even, on float64)"""
    c = np.asarray(coords)[:, :len(shape)].astype(np.float64)
    size = np.ones(len(shape), dtype=np.int32) * crop_size
    lo = np.maximum(np.round(c - size), 0.0)
    hi = np.minimum(np.round(c + size + 1), np.array(shape, dtype=np.int32))
    out = np.empty((len(c), 2 * len(shape)), dtype=np.int32)
    out[:, 0::2] = lo.astype(np.int32)
    out[:, 1::2] = hi.astype(np.int32)
    return out


def _image_backgrounds(im, stack, boxes, dtype='uint16', bin_size=10, make_plot=False, max_iter=10, _resident=False):
    """find_image_background (io_tools/load.py:642-686) for every box of ``im``.  uint16 stacks that are
    resident on the device are histogrammed there (one CTA per box); anything else (float images, other
    ``dtype`` arguments) keeps the reference's own numpy/scipy expressions."""
    boxes = np.asarray(boxes, dtype=np.int32).reshape(-1, 6)
    if dtype is None:
        dtype = im.dtype
    on_device = (stack is not None and im.dtype == np.uint16 and np.dtype(dtype).kind in "ui"
                 and float(bin_size) == int(bin_size) and int(bin_size) >= 1)
    if on_device:
        info = np.iinfo(dtype)
        first, last = int(info.min), int(info.max)
        if first >= 0 and last <= 65535:
            empty = (boxes[:, 1] <= boxes[:, 0]) | (boxes[:, 3] <= boxes[:, 2]) | (boxes[:, 5] <= boxes[:, 4])
            safe = boxes.copy()
            safe[empty] = 0
            out = stack.box_background(safe, first, last, int(bin_size), int(max_iter))
            out[empty] = np.nan
            return out
    if _resident:
        raise NotImplementedError("these background_args need the image on the host: pass the numpy image, not a resident stack")
    return np.array([_find_image_background(im[b[0]:b[1], b[2]:b[3], b[4]:b[5]], dtype=dtype, bin_size=bin_size, max_iter=max_iter)
                     for b in boxes])


def _neighbor_slices(coord, crop_size, shape):
    """io_tools/crop.py:59-88 (sub_pixel_precision=False) -> tuple of slices"""
    coord = np.array(coord)[:len(shape)]
    size = np.ones(len(shape), dtype=np.int32) * crop_size
    lo = np.max([np.round(coord - size), np.zeros(len(shape))], axis=0)
    hi = np.min([np.round(coord + size + 1), np.array(shape, dtype=np.int32)], axis=0)
    return tuple(slice(int(a), int(b)) for a, b in zip(lo, hi))


# Stacks that hold seed-stage work volumes at the same time (callers that keep many stacks in flight,
# sharding.map_stacks).  The seed kernels fill the GPU with a few stacks; each one needs three more
# copies of the stack in HBM, and letting all of them seed at once only inflates the allocation pool
# (cudaMalloc while other stacks' kernels run stalls every stream).
_SEED_GATE = threading.BoundedSemaphore(max(1, int(os.environ.get("IA3_SEED_INFLIGHT", "6"))))
# Stacks between "upload started" and "fit done" (they hold their image in HBM and a place in the upload
# queue): callers may keep more host threads than this in flight, the rest wait here.
_ADMIT_GATE = threading.BoundedSemaphore(max(1, int(os.environ.get("IA3_ADMIT_INFLIGHT", "64"))))


def fit_fov_image(im, channel, *args, **kwargs):
    """Seeding + fitting of a whole field-of-view stack -> (M, 11) spots
    [height, z, x, y, background, sigma_z, sigma_x, sigma_y, sin_t, sin_p, eps]
    (signature and defaults of the reference: see _fit_fov_image)."""
    state = {"held": True}

    def front_done():
        if state["held"]:
            state["held"] = False
            _ADMIT_GATE.release()
    _ADMIT_GATE.acquire()
    try:
        return _fit_fov_image(im, channel, *args, _front_done=front_done, **kwargs)
    finally:
        front_done()


def _public_signature():
    import inspect
    sig = inspect.signature(_fit_fov_image)
    return sig.replace(parameters=[p for n, p in sig.parameters.items() if n != "_front_done"])


def _fit_fov_image(im, channel, seeds=None,
                  seed_mask=None,
                  max_num_seeds=500,
                  th_seed=300, th_seed_per=95, use_percentile=False,
                  use_dynamic_th=True,
                  dynamic_niters=10, min_dynamic_seeds=1,
                  remove_hot_pixel=True, seeding_kwargs={},
                  fit_radius=5,
                  normalize_background=False, normalize_local=False,
                  background_args={},
                  fitting_args={},
                  remove_boundary_points=True, verbose=True, _stack=None, _front_done=None):
    """spot_tools/fitting.py:169-262.  `_stack` (not in the reference): a `_lib.Stack` that already holds
    `im` in HBM -- it keeps its image (only the seed stage's scratch volumes are handed back); `_front_done`:
    called once the image is no longer needed on the device."""
    th_seed = float(th_seed)
    if verbose:
        print(f"-- start fitting spots in channel:{channel}, ", end='')
        t0 = time.time()
    resident = isinstance(im, _lib.Stack)
    if resident:
        # a stack that is already in HBM (e.g. from io_tools.load.correct_fov_image(..., return_stacks=True)): `im`
        # below only carries shape and dtype, its values are never read
        if im.dtype != np.uint16 or len(im.shape) != 3:
            raise NotImplementedError("resident input: uint16 (Z, X, Y) stacks")
        if 'sel_center' in seeding_kwargs:
            raise NotImplementedError("sel_center crops the image on the host: pass the numpy image, not a resident stack")
        _stack = im
        im = np.broadcast_to(np.zeros((), dtype=np.uint16), _stack.shape)
    stack = _stack
    own_stack = _stack is None
    if stack is None and isinstance(im, np.ndarray) and im.ndim == 3 and im.dtype.kind in "uibf" and im.dtype.itemsize >= (4 if im.dtype.kind == "f" else 1):
        stack = _lib.Stack(_device_image(im))     # one upload shared by the seed and the fit stage
    if seeds is None:
        with _SEED_GATE:
            _seeds = get_seeds(im, max_num_seeds=max_num_seeds,
                               th_seed=th_seed, th_seed_per=th_seed_per,
                               use_percentile=use_percentile,
                               use_dynamic_th=use_dynamic_th,
                               dynamic_niters=dynamic_niters,
                               min_dynamic_seeds=min_dynamic_seeds,
                               remove_hot_pixel=remove_hot_pixel,
                               return_h=False, verbose=False,
                               _stack=stack if 'sel_center' not in seeding_kwargs else None,
                               **seeding_kwargs)
            if stack is not None:
                stack.trim(1)          # the seed stage's scratch volumes go back to the other stacks in flight (never the image)
        if verbose:
            print(f"{len(_seeds)} seeded with th={th_seed}, ", end='')
    else:
        _seeds = np.array(seeds)[:, :len(np.shape(im))]
        if verbose:
            print(f"{len(_seeds)} given, ", end='')
    if len(_seeds) == 0:
        return np.array([])
    if seed_mask is not None:
        picked = [s for s in _seeds
                  if seed_mask[tuple(np.round(s[:len(np.shape(im))]).astype(np.int32))] > 0]
        _seeds = np.array(picked)
        if verbose:
            print(f"{len(_seeds)} selected by mask, ", end='')

    fitter = Fitting_v4.iter_fit_seed_points(im, _seeds.T, radius_fit=fit_radius, _stack=stack, **fitting_args)
    fitter._fit_all()          # firstfit() + repeatfit() as one device run
    _need_image = normalize_local or normalize_background
    if stack is not None and own_stack and not _need_image:
        stack.trim(2)          # only a stack created here: a caller's stack keeps its image for the caller's next call
    if _front_done is not None:
        _front_done()
    _spots = fitter._ps_array()                 # == np.array(fitter.ps)
    _spots = _spots[np.sum(np.isnan(_spots), axis=1) == 0]
    if remove_boundary_points:
        inside = (_spots[:, 1:4] > np.zeros(3)).all(1) * (_spots[:, 1:4] < np.array(np.shape(im))).all(1)
        _spots = _spots[np.where(inside)[0]]
    if normalize_background and not normalize_local:
        back = _image_backgrounds(im, stack, np.array([[0, im.shape[0], 0, im.shape[1], 0, im.shape[2]]]), _resident=resident, **background_args)[0]
        if verbose:
            print(f"normalize total background:{back:.2f}, ", end='')
        _spots[:, 0] = _spots[:, 0] / back
    elif normalize_local:
        backs = _image_backgrounds(im, stack, _neighbor_boxes(_spots[:, 1:4], fit_radius * 2, np.shape(im)), _resident=resident, **background_args)
        if verbose:
            print(f"normalize local background for each spot, ", end='')
        _spots[:, 0] = _spots[:, 0] / np.array(backs)
    if verbose:
        print(f"{len(_spots)} fitted in {time.time()-t0:.3f}s.")
    return _spots


def get_centers(im, seeds=None, th_seed=150,
                th_seed_per=98, use_percentile=False,
                sel_center=None, seed_radius=40,
                max_num_seeds=None, use_dynamic_th=True,
                min_num_seeds=1,
                remove_hot_pixel=True, hot_pixel_th=3,
                seed_kwargs={},
                fit_radius=5,
                remove_close_pts=True, close_threshold=0.1,
                verbose=False):
    """Fitted centres (bead / spot) of one image -> (K, 3) array."""
    if seeds is None:
        seeds = get_seeds(im, max_num_seeds=max_num_seeds,
                          th_seed=th_seed, th_seed_per=th_seed_per,
                          use_percentile=use_percentile,
                          sel_center=sel_center, seed_radius=seed_radius,
                          use_dynamic_th=use_dynamic_th,
                          min_dynamic_seeds=min_num_seeds,
                          remove_hot_pixel=remove_hot_pixel,
                          hot_pixel_th=hot_pixel_th,
                          return_h=False, verbose=verbose,
                          **seed_kwargs)
    fitter = Fitting_v4.iter_fit_seed_points(im, seeds.T, radius_fit=fit_radius)
    fitter._fit_all()                           # firstfit() + repeatfit()
    pfits = fitter._ps_array()                  # == np.array(fitter.ps)
    if len(pfits) > 0:
        centers = pfits[:, 1:4]
        if verbose:
            print(f"-- fitting {len(pfits)} points.")
        if remove_close_pts:
            centers = _drop_close_points(centers, im.shape, close_threshold, verbose)
    else:
        centers = np.array([])
        if verbose:
            print(f"-- no points fitted, return empty array.")
    return centers


def _drop_close_points(centers, shape, close_threshold, verbose=False):
    """fitting.py:319-326: drop NaN centres, centres with another centre within d^2 <
    close_threshold, and centres outside [0, shape]."""
    remove = np.zeros(len(centers), dtype=bool)
    shp = np.array(shape)
    for i, bead in enumerate(centers):
        if np.isnan(bead).any() or np.sum(np.sum((centers - bead) ** 2, axis=1) < close_threshold) > 1:
            remove[i] = True
        if (bead < 0).any() or (bead > shp).any():
            remove[i] = True
    if verbose:
        print(f"-- {np.sum(remove)} points removed, given miminum distance {close_threshold}.")
    return centers[remove == False]


def select_sparse_centers(centers, distance_th=9,
                          distance_norm=np.inf,
                          verbose=False):
    """Greedy selection: keep a centre if it is farther than distance_th (in the given norm)
    from every centre kept so far."""
    kept = []
    for ct in centers:
        if kept:
            d = np.linalg.norm(np.array(kept) - ct[np.newaxis, :], axis=1, ord=distance_norm)
            if (d <= distance_th).any():
                continue
        kept.append(ct)
    if verbose:
        print(f"-- {len(kept)} among {len(centers)} centers are selected by th={distance_th}")
    return np.array(kept)


fit_fov_image.__signature__ = _public_signature()      # the reference's parameters and defaults (+ _stack)
