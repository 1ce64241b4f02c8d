"""Drop-in for the legacy seeding wrappers of the reference's ``visual_tools.py``
(get_STD_centers :260-345, get_seed_points_base :348-381, get_seed_in_distance :1775-1870,
find_matched_seeds :3081-3139, select_sparse_centers :3142-3162), which feed ``Fitting_v3``.
Both Gaussian blurs, the rank filters and the candidate mask run on the B200
(libia3b200.so, variant 1 of the seed stage); thresholds above the device floor, the hot-pixel
multiplicity filter and the top-N selection are replayed here on the candidate list.
"""
import numpy as np

from . import _lib, _sigma_zxy  # noqa: F401
from .External import Fitting_v3
from .spot_tools.fitting import _device_image, _gauss_half_kernel, _drop_close_points, select_sparse_centers as _ssc


def _legacy_candidates(im, gfilt_size, background_gfilt_size, filt_size, th_floor):
    """device: candidates (z, x, y, h) in C order with h = max_filt - min_filt > th_floor"""
    if not isinstance(im, np.ndarray) or im.ndim != 3:
        raise NotImplementedError("the GPU seed stage needs a 3D numpy stack")
    dev = _device_image(im)
    if dev.dtype != np.uint16:
        raise NotImplementedError("legacy seeder on the GPU supports integer (uint16) stacks")
    stack = _lib.Stack(dev)
    # strict '>' on integers: h > th  <=>  h >= floor(th) + 1
    h_min = np.floor(float(th_floor)) + 1.0
    zxy, h, _ = stack.seed_candidates(
        _gauss_half_kernel(gfilt_size) if gfilt_size else None,
        _gauss_half_kernel(background_gfilt_size) if background_gfilt_size else None,
        int(filt_size), 1, 0.0, h_min)
    return zxy.astype(np.int64), h.astype(np.int64)


def _legacy_select(zxy, h, th_seed, hot_pix_th):
    keep = h > th_seed
    zxy, h = zxy[keep], h[keep]
    if hot_pix_th > 0:
        key = zxy[:, 1] * (int(zxy[:, 2].max()) + 1 if len(zxy) else 1) + zxy[:, 2]
        _, inv, cts = np.unique(key, return_inverse=True, return_counts=True)
        ok = ~(cts[inv] > hot_pix_th) if len(key) else np.zeros(0, dtype=bool)
        zxy, h = zxy[ok], h[ok]
    return zxy, h


def get_seed_points_base(im, gfilt_size=0.75, background_gfilt_size=10, filt_size=3,
                         th_seed=300, hot_pix_th=0, return_h=False):
    """Base function to do seeding -> (3|4, N) int64 [z, x, y(, h)] in C order"""
    zxy, h = _legacy_candidates(im, gfilt_size, background_gfilt_size, filt_size, th_seed)
    zxy, h = _legacy_select(zxy, h, th_seed, hot_pix_th)
    if return_h:
        return np.array([zxy[:, 0], zxy[:, 1], zxy[:, 2], h])
    return np.array([zxy[:, 0], zxy[:, 1], zxy[:, 2]])


def get_seed_in_distance(im, center=None, num_seeds=0, seed_radius=30,
                         gfilt_size=0.75, background_gfilt_size=10, filt_size=3,
                         seed_by_per=False, th_seed_percentile=95,
                         th_seed=300,
                         dynamic=True, dynamic_iters=10, min_dynamic_seeds=2,
                         distance_to_edge=1, hot_pix_th=4,
                         return_h=False, verbose=False):
    """Seeds within seed_radius of a centre (or of the whole image when center is None),
    brightest first, at most num_seeds (0 = all)."""
    from scipy.spatial.distance import cdist
    from scipy.stats import scoreatpercentile
    if center is not None and len(center) != 3:
        raise ValueError('wrong input dimension of center!')
    _dim = np.shape(im)
    if seed_by_per:
        if isinstance(im, np.ndarray) and im.dtype == np.uint16 and im.ndim == 3:
            # order statistics of the whole image from the device histogram (exactly scipy's numbers)
            from .spot_tools.fitting import _score_from_counts
            _st = _lib.Stack(im)
            _counts = _st.histogram()
            _st.close()
            _th_seed = _score_from_counts(_counts, th_seed_percentile) - _score_from_counts(_counts, 100 - th_seed_percentile)
        else:
            ints = im[np.isnan(im) == False].astype(float)
            _th_seed = scoreatpercentile(ints, th_seed_percentile) - scoreatpercentile(ints, 100 - th_seed_percentile)
    else:
        _th_seed = th_seed
    if verbose:
        print(f"-- seeding with threshold: {_th_seed}, per={th_seed_percentile}")
    if center is not None:
        _center = np.array(center, dtype=float)
        _limits = np.zeros([2, 3], dtype=int)
        _limits[0, 1:] = np.array([np.max([x, y]) for x, y in zip(np.zeros(2), _center[1:] - seed_radius)], dtype=int)
        _limits[0, 0] = np.array(np.max([0, _center[0] - seed_radius / 2]), dtype=int)
        _limits[1, 1:] = np.array([np.min([x, y]) for x, y in zip(_dim[1:], _center[1:] + seed_radius)], dtype=int)
        _limits[1, 0] = np.array(np.min([_dim[0], _center[0] + seed_radius / 2]), dtype=int)
        _local_center = _center - _limits[0]
        _cim = np.ascontiguousarray(im[_limits[0, 0]:_limits[1, 0], _limits[0, 1]:_limits[1, 1], _limits[0, 2]:_limits[1, 2]])
        if dynamic:
            ratios = np.linspace(1, 1 / dynamic_iters, dynamic_iters)
            # one device pass at the lowest threshold, then replay the descent on the candidates
            zxy_all, h_all = _legacy_candidates(_cim, gfilt_size, background_gfilt_size, filt_size,
                                                min(_th_seed * r for r in ratios))
            for _dy_ratio in ratios:
                zxy, h = _legacy_select(zxy_all, h_all, _th_seed * _dy_ratio, hot_pix_th)
                cand = np.array([zxy[:, 0], zxy[:, 1], zxy[:, 2], h])
                dist = cdist(cand[:3].transpose(), _local_center[np.newaxis, :3]).transpose()[0]
                _seeds = cand[:, dist < seed_radius]
                _seeds[:3, :] += _limits[0][:, np.newaxis]
                if num_seeds > 0 and _seeds.shape[1] >= min(num_seeds, min_dynamic_seeds):
                    break
                elif num_seeds == 0 and _seeds.shape[1] >= min_dynamic_seeds:
                    break
        else:
            _seeds = get_seed_points_base(_cim, gfilt_size=gfilt_size, filt_size=filt_size,
                                          th_seed=th_seed, hot_pix_th=hot_pix_th, return_h=True)
    else:
        _seeds = get_seed_points_base(im, gfilt_size=gfilt_size, filt_size=filt_size,
                                      th_seed=_th_seed, hot_pix_th=hot_pix_th, return_h=True)
    if _seeds.shape[1] > 1:
        order = np.argsort(_seeds[-1])
        _seeds = _seeds[:, np.flipud(order[-num_seeds:])]
    if not return_h:
        return _seeds[:3].transpose()
    return _seeds[:4].transpose()


def get_STD_centers(im, seeds=None, th_seed=150,
                    dynamic=False, seed_by_per=False, th_seed_percentile=95,
                    min_num_seeds=1,
                    remove_close_pts=True, close_threshold=0.1, fit_radius=5,
                    sort_by_h=False, save=False, save_folder='', save_name='',
                    plt_val=False, force=False, verbose=False):
    """Fit beads of one image (legacy seeder + Fitting_v3 firstfit) -> (n, 3) centres"""
    import os
    import pickle
    if not force and os.path.exists(save_folder + os.sep + save_name) and save_name != '':
        if verbose:
            print("- loading file:,", save_folder + os.sep + save_name)
        beads = pickle.load(open(save_folder + os.sep + save_name, 'rb'))
        if verbose:
            print("--", len(beads), " of beads loaded.")
        return beads
    if seeds is None:
        seeds = get_seed_in_distance(im, center=None, dynamic=dynamic,
                                     th_seed_percentile=th_seed_percentile,
                                     seed_by_per=seed_by_per,
                                     min_dynamic_seeds=min_num_seeds,
                                     gfilt_size=0.75, filt_size=3,
                                     th_seed=th_seed, hot_pix_th=4, verbose=verbose)
    fitter = Fitting_v3.iter_fit_seed_points(im, seeds.T, radius_fit=5)
    fitter.firstfit()
    pfits = fitter.ps
    if len(pfits) > 0:
        if sort_by_h:
            order = np.argsort(np.array(pfits)[:, 0])
            beads = np.array(pfits)[np.flipud(order), 1:4]
        else:
            beads = np.array(pfits)[:, 1:4]
        if remove_close_pts:
            beads = _drop_close_points(beads, im.shape, close_threshold, verbose)
    else:
        beads = None
    if verbose:
        print(f"- fitting {len(pfits)} points")
    if save:
        if not os.path.exists(save_folder):
            os.makedirs(save_folder)
        if verbose:
            print("-- saving fitted spots to", save_folder + os.sep + save_name)
        pickle.dump(beads[:, -3:], open(save_folder + os.sep + save_name, 'wb'))
    return beads


def find_matched_seeds(im, ref_centers, search_distance=3,
                       gfilt_size=0.75, background_gfilt_size=10, filt_size=3,
                       dynamic=False, th_seed_percentile=95, th_seed=200,
                       keep_unique=False, verbose=True):
    """Seeds of `im` lying within search_distance of the given reference centres."""
    if not isinstance(im, np.ndarray) and not isinstance(im, np.memmap):
        raise TypeError(f"Wrong input data type for im, should be np.ndarray or memmap, {type(im)} given!")
    ref_centers = np.array(ref_centers)[:, :3]
    if verbose:
        print(f"- find seeds paired with {len(ref_centers)} centers in given image")
    _seeds = get_seed_in_distance(im, center=None, gfilt_size=gfilt_size,
                                  background_gfilt_size=background_gfilt_size,
                                  filt_size=filt_size, dynamic=dynamic,
                                  th_seed_percentile=th_seed_percentile,
                                  th_seed=th_seed, return_h=True)
    matched, found = [], []
    for ct in ref_centers:
        d = np.linalg.norm(_seeds[:, :3] - ct[np.newaxis, :], axis=1)
        idx, = np.where(d < search_distance)
        if len(idx) == 1:
            matched.append(_seeds[idx[0], :3]); found.append(True)
        elif len(idx) > 1 and not keep_unique:
            cand = _seeds[idx, :]
            matched.append(cand[np.argsort(cand[:, -1])[-1], :3]); found.append(True)
        else:
            found.append(False)
    matched = np.array(matched)
    found = np.array(found, dtype=bool)
    if verbose:
        print(f"-- {len(matched)} paired seeds are found. ")
    return matched, found


def select_sparse_centers(centers, distance_th=9, distance_norm=np.inf):
    return _ssc(centers, distance_th=distance_th, distance_norm=distance_norm)
