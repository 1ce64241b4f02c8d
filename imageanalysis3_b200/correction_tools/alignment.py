"""Drift between two bead images (reference correction_tools/alignment.py), SURVEY 8(f) rank 3.

``align_image`` has two modes in the reference.  ``use_autocorr=False`` fits the beads of up to eight crops of both
images with ``fit_fov_image`` (the hot path: on the device here), pairs them (``align_beads``: FFT cross-correlation of
max projections for the pixel-level shift, then unique nearest pairs and a Delaunay-neighbour outlier test) and averages
the pair differences -- that mode is built, with the reference's crop schedule and stopping rule.  ``use_autocorr=True``
(the reference's default) calls ``skimage.registration.phase_cross_correlation``; scikit-image is not installed where this
was built, so neither the reference nor a restatement of it could be pinned, and that mode raises NotImplementedError.
"""
import os
import time

import numpy as np

from .. import _allowed_colors, _correction_folder, _image_size

_num_buffer_frames, _num_empty_frames = 10, 0        # reference __init__.py

_default_align_corr_args = {
    'single_im_size': _image_size,
    'num_buffer_frames': _num_buffer_frames,
    'num_empty_frames': _num_empty_frames,
    'correction_folder': _correction_folder,
    'illumination_corr': True,
    'bleed_corr': False,
    'chromatic_corr': False,
    'z_shift_corr': False,
    'hot_pixel_corr': True,
    'normalization': False,
}

_default_align_fitting_args = {
    'th_seed': 300,
    'th_seed_per': 95,
    'use_percentile': False,
    'use_dynamic_th': True,
    'min_dynamic_seeds': 10,
    'max_num_seeds': 200,
}


def _find_boundary(_ct, _radius, _im_size):
    return np.array([[max(c - _radius, 0), min(c + _radius, sz)] for c, sz in zip(_ct, _im_size)], dtype=int)


def generate_drift_crops(single_im_size=_image_size, coord_sel=None, drift_size=None):
    """eight (3, 2) crop limits around ``coord_sel`` (default: the image centre): the four quadrant centres, then the
    four edge midpoints, each ``drift_size`` (default: a quarter of the largest dimension) wide
    (reference correction_tools/alignment.py:87-136)"""
    size = np.array(single_im_size)
    if coord_sel is None:
        coord_sel = np.array(size / 2, dtype=int)
    if coord_sel[-2] >= size[-2] or coord_sel[-1] >= size[-1]:
        raise ValueError(f"wrong input coord_sel:{coord_sel}, should be smaller than single_im_size:{single_im_size}")
    if drift_size is None:
        drift_size = int(np.max(size) / 4)
    z = coord_sel[-3] / 2
    x0, x1, x2 = coord_sel[-2] / 2, coord_sel[-2], (coord_sel[-2] + size[-2]) / 2
    y0, y1, y2 = coord_sel[-1] / 2, coord_sel[-1], (coord_sel[-1] + size[-1]) / 2
    centres = [(z, x0, y0), (z, x2, y2), (z, x2, y0), (z, x0, y2), (z, x1, y0), (z, x1, y2), (z, x0, y1), (z, x2, y1)]
    return np.array([_find_boundary(np.array(c), _radius=drift_size / 2, _im_size=single_im_size) for c in centres])


def align_beads(tar_cts, ref_cts, tar_im=None, ref_im=None, use_fft=True, fft_filt_size=0, match_distance_th=2.,
                check_paired_cts=True, outlier_sigma=1.5, return_paired_cts=True, verbose=True):
    """drift (tar - ref) from two lists of bead centres (reference correction_tools/alignment.py:139-217); use_fft: the
    pixel-level shift comes from the images, then centres are paired uniquely within match_distance_th"""
    from ..alignment_tools import fft3d_from2d
    from ..spot_tools.matching import check_paired_centers, find_paired_centers
    tar, ref = np.array(tar_cts), np.array(ref_cts)
    if not use_fft:
        raise NotImplementedError("align_beads(use_fft=False) (alignment_tools.translation_align_pts) is not built")
    if tar_im is None or ref_im is None:
        raise ValueError("both tar_im and ref_im should be given if use FFT!")
    if np.shape(tar_im) != np.shape(ref_im):
        raise IndexError(f"tar_im shape:{np.shape(tar_im)} should match ref_im shape:{np.shape(ref_im)}")
    rough = fft3d_from2d(tar_im, ref_im, gb=fft_filt_size, max_disp=np.max(np.shape(tar_im)) / 2)
    drift, p_tar, p_ref = find_paired_centers(tar, ref, rough, cutoff=float(match_distance_th), return_paired_cts=True, verbose=verbose)
    if verbose:
        print("before check:", drift, len(p_ref))
    if check_paired_cts and len(p_ref) > 3:
        drift, p_tar, p_ref = check_paired_centers(p_tar, p_ref, outlier_sigma=outlier_sigma, return_paired_cts=True, verbose=verbose)
    return (drift, p_tar, p_ref) if return_paired_cts else (drift,)


def _bead_image(im, channel, all_channels, correction_args, verbose):
    if isinstance(im, np.ndarray):
        return im
    if isinstance(im, str):
        if not os.path.isfile(im) or im.split('.')[-1] != 'dax':
            raise IOError(f"input image: {im} should be a .dax file!")
        from ..io_tools.load import correct_fov_image
        return correct_fov_image(im, [channel], all_channels=all_channels, calculate_drift=False, return_drift=False,
                                 verbose=verbose, **correction_args)[0][0]
    raise IOError(f"Wrong input file type, {type(im)} should be .dax file or np.ndarray")


def align_image(src_im, ref_im, crop_list=None, use_autocorr=True, precision_fold=100,
                min_good_drifts=3, drift_diff_th=1., all_channels=_allowed_colors, ref_all_channels=None,
                drift_channel='488', correction_args={}, fitting_args={}, match_distance_th=2.,
                verbose=True, detailed_verbose=False):
    """(drift, flag) of ``src_im`` against ``ref_im`` (arrays or .dax files), reference correction_tools/alignment.py:
    527-696: crops are processed in order until ``min_good_drifts`` of their drifts lie within ``drift_diff_th`` of the
    running mean (flag 0); otherwise the mean of the closest pair of drifts and the one nearest to both (flag 1)"""
    from ..sharding import map_stacks
    from ..spot_tools.fitting import fit_fov_image, select_sparse_centers
    corr_args = dict(_default_align_corr_args)
    corr_args.update(correction_args)
    fit_args = dict(_default_align_fitting_args)
    fit_args.update(fitting_args)
    if crop_list is None:
        crop_list = generate_drift_crops(corr_args['single_im_size'])
    for crop in crop_list:
        if np.shape(np.array(crop)) != (3, 2):
            raise IndexError("crop should be 3x2 np.ndarray.")
    every = [str(ch) for ch in all_channels]
    channel = str(drift_channel)
    if channel not in all_channels:
        raise ValueError(f"bead channel {channel} not exist in all channels given:{every}")
    ref_every = every if ref_all_channels is None else [str(ch) for ch in ref_all_channels]
    if use_autocorr:
        raise NotImplementedError("align_image(use_autocorr=True) needs skimage.registration.phase_cross_correlation, which could "
                                  "not be pinned where this framework was built; use_autocorr=False (bead fitting) runs on the device")
    if verbose:
        print("-- start aligning", "given source image" if isinstance(src_im, np.ndarray) else f"file {src_im}",
              "to", "given reference image." if isinstance(ref_im, np.ndarray) else f"reference file:{ref_im}.")
    src = _bead_image(src_im, channel, every, corr_args, detailed_verbose)
    ref = _bead_image(ref_im, channel, ref_every, corr_args, detailed_verbose)
    if np.shape(src) != np.shape(ref):
        raise IndexError(f"shape of target image:{np.shape(src)} and reference image:{np.shape(ref)} doesnt match!")
    drifts, result, flag = [], None, 0
    for i, crop in enumerate(crop_list):
        t0 = time.time()
        sel = tuple(slice(*np.array(c, dtype=int)) for c in crop)
        sim, rim = np.ascontiguousarray(src[sel]), np.ascontiguousarray(ref[sel])
        # the two fits of a crop are independent: both in flight at once (each is latency-bound on its own)
        spots = map_stacks(lambda im: fit_fov_image(im, channel, verbose=detailed_verbose, **fit_args), [sim, rim], inflight=2)
        src_cts = select_sparse_centers(spots[0][:, 1:4], match_distance_th)
        ref_cts = select_sparse_centers(spots[1][:, 1:4], match_distance_th, verbose=detailed_verbose)
        dft, _, _ = align_beads(src_cts, ref_cts, sim, rim, use_fft=True, match_distance_th=match_distance_th,
                                return_paired_cts=True, verbose=detailed_verbose)
        drifts.append(dft * -1)                       # bead centres move opposite to the cross-correlation convention
        if verbose:
            print(f"-- drift {i}: {np.around(drifts[-1], 2)} in {time.time() - t0:.3f}s.")
        mean = np.nanmean(drifts, axis=0)
        if len(drifts) >= min_good_drifts:
            kept = np.where(np.linalg.norm(drifts - mean, axis=1) <= drift_diff_th)[0]
            if len(kept) >= min_good_drifts:
                result = np.nanmean(np.array(drifts)[kept], axis=0)
                if verbose:
                    print(f"--- drifts for crops:{kept} pass the thresold, exit cycle.")
                break
    if result is None:
        if verbose:
            print("-- return a sub-optimal drift")
        from scipy.spatial.distance import pdist, squareform
        arr = np.array(drifts)
        dist = squareform(pdist(arr))
        np.fill_diagonal(dist, np.inf)
        pair = np.array(np.unravel_index(np.argmin(dist), np.shape(dist)))
        chosen = list(arr[pair]) + [arr[np.argmin(dist[:, pair].sum(1))]]
        result = np.nanmean(chosen, axis=0)
        flag = 1
    return result, flag
