"""ORACLE (test infrastructure, CPU) -- the compute core of io_tools/load.py correct_fov_image restated.

Not part of the product.  Restates, on in-memory channel stacks (file reading stays with the caller),
  hot-pixel removal        corrections.py:490-510 Remove_Hot_Pixels (called at io_tools/load.py:323-334)
  z-shift correction       corrections.py:479-487 Z_Shift_Correction (io_tools/load.py:336-345)
  bleed-through mixing     io_tools/load.py:347-367
  illumination division    io_tools/load.py:369-381
  drift + chromatic warp   io_tools/load.py:424-459 (scipy.ndimage.map_coordinates, cubic, mode='nearest')
  Gaussian high pass       correction_tools/filter.py:14-19 (io_tools/load.py:487-497)
with numpy / scipy.ndimage, the reference's own third-party layer.  Pinned: oracle/make_golden.py runs the unmodified
reference function (lifted by oracle/ref_loader.load_corrections) on a synthetic .dax and asserts equality.
"""
import numpy as np
from scipy.ndimage import map_coordinates


def remove_hot_pixels(im, dtype=np.uint16, hot_pix_th=0.50, hot_th=4):
    """im float32 (Z, X, Y): columns brighter than hot_th x the mean of their neighbours (np.roll: wraps around, and
    the y + 1 neighbour is taken twice, as in the reference) in more than hot_pix_th of the planes are replaced, in
    np.where order and in place, by the mean of their four neighbours; interior columns only"""
    conv = (np.roll(im, 1, 1) + np.roll(im, -1, 1) + np.roll(im, 1, 2) + np.roll(im, 1, 2)) / 4
    hot2d = np.sum(im > hot_th * conv, 0)
    cand = np.where(hot2d > hot_pix_th * np.shape(im)[0])
    if len(cand[0]) == 0:
        return im
    nim = im.copy()
    for x, y in zip(cand[0], cand[1]):
        if x > 0 and y > 0 and x < im.shape[1] - 1 and y < im.shape[2] - 1:
            nim[:, x, y] = (nim[:, x + 1, y] + nim[:, x - 1, y] + nim[:, x, y + 1] + nim[:, x, y - 1]) / 4
    return nim.astype(dtype)


def z_shift_correction(im, dtype=np.uint16):
    """im float32 (Z, X, Y): every plane divided by its median and multiplied by the stack's median (corrections.py:479-487)"""
    return (im / np.median(im, axis=(1, 2))[:, np.newaxis, np.newaxis] * np.median(im)).astype(dtype)


def gaussian_high_pass(image, sigma=5, truncate=2):
    """image - gaussian_filter(image, sigma, mode='nearest', truncate) where positive, else 0 (correction_tools/filter.py:14-19)"""
    from scipy.ndimage import gaussian_filter
    low = gaussian_filter(image, sigma, mode='nearest', truncate=truncate)
    high = image - low
    high[low > image] = 0
    return high


def correct_stacks(ims, load_channels, sel_channels, corr_channels, drift=None, hot_pixel_corr=True, hot_pixel_th=4, z_shift_corr=False,
                   gaussian_highpass=False, gauss_sigma=3, gauss_truncate=2,
                   illumination_corr=True, illumination_profile=None, bleed_corr=True, bleed_profile=None,
                   chromatic_ref_channel='647', chromatic_corr=True, chromatic_profile=None, warp_image=True,
                   output_dtype=np.uint16, verbose=True):
    """ims: list of (Z, X, Y) uint16 stacks, one per channel of load_channels (the order correct_fov_image builds:
    corr_channels first when bleed-through applies, then the remaining selected ones) -> list of corrected stacks
    for sel_channels.  ``verbose``: the reference's warp code sits INSIDE its ``if verbose:`` block
    (io_tools/load.py:437-459 are indented under :436), so images are only warped when verbose is true."""
    ims = [np.array(im) for im in ims]
    if hot_pixel_corr:
        ims = [remove_hot_pixels(im.astype(np.float32), dtype=output_dtype, hot_th=hot_pixel_th) for im in ims]
    if z_shift_corr:
        ims = [z_shift_correction(im.astype(np.float32), dtype=output_dtype) for im in ims]
    overlap = [ch for ch in corr_channels if ch in sel_channels]
    if len(overlap) > 0 and bleed_corr:
        bld = [ims[load_channels.index(ch)] for ch in corr_channels]
        mixed = [np.sum([im * bleed_profile[i, j] for j, im in enumerate(bld)], axis=0) for i in range(len(corr_channels))]
        for nim, ch in zip(mixed, corr_channels):
            nim[nim > np.iinfo(output_dtype).max] = np.iinfo(output_dtype).max
            nim[nim < np.iinfo(output_dtype).min] = np.iinfo(output_dtype).min
            ims[load_channels.index(ch)] = nim.astype(output_dtype)
    if illumination_corr:
        ims = [(im.astype(np.float32) / illumination_profile[ch][np.newaxis, :]).astype(output_dtype) for im, ch in zip(ims, load_channels)]
    drift = np.zeros(3, dtype=np.float32) if drift is None else np.array(drift, dtype=np.float32)
    chroma = [ch for ch in corr_channels if ch in sel_channels and ch != chromatic_ref_channel]
    if warp_image:
        for ch in sel_channels:
            if ((chromatic_corr and ch in chroma) or drift.any()) and verbose:
                im = ims[load_channels.index(ch)]
                coords = np.stack(np.meshgrid(*[np.arange(s) for s in im.shape])).transpose((0, 2, 1, 3))
                if chromatic_corr and ch in chroma and chromatic_profile[ch] is not None:
                    coords = coords + chromatic_profile[ch]
                if drift.any():
                    coords = coords - drift[:, np.newaxis, np.newaxis, np.newaxis]
                ims[load_channels.index(ch)] = map_coordinates(im, coords.reshape(3, -1), mode='nearest').astype(output_dtype).reshape(im.shape)
    if gaussian_highpass:
        ims = [gaussian_high_pass(im, gauss_sigma, gauss_truncate) for im in ims]
    return [ims[load_channels.index(ch)].astype(output_dtype).copy() for ch in sel_channels]
