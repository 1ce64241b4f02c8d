"""Pixel-level translation between two images by FFT cross-correlation of max projections (reference
alignment_tools.py:286-353).  Host code on small 2-D projections: scipy.signal.fftconvolve, the reference's own
third-party layer, so the integer shifts are the reference's."""
import numpy as np


def blurnorm2d(im, gb):
    """im / cv2.blur(im, (gb, gb)) in float32 (reference alignment_tools.py:277-283)"""
    import cv2
    im32 = im.astype(np.float32)
    return im32 / cv2.blur(im32, (gb, gb))


def _standardised(im):
    out = np.array(im, dtype=float)
    out -= np.mean(out)
    out /= np.std(out)
    return out


def fftalign_2d(im1, im2, center=[0, 0], max_disp=150, plt_val=False):
    """integer (tx, ty) maximising the cross-correlation of two 2-D images within max_disp of ``center``
    (reference alignment_tools.py:286-328)"""
    from scipy.signal import fftconvolve
    cor = fftconvolve(_standardised(im1), _standardised(im2[::-1, ::-1]), mode='full')
    n0, n1 = cor.shape
    mid = np.array(center) + np.array([n0, n1]) / 2.
    lo0, hi0 = (int(min(max(mid[0] + s * max_disp, 0), n0)) for s in (-1, 1))
    lo1, hi1 = (int(min(max(mid[1] + s * max_disp, 0), n1)) for s in (-1, 1))
    window = np.zeros_like(cor)
    window[lo0:hi0, lo1:hi1] = 1
    cor = cor * window
    p0, p1 = np.unravel_index(np.argmax(cor), cor.shape)
    t0, t1 = (-np.floor(np.array(cor.shape) / 2) + [p0, p1]).astype(int)
    return t0, t1


def fft3d_from2d(im1, im2, gb=5, max_disp=150):
    """(tz, tx, ty): tx, ty from the z max projections, then tz from the y max projections of the overlapping
    parts (reference alignment_tools.py:330-353); gb > 1 normalises the projections with blurnorm2d first"""
    prep = (lambda p: blurnorm2d(p, gb)) if gb > 1 else (lambda p: p)
    top1, top2 = prep(np.max(im1, 0)), prep(np.max(im2, 0))
    tx, ty = fftalign_2d(top1, top2, center=[0, 0], max_disp=max_disp, plt_val=False)
    sx, sy = top1.shape
    side1 = prep(np.max(im1[:, max(tx, 0):sx + tx, max(ty, 0):sy + ty], axis=-1))
    side2 = prep(np.max(im2[:, max(-tx, 0):sx - tx, max(-ty, 0):sy - ty], axis=-1))
    tz, _ = fftalign_2d(side1, side2, center=[0, 0], max_disp=max_disp, plt_val=False)
    return np.array([tz, tx, ty])
