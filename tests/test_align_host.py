"""CPU: host side of the drift estimation (correction_tools/alignment.py mirror, alignment_tools.py, spot_tools/matching.py)
against the fixture written from the unmodified reference (``python -m oracle.make_golden align`` -> tests/golden/align_r2.npz)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "align_r2.npz"))


def bead_images(g):
    from imageanalysis3_b200.synth import bead_pair
    ref, src, _ = bead_pair(tuple(int(v) for v in g["shape"]), int(g["n"]), tuple(g["drift_planted"]), int(g["seed"]))
    chk = np.array([int(ref.astype(np.uint64).sum()), int(src.astype(np.uint64).sum()), int((ref.astype(np.int64) * 3 + src).std() * 1e6)])
    assert np.array_equal(chk, g["checksum"]), "imageanalysis3_b200.synth.bead_pair no longer produces the fixture's images"
    return ref, src


def test_crops_rough_shift_and_pairing_match_reference(g):
    from imageanalysis3_b200 import alignment_tools
    from imageanalysis3_b200.correction_tools import alignment
    from imageanalysis3_b200.spot_tools import matching
    ref, src = bead_images(g)
    crops = alignment.generate_drift_crops([int(v) for v in g["shape"]])
    assert np.array_equal(crops, g["crops"])
    s = tuple(slice(*r) for r in crops[0])
    rough = alignment_tools.fft3d_from2d(src[s], ref[s], gb=0, max_disp=np.max(src[s].shape) / 2)
    assert np.array_equal(rough, g["crop0_rough"])
    dft, p_t, p_r = matching.find_paired_centers(g["crop0_src_cts"], g["crop0_ref_cts"], rough, cutoff=2., return_paired_cts=True)
    assert np.array_equal(dft, g["crop0_drift_paired"]) and np.array_equal(p_t, g["crop0_paired_tar"]) and np.array_equal(p_r, g["crop0_paired_ref"])
    dft2, k_t, k_r = matching.check_paired_centers(p_t, p_r, outlier_sigma=1.5, return_paired_cts=True)
    assert np.array_equal(dft2, g["crop0_drift_checked"]) and np.array_equal(k_t, g["crop0_kept_tar"]) and np.array_equal(k_r, g["crop0_kept_ref"])
    # the full align_beads call on the same inputs (host only: centres are given)
    d3, t3, r3 = alignment.align_beads(g["crop0_src_cts"], g["crop0_ref_cts"], src[s], ref[s], verbose=False)
    assert np.array_equal(d3, g["crop0_drift_checked"]) and len(t3) == len(g["crop0_kept_tar"])


def test_pairing_is_unique_and_respects_cutoff():
    from imageanalysis3_b200.spot_tools import matching
    ref = np.array([[0., 0, 0], [10, 0, 0], [10, 1.5, 0], [30, 30, 30]])
    tar = np.array([[0.5, 0, 0], [10, 0.7, 0], [50, 50, 50]])
    dft, p_t, p_r, i_t, i_r = matching.find_paired_centers(tar, ref, None, cutoff=1.0, return_paired_cts=True, return_kept_inds=True)
    assert i_t.tolist() == [0] and i_r.tolist() == [0]               # tar 1 has two partners within the cutoff: dropped
    assert np.allclose(dft, [0.5, 0, 0])
    spots = np.concatenate([np.ones((3, 1)), tar, np.zeros((3, 7))], axis=1)     # 11-column spot tables are accepted
    assert np.array_equal(matching.find_paired_centers(spots, ref, None, cutoff=1.0)[1], p_t)


def test_align_image_argument_errors():
    from imageanalysis3_b200.correction_tools import alignment
    im = np.zeros((8, 32, 32), np.uint16)
    kw = dict(correction_args=dict(single_im_size=[8, 32, 32]), verbose=False)
    with pytest.raises(NotImplementedError):
        alignment.align_image(im, im, use_autocorr=True, **kw)
    with pytest.raises(IndexError):
        alignment.align_image(im, im, use_autocorr=False, crop_list=[np.zeros((2, 2))], **kw)
    with pytest.raises(ValueError):
        alignment.align_image(im, im, use_autocorr=False, drift_channel='999', **kw)
    with pytest.raises(IOError):
        alignment.align_image(3.5, im, use_autocorr=False, **kw)
    with pytest.raises(IOError):
        alignment.align_image("not_there.dax", im, use_autocorr=False, **kw)
    with pytest.raises(IndexError):
        alignment.align_image(im, im[:, :16], use_autocorr=False, **kw)
    with pytest.raises(ValueError):
        alignment.generate_drift_crops([8, 32, 32], coord_sel=np.array([4, 40, 10]))
    with pytest.raises(NotImplementedError):
        alignment.align_beads(np.zeros((3, 3)), np.zeros((3, 3)), use_fft=False)
    with pytest.raises(ValueError):
        alignment.align_beads(np.zeros((3, 3)), np.zeros((3, 3)), use_fft=True)
