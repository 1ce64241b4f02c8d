"""GPU: drift estimation in its bead-fitting mode (correction_tools/alignment.py align_image(use_autocorr=False): the device
fit_fov_image on crops + the host pairing) against the unmodified reference's results (tests/golden/align_r2.npz), directly
and through correct_fov_image(calculate_drift=True) on two .dax movies."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_align_host import bead_images

pytestmark = pytest.mark.gpu

TOL_DRIFT_PX = 1e-3          # fitted centres agree to float32 on well-posed beads; a drift is a mean over tens of pairs


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "align_r2.npz"))


def test_align_image_matches_reference(lib, g):
    from imageanalysis3_b200.correction_tools import alignment
    from imageanalysis3_b200.spot_tools.fitting import fit_fov_image, select_sparse_centers
    ref, src = bead_images(g)
    shape = [int(v) for v in g["shape"]]
    kw = dict(use_autocorr=False, correction_args=dict(single_im_size=shape), verbose=False)
    d0, f0 = alignment.align_image(src, ref, **kw)
    assert f0 == int(g["flag_default"]) == 0 and np.abs(d0 - g["drift_default"]).max() <= TOL_DRIFT_PX
    assert np.abs(d0 + g["drift_planted"]).max() < 0.05                  # and it is the planted drift (sign: cross-correlation convention)
    d1, f1 = alignment.align_image(src, ref, drift_diff_th=1e-4, **kw)      # no three crops agree that closely: the sub-optimal branch
    assert f1 == int(g["flag_suboptimal"]) == 1 and np.abs(d1 - g["drift_suboptimal"]).max() <= TOL_DRIFT_PX
    d2, f2 = alignment.align_image(src, ref, crop_list=g["crops"][[5, 2, 7]], min_good_drifts=2, match_distance_th=1.5,
                                   fitting_args=dict(max_num_seeds=12), **kw)
    assert f2 == int(g["flag_custom"]) and np.abs(d2 - g["drift_custom"]).max() <= TOL_DRIFT_PX
    # the bead centres of one crop, as the reference's own fit_fov_image found them
    s = tuple(slice(*r) for r in g["crops"][0])
    cts = select_sparse_centers(fit_fov_image(np.ascontiguousarray(src[s]), '488', verbose=False, **alignment._default_align_fitting_args)[:, 1:4], 2.)
    assert cts.shape == g["crop0_src_cts"].shape and np.abs(cts - g["crop0_src_cts"]).max() <= 1e-3


def test_correct_fov_image_estimates_the_drift_from_files(lib, g, tmp_path):
    from imageanalysis3_b200.io_tools import load
    from imageanalysis3_b200.synth import synth
    from oracle.make_golden import align_files
    ref, src = bead_images(g)
    shape = tuple(int(v) for v in g["shape"])
    other_ref, other_src = (synth(shape, 40, int(sd)) for sd in g["file_other_seeds"])
    src_dax, ref_dax = align_files(str(tmp_path), ref, src, other_ref, other_src)
    kw = dict(single_im_size=list(shape), all_channels=['647', '488'], num_buffer_frames=2, num_empty_frames=0, drift_channel='488',
              ref_filename=ref_dax, corr_channels=['647'], correction_folder=str(tmp_path), bleed_corr=False, chromatic_corr=False, return_drift=True)
    ims, drift, flag = load.correct_fov_image(src_dax, ['647'], calculate_drift=True, use_autocorr=False, verbose=True, **kw)
    assert flag == int(g["file_flag"]) and np.abs(drift - g["file_drift"]).max() <= TOL_DRIFT_PX
    got, want = ims[0][4:12, 64:192, 64:192].astype(np.int64), g["file_im_647_slab"].astype(np.int64)
    # the warp runs with OUR drift, which differs from the reference's in the 6th decimal: a few roundings may flip
    d = np.abs(got - want)
    assert d.max() <= 1 and (d > 0).mean() <= 1e-3, (d.max(), (d > 0).mean())
    # with the reference's drift given, the corrected image is the reference's
    ims2, d2, f2 = load.correct_fov_image(src_dax, ['647'], drift=None, verbose=False, **{**kw, 'ref_filename': None})
    assert f2 == 0 and not d2.any()
    with pytest.raises(NotImplementedError):
        load.correct_fov_image(src_dax, ['647'], calculate_drift=True, use_autocorr=True, verbose=False, **kw)
