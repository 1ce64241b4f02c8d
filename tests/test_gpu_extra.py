"""GPU parity for the argument paths and functions that had no test in round 1, against fixtures produced by the
unmodified reference (oracle/make_golden.py extra -> tests/golden/extra_r2.npz): percentile thresholds,
find_matched_seeds, standalone Fitting_v3.GaussianFit, other input dtypes / 2-D images in get_seeds, and the
alternative seeders of External/Fitting_v4.py (get_seed_points_base_v2: cv2.blur based; get_seed_points_base /
fft_gaussian_fast: FFT based, compared by tolerance)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_spots_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def extra():
    return np.load(os.path.join(GOLDEN, "extra_r2.npz"))


def test_percentile_thresholds_match_reference(lib, extra):
    """use_percentile / seed_by_per: scipy.stats.scoreatpercentile over the whole image, read off the device
    histogram (spot_tools/fitting.py:75-76, visual_tools.py:1808-1811)"""
    from imageanalysis3_b200 import visual_tools as vt
    from imageanalysis3_b200.spot_tools.fitting import fit_fov_image, get_seeds
    g = extra
    im = g["im"]
    assert np.array_equal(get_seeds(im, use_percentile=True, th_seed_per=95), g["seeds_percentile95"])
    assert np.array_equal(get_seeds(im, use_percentile=True, th_seed_per=99.5, return_h=True), g["seeds_percentile99_h"])
    got = fit_fov_image(im, '647', use_percentile=True, th_seed_per=99.5, max_num_seeds=30, verbose=False)
    assert_spots_close(got, g["fov_percentile"], "fit_fov_image(use_percentile)", g["fov_percentile_comparable"])
    assert np.array_equal(vt.get_seed_in_distance(im, center=None, seed_by_per=True, th_seed_percentile=99.5, return_h=True), g["legacy_by_per"])
    assert np.array_equal(vt.get_seed_in_distance(im, center=[10, 40, 50], seed_by_per=True, th_seed_percentile=99.5, num_seeds=5, return_h=True),
                          g["legacy_by_per_center"])
    st = lib.Stack(im)
    assert np.array_equal(st.histogram(), np.bincount(im.ravel(), minlength=65536).astype(np.uint64))


def test_find_matched_seeds_matches_reference(lib, extra):
    from imageanalysis3_b200 import visual_tools as vt
    g = extra
    for tag, kw in (("default", {}), ("unique_d5", dict(keep_unique=True, search_distance=5)), ("th600", dict(th_seed=600, search_distance=2))):
        m, f = vt.find_matched_seeds(g["im"], g["matched_ref"], verbose=False, **kw)
        assert np.array_equal(f, g[f"matched_{tag}_found"]), tag
        assert m.shape == g[f"matched_{tag}"].shape and np.array_equal(m, g[f"matched_{tag}"]), tag
    with pytest.raises(TypeError):
        vt.find_matched_seeds([[1, 2], [3, 4]], g["matched_ref"])


def test_get_seeds_other_dtypes_and_2d(lib, extra):
    """int32 / int64 / float64 stacks and 2-D images (the reference's filters are dtype- and rank-generic)"""
    from imageanalysis3_b200.spot_tools.fitting import fit_fov_image, get_seeds
    g = extra
    im = g["im"]
    assert np.array_equal(get_seeds(im.astype(np.int32), th_seed=300), g["seeds_i32"])
    assert np.array_equal(get_seeds(im.astype(np.int64), th_seed=300, return_h=True, max_num_seeds=30), g["seeds_i64_h"])
    imf64 = im.astype(np.float64) / 2.5
    assert np.array_equal(get_seeds(imf64, th_seed=120), g["seeds_f64"])
    got = fit_fov_image(imf64, '647', th_seed=120, max_num_seeds=20, verbose=False)
    assert got.dtype == g["fov_f64"].dtype
    assert_spots_close(got, g["fov_f64"], "float64 image")
    assert np.array_equal(get_seeds(im[10], th_seed=300, remove_hot_pixel=False), g["seeds_2d"])
    assert np.array_equal(get_seeds(im[10].astype(np.float32), th_seed=300, remove_hot_pixel=False, return_h=True), g["seeds_2d_f32_h"])
    with pytest.raises(IndexError):
        get_seeds(im[10], th_seed=300)                  # the reference's hot-pixel filter needs a 3-D stack
    with pytest.raises(NotImplementedError):
        get_seeds(im.astype(np.int32) - 400, th_seed=300)


def test_standalone_v3_gaussianfit(lib, extra):
    """Fitting_v3.GaussianFit on its own (External/Fitting_v3.py:50-257), without and with the width prior"""
    from imageanalysis3_b200.External import Fitting_v3
    g = extra
    X, im = g["gf3_X"], g["im"]
    # the fixture's window is int(c) + offsets (-5..4) around the brightest seed c, which lies inside the image
    c = [float(X[0].min() + 5), float(X[1].min() + 5), float(X[2].min() + 5)]
    for ws in (0, 1000):
        obj = Fitting_v3.GaussianFit(im[X[0], X[1], X[2]], X, center=c, delta_center=2.5, weight_sigma=ws)
        obj.fit()
        assert obj.success
        assert_spots_close([obj.p], [g[f"gf3_ws{ws}_p"]], f"v3 GaussianFit ws={ws}")
        assert np.allclose(obj.get_im(), g[f"gf3_ws{ws}_rec"], rtol=1e-4, atol=1e-3)


def test_fitting_v4_alternative_seeders(lib, extra):
    """a12: get_seed_points_base_v2 (exact: the blur is bit-identical to cv2.blur, the fixtures' seeds and heights are
    reproduced exactly, std to float32 rounding) and get_seed_points_base / fft_gaussian_fast (direct FP64 convolution
    against the reference's single-precision FFT: 1e-5 relative)"""
    from imageanalysis3_b200.External import Fitting_v4
    g = extra
    im = g["im"]
    imf = im.astype(np.float32)
    if "v2_u16_centers" in g.files:
        for tag, arr, kw in (("u16", im, dict(th_seed=3.)), ("f32_g7_f5", imf, dict(gfilt_size=7, filt_size=5, th_seed=2.5)),
                             ("u16_top10", im, dict(th_seed=3., max_num=10))):
            cen, std = Fitting_v4.get_seed_points_base_v2(arr, **kw)
            want = g[f"v2_{tag}_centers"]
            assert cen.shape == want.shape and cen.dtype == want.dtype, tag
            assert np.array_equal(cen, want), tag
            assert abs(float(std) - float(g[f"v2_{tag}_std"])) <= 2e-6 * float(g[f"v2_{tag}_std"]), tag
    blur = Fitting_v4.fft_gaussian_fast(imf, gaus=[2.5, 5, 5])
    assert blur.shape == g["fftg_5"].shape
    assert np.abs(blur - g["fftg_5"]).max() <= 1e-5 * np.abs(g["fftg_5"]).max()
    cen, std = Fitting_v4.get_seed_points_base(imf, gfilt_size=2.5, th_seed=3.)
    want = g["lr_centers"]
    assert abs(std - float(g["lr_std"])) <= 1e-5 * float(g["lr_std"])
    assert cen.shape == want.shape
    # same voxels; heights to 1e-5; order by height may swap only between heights closer than that
    key = lambda c: sorted(map(tuple, c[:3].T.astype(np.int64)))
    assert key(cen) == key(want)
    assert np.allclose(np.sort(cen[3]), np.sort(want[3]), rtol=0, atol=1e-5)
    cen, _ = Fitting_v4.get_seed_points_base(imf, gfilt_size=2.5, th_seed=4., filt_size=5, max_num=12)
    assert key(cen) == key(g["lr_centers_f5_top12"])
    # host helpers of the module
    s, si = Fitting_v4.to_sigmas(0.3, -0.2, 1.3, 1.9, 2.1)
    assert np.allclose(s @ si, np.eye(3), atol=1e-12)
