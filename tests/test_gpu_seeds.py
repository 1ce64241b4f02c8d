"""GPU parity, seed stage: the CUDA path (through the C ABI) against the oracle and the committed
fixtures of the unmodified reference.  Bar: bit-exact (integer work)."""
import numpy as np
import pytest

from oracle import seed_oracle
from oracle.make_golden import SEED_CASES

pytestmark = pytest.mark.gpu


def _half(sigma):
    w = seed_oracle.gaussian_weights(sigma)
    return np.ascontiguousarray(w[len(w) // 2:])


@pytest.mark.parametrize("sigma", [0.75, 7.5, 10.0, 2.3, 1.0])
def test_gaussian_volumes_bit_exact_u16(lib, golden_seeds, sigma):
    im = golden_seeds["im"]
    st = lib.Stack(im)
    st.seed_candidates(_half(sigma), _half(sigma), 3, 0, 2.0, 1e9)
    want = seed_oracle.gaussian_filter_c(im, sigma)
    assert np.array_equal(st.seed_volume(0), want)
    assert np.array_equal(st.seed_volume(1), want)


@pytest.mark.parametrize("shape", [(5, 7, 9), (30, 33, 70), (3, 130, 257), (61, 40, 8), (50, 24, 72), (50, 130, 33)])
def test_gaussian_ragged_shapes(lib, shape):
    """radius > axis length (repeated reflection), sizes that are not multiples of the tile"""
    rng = np.random.default_rng(3)
    im = rng.integers(0, 65536, size=shape, dtype=np.uint16)
    st = lib.Stack(im)
    st.seed_candidates(_half(0.75), _half(7.5), 3, 0, 2.0, 1e9)
    assert np.array_equal(st.seed_volume(0), seed_oracle.gaussian_filter_c(im, 0.75))
    assert np.array_equal(st.seed_volume(1), seed_oracle.gaussian_filter_c(im, 7.5))


def test_gaussian_extremes(lib):
    """constant / saturated volumes: sums that are exact integers in exact arithmetic sit on the
    truncation boundary, so the FP64 accumulation order decides the result"""
    for val, shape in ((0, (12, 40, 48)), (1, (12, 40, 48)), (300, (12, 40, 48)), (65535, (12, 40, 48)),
                       (300, (50, 16, 24)), (65535, (30, 16, 24)), (0, (50, 16, 24))):   # + the whole-line z kernels
        im = np.full(shape, val, dtype=np.uint16)
        st = lib.Stack(im)
        st.seed_candidates(_half(0.75), _half(7.5), 3, 0, 2.0, 1e9)
        assert np.array_equal(st.seed_volume(0), seed_oracle.gaussian_filter_c(im, 0.75)), val
        assert np.array_equal(st.seed_volume(1), seed_oracle.gaussian_filter_c(im, 7.5)), val


def test_gaussian_float32(lib, golden_seeds):
    imf = golden_seeds["im"].astype(np.float32) / np.float32(301.7)
    st = lib.Stack(imf)
    st.seed_candidates(_half(0.75), _half(7.5), 3, 0, 2.0, 1e9)
    assert np.array_equal(st.seed_volume(0), seed_oracle.gaussian_filter_c(imf, 0.75))
    assert np.array_equal(st.seed_volume(1), seed_oracle.gaussian_filter_c(imf, 7.5))


def test_candidates_are_np_where_ordered(lib, golden_seeds):
    im = golden_seeds["im"]
    st = lib.Stack(im)
    zxy, h, _ = st.seed_candidates(_half(0.75), _half(7.5), 3, 0, 2.0, 30.0)
    fg, bg, mask, diff = seed_oracle.seed_maps(im)
    pts = np.array(np.where(mask & (diff >= 30.0))).T
    ok = ((pts >= 2) & (pts <= np.array(im.shape) - 2)).all(1)
    pts = pts[ok]
    assert np.array_equal(zxy, pts)
    assert np.array_equal(h, diff[pts[:, 0], pts[:, 1], pts[:, 2]])


@pytest.mark.parametrize("name", sorted(SEED_CASES))
def test_get_seeds_matches_reference_fixture(lib, golden_seeds, name):
    from imageanalysis3_b200.spot_tools.fitting import get_seeds
    got = get_seeds(golden_seeds["im"], **SEED_CASES[name])
    want = golden_seeds["seeds_" + name]
    assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want)


def test_get_seeds_float32_and_uint8(lib, golden_seeds):
    from imageanalysis3_b200.spot_tools.fitting import get_seeds
    imf = golden_seeds["im"].astype(np.float32) / np.float32(301.7)
    assert np.array_equal(get_seeds(imf, th_seed=1.0), golden_seeds["seeds_f32"])
    im8 = (golden_seeds["im"] >> 5).astype(np.uint8)
    assert np.array_equal(get_seeds(im8, th_seed=8), seed_oracle.get_seeds_oracle(im8, th_seed=8))


def test_get_seeds_empty_and_errors(lib):
    from imageanalysis3_b200.spot_tools.fitting import get_seeds
    flat = np.full((10, 32, 32), 300, dtype=np.uint16)
    out = get_seeds(flat, th_seed=100)
    assert out.shape == (0, 3)
    with pytest.raises(TypeError):
        get_seeds("not an array")
    with pytest.raises(IndexError):
        get_seeds(flat, sel_center=[1, 2])


def test_get_seeds_random_medium(lib):
    """seeded synthetic stack, several kwargs, oracle computed on the fly"""
    from imageanalysis3_b200.spot_tools.fitting import get_seeds
    from imageanalysis3_b200.synth import synth
    im = synth((30, 160, 176), 160, 7)
    for kw in (dict(th_seed=300), dict(th_seed=300, max_num_seeds=77, return_h=True), dict(th_seed=150, hot_pixel_th=2),
               dict(th_seed=90000, min_dynamic_seeds=40, dynamic_niters=7)):
        want = seed_oracle.get_seeds_oracle(im, backend="c", **kw)
        got = get_seeds(im, **kw)
        assert got.shape == want.shape and np.array_equal(got, want), kw


def test_legacy_seeders_match_reference_fixture(lib, golden_seeds):
    from imageanalysis3_b200 import visual_tools as vt
    im = golden_seeds["im"]
    a = vt.get_seed_in_distance(im, center=None, th_seed=300, return_h=True)
    assert a.dtype == np.int64 and np.array_equal(a, golden_seeds["legacy_all_h"])
    b = vt.get_seed_in_distance(im, center=[10, 40, 50], th_seed=3000, num_seeds=6)
    assert np.array_equal(b, golden_seeds["legacy_center"])
    c = vt.get_seed_points_base(im, th_seed=200, hot_pix_th=3, return_h=True)
    assert np.array_equal(c, golden_seeds["legacy_base"])


def test_full_size_stack_properties(lib):
    """BASELINE config C2 size (50 x 2048 x 2048): size-independent checks.
    (i) a z-slab of the device blur equals the C oracle on a sub-volume deep inside x/y (the filter
    is separable: interior columns only see their own 61-wide neighbourhood);
    (ii) candidates are sorted in C order, inside the edge margin, and idempotent across runs."""
    import torch
    from imageanalysis3_b200.synth import synth_torch
    shape = (50, 2048, 2048)
    d = synth_torch(shape, 5000, 1, "cuda")
    im = d.cpu().numpy().view(np.uint16)
    del d
    torch.cuda.empty_cache()
    st = lib.Stack(im)
    zxy, h, t = st.seed_candidates(_half(0.75), _half(7.5), 3, 0, 2.0, 30.0)
    lin = (zxy[:, 0].astype(np.int64) * shape[1] + zxy[:, 1]) * shape[2] + zxy[:, 2]
    assert len(zxy) > 1000 and np.all(np.diff(lin) > 0)
    assert zxy.min() >= 2 and np.all(zxy <= np.array(shape) - 2)
    zxy2, h2, _ = st.seed_candidates(_half(0.75), _half(7.5), 3, 0, 2.0, 30.0)
    assert np.array_equal(zxy, zxy2) and np.array_equal(h, h2)
    # sub-volume parity: x in [900, 1100), y in [500, 740): compare the interior (away from the
    # crop's own reflecting borders by the background radius 30 + 1)
    sub = np.ascontiguousarray(im[:, 900:1100, 500:740])
    fg = st.seed_volume(0)[:, 900:1100, 500:740]
    bg = st.seed_volume(1)[:, 900:1100, 500:740]
    assert np.array_equal(fg[:, 4:-4, 4:-4], seed_oracle.gaussian_filter_c(sub, 0.75)[:, 4:-4, 4:-4])
    assert np.array_equal(bg[:, 31:-31, 31:-31], seed_oracle.gaussian_filter_c(sub, 7.5)[:, 31:-31, 31:-31])


def test_fitting_v3_seeder_matches_reference_fixture(lib, golden_seeds):
    """External/Fitting_v3.py:261-306 get_seed_points_base (a11): float input and the uint16 wrap-around"""
    from imageanalysis3_b200.External import Fitting_v3
    g = golden_seeds
    im = g["im"]
    imf = im.astype(np.float32) / np.float32(301.7)
    for name, arr, kw in (("v3base_f32", imf, dict(th_seed=0.4, hot_pix_th=3)),
                          ("v3base_f32_h", imf, dict(th_seed=0.4, return_h=True, max_num=50)),
                          ("v3base_f32_snr", imf, dict(th_seed=1.3, use_snr=True)),
                          ("v3base_u16", im, dict(th_seed=200, max_num=300))):
        got = Fitting_v3.get_seed_points_base(arr, **kw)
        want = g[name]
        assert got.shape == want.shape and got.dtype == want.dtype and np.array_equal(got, want), name
