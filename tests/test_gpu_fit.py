"""GPU parity, fit stage: CUDA (through the C ABI) against the committed fixtures produced by the
unmodified reference and against the oracle on seeded inputs.  Tolerances are BASELINE.json's:
|d centre| <= 1e-3 px, sigma / height relative error <= 1e-4, identical accept/reject."""
import numpy as np
import pytest

from conftest import assert_spots_close
from oracle import fit_oracle, seed_oracle

pytestmark = pytest.mark.gpu


def test_v4_firstfit_and_repeatfit_match_fixture(lib, golden_fits):
    from imageanalysis3_b200.External import Fitting_v4
    g = golden_fits
    f = Fitting_v4.iter_fit_seed_points(g["im"], g["seeds"].T)
    f.firstfit()
    assert g["v4_comparable"].all()
    assert_spots_close(f.ps, g["v4_first"], "v4 firstfit")
    assert all(f.success)
    f.repeatfit()
    assert_spots_close(f.ps, g["v4_final"], "v4 repeatfit")
    assert f.n_iter == int(g["v4_n_iter"])
    assert np.array_equal(f.converged, g["v4_converged"])
    assert np.array(f.ps).dtype == np.float32


@pytest.mark.parametrize("ws", [0, 1000])
def test_v3_matches_fixture(lib, golden_fits, ws):
    from imageanalysis3_b200.External import Fitting_v3
    g = golden_fits
    f = Fitting_v3.iter_fit_seed_points(g["im"], g["seeds"].T, weight_sigma=ws)
    f.firstfit()
    assert g[f"v3_ws{ws}_comparable"].all()
    assert_spots_close(f.ps, g[f"v3_ws{ws}_first"], f"v3 ws={ws} firstfit")
    f.repeatfit()
    assert_spots_close(f.ps, g[f"v3_ws{ws}_final"], f"v3 ws={ws} repeatfit")
    assert f.n_iter == int(g[f"v3_ws{ws}_n_iter"])
    assert np.array_equal(f.converged, g[f"v3_ws{ws}_converged"])


def test_edge_seeds_and_failed_fit_dtype(lib, golden_fits):
    """junk seeds on and beyond the image border next to real ones: a seed with < 10 voxels gives a
    NaN row and makes np.array(ps) float64 (SURVEY App. D dtype trap); corner windows on pure noise
    are ill-posed in the reference itself (rank-deficient Jacobian, maxfev hit) and are only checked
    for accept/reject; the real spots beside them must stay in tolerance."""
    from imageanalysis3_b200.External import Fitting_v4
    g = golden_fits
    f = Fitting_v4.iter_fit_seed_points(g["im"], g["edge_seeds"].T)
    f.firstfit()
    f.repeatfit()
    assert g["edge_comparable"].sum() >= 6 and not g["edge_comparable"].all()
    assert_spots_close(f.ps, g["edge_final"], "edge seeds", g["edge_comparable"])
    assert np.array(f.ps).dtype == np.float64
    assert f.success.count(False) == 1


def test_fit_fov_image_matches_fixture(lib, golden_fits):
    from imageanalysis3_b200.spot_tools.fitting import fit_fov_image, get_centers
    g = golden_fits
    spots = fit_fov_image(g["im"], '647', th_seed=300, max_num_seeds=None, verbose=False)
    assert spots.dtype == np.float32
    assert_spots_close(spots, g["fov_spots"], "fit_fov_image", g["fov_spots_comparable"])
    spots = fit_fov_image(g["im"], '647', th_seed=300, max_num_seeds=20, verbose=False)
    assert_spots_close(spots, g["fov_spots_top20"], "fit_fov_image top20", g["fov_spots_top20_comparable"])
    c = get_centers(g["im"], th_seed=300)
    assert c.shape == g["centers"].shape and np.abs(c - g["centers"]).max() <= 1e-3
    assert fit_fov_image(np.full((10, 32, 32), 300, np.uint16), '647', verbose=False).shape == (0,)


def test_legacy_bead_centres_match_fixture(lib, golden_fits):
    from imageanalysis3_b200 import visual_tools as vt
    g = golden_fits
    c = vt.get_STD_centers(g["im"], th_seed=300)
    assert c.shape == g["std_centers"].shape and np.abs(c - g["std_centers"]).max() <= 1e-3


def test_standalone_gaussianfit(lib, golden_fits):
    from imageanalysis3_b200.External import Fitting_v4
    g = golden_fits
    X = g["gf_X"]
    obj = Fitting_v4.GaussianFit(g["im"][X[0], X[1], X[2]], X, center=None, delta_center=2.5)
    obj.fit()
    assert obj.success
    assert_spots_close([obj.p], [g["gf_p"]], "GaussianFit")
    rec = obj.get_im()
    assert np.allclose(rec, g["gf_rec"], rtol=1e-4, atol=1e-3)
    few = Fitting_v4.GaussianFit(np.arange(5.), np.zeros((3, 5)), center=[0, 0, 0])
    few.fit()
    assert few.success is False


def test_zero_seed_errors_mirror_reference(lib, golden_fits):
    from imageanalysis3_b200.External import Fitting_v3, Fitting_v4
    im = golden_fits["im"]
    with pytest.raises(ValueError):
        Fitting_v3.iter_fit_seed_points(im, np.zeros((3, 0))).firstfit()
    with pytest.raises(AttributeError):
        Fitting_v4.iter_fit_seed_points(im, np.zeros((3, 0))).firstfit()


def test_dense_overlapping_spots_keep_sequential_semantics(lib):
    """dense stack: many overlapping windows, Voronoi ties, several dependency levels -- the level
    schedule must reproduce the reference's in-order Gauss-Seidel sweeps."""
    from imageanalysis3_b200.External import Fitting_v4
    from imageanalysis3_b200.synth import synth
    im = synth((30, 128, 128), 400, 31, h_range=(500.0, 3000.0))
    seeds = seed_oracle.get_seeds_oracle(im, th_seed=200, backend="c")
    assert len(seeds) > 200
    o = fit_oracle.iter_fit(im, seeds.T, version=4)
    f = Fitting_v4.iter_fit_seed_points(im, seeds.T)
    f.firstfit()
    assert f._h.num_levels >= 3
    cmp_ok = o["comparable"]
    assert cmp_ok.mean() > 0.9
    assert_spots_close(f.ps, o["first_ps"], "dense firstfit", cmp_ok)
    if cmp_ok.all():
        assert np.allclose(f.im_subtr, o["im_subtr"], rtol=0, atol=1e-6)
    f.repeatfit()
    assert np.array_equal(f.converged[cmp_ok], o["converged"][cmp_ok])
    assert_spots_close(f.ps, o["ps"], "dense repeatfit", cmp_ok)
    if cmp_ok.all():
        assert f.n_iter == o["n_iter"]
        assert np.allclose(f.im_add, o["im_add"], rtol=0, atol=1e-4)


def test_float_seeds_and_duplicates(lib, golden_fits):
    """user-supplied non-integer seeds (int() picks the window, the float is the sigmoid centre)
    and an exact duplicate seed"""
    from imageanalysis3_b200.External import Fitting_v3, Fitting_v4
    g = golden_fits
    rng = np.random.default_rng(5)
    seeds = np.concatenate([g["seeds"][:12] + rng.uniform(-0.45, 0.45, size=(12, 3)), g["seeds"][:1]])
    for mod, ver in ((Fitting_v4, 4), (Fitting_v3, 3)):
        o = fit_oracle.iter_fit(g["im"], seeds.T, version=ver)
        f = mod.iter_fit_seed_points(g["im"], seeds.T)
        f.firstfit()
        assert o["comparable"].sum() >= 10
        assert_spots_close(f.ps, o["first_ps"], f"v{ver} float seeds first", o["comparable"])
        f.repeatfit()
        assert_spots_close(f.ps, o["ps"], f"v{ver} float seeds final", o["comparable"])


def test_medium_fov_against_oracle(lib):
    from imageanalysis3_b200.spot_tools.fitting import fit_fov_image
    from imageanalysis3_b200.synth import synth
    im = synth((30, 128, 128), 110, 17)
    want, seeds = fit_oracle.fit_fov_image_oracle(im, th_seed=300, max_num_seeds=None, seed_backend="c")
    got = fit_fov_image(im, '647', th_seed=300, max_num_seeds=None, verbose=False)
    cmp_ok = fit_oracle.fit_fov_image_oracle.last_comparable
    assert cmp_ok.mean() > 0.95
    assert_spots_close(got, want, "medium fov", cmp_ok)


def test_concurrent_stacks_match_sequential(lib):
    """several stacks in flight (host threads, one CUDA stream per stack) give exactly the results
    of running them one after the other"""
    from imageanalysis3_b200 import sharding
    from imageanalysis3_b200.spot_tools.fitting import fit_fov_image
    from imageanalysis3_b200.synth import synth
    ims = [synth((20, 96, 104), 40, 50 + i) for i in range(6)]
    run = lambda im: fit_fov_image(im, '647', th_seed=300, max_num_seeds=None, verbose=False)
    seq = [run(im) for im in ims]
    par = sharding.map_stacks(run, ims * 3, inflight=6)
    for i, got in enumerate(par):
        want = seq[i % len(ims)]
        assert got.shape == want.shape and np.array_equal(got, want, equal_nan=True)


def test_trim_releases_and_guards(lib, golden_fits):
    from imageanalysis3_b200.External import Fitting_v4
    g = golden_fits
    st = lib.Stack(g["im"])
    from imageanalysis3_b200.spot_tools.fitting import get_seeds
    seeds = get_seeds(g["im"], th_seed=300, _stack=st)
    st.trim(1)                                   # seed buffers gone, image still there
    with pytest.raises(lib.IA3Error):
        st.seed_volume(0)
    f = Fitting_v4.iter_fit_seed_points(g["im"], seeds.T, _stack=st)
    f.firstfit()
    st.trim(2)                                   # the library's image copy gone: repeatfit still works
    f.repeatfit()
    assert_spots_close(f.ps, g["v4_final"], "after trim")
    with pytest.raises(lib.IA3Error):
        f.im_subtr
    with pytest.raises(lib.IA3Error):
        get_seeds(g["im"], th_seed=300, _stack=st)
