"""CPU: the restated oracle against (i) the committed fixtures the unmodified reference produced,
(ii) scipy itself for the C restatement, (iii) the live reference when /root/reference exists."""
import numpy as np
import pytest
from scipy import ndimage

from oracle import fit_oracle, ref_loader, seed_oracle
from oracle.make_golden import SEED_CASES


@pytest.mark.parametrize("name", sorted(SEED_CASES))
@pytest.mark.parametrize("backend", ["scipy", "c"])
def test_get_seeds_oracle_matches_golden(golden_seeds, name, backend):
    got = seed_oracle.get_seeds_oracle(golden_seeds["im"], backend=backend, **SEED_CASES[name])
    want = golden_seeds["seeds_" + name]
    assert got.dtype == want.dtype and np.array_equal(got, want)


def test_get_seeds_oracle_float32(golden_seeds):
    imf = golden_seeds["im"].astype(np.float32) / np.float32(301.7)
    for backend in ("scipy", "c"):
        assert np.array_equal(seed_oracle.get_seeds_oracle(imf, th_seed=1.0, backend=backend), golden_seeds["seeds_f32"])


def test_legacy_seed_oracle_matches_golden(golden_seeds):
    im = golden_seeds["im"]
    for backend in ("scipy", "c"):
        a = seed_oracle.legacy_seed_in_distance(im, center=None, th_seed=300, return_h=True, backend=backend)
        assert a.dtype == np.int64 and np.array_equal(a, golden_seeds["legacy_all_h"])
        b = seed_oracle.legacy_seed_in_distance(im, center=[10, 40, 50], th_seed=3000, num_seeds=6, backend=backend)
        assert np.array_equal(b, golden_seeds["legacy_center"])
        c = seed_oracle.legacy_seed_points_base(im, th_seed=200, hot_pix_th=3, return_h=True, backend=backend)
        assert np.array_equal(c, golden_seeds["legacy_base"])


@pytest.mark.parametrize("sigma", [0.75, 7.5, 10.0, 2.3])
def test_c_gaussian_is_scipy_bit_exact(golden_seeds, sigma):
    im = golden_seeds["im"]
    assert np.array_equal(seed_oracle.gaussian_filter_c(im, sigma), ndimage.gaussian_filter(im, sigma))
    imf = im.astype(np.float32) * np.float32(0.37)
    assert np.array_equal(seed_oracle.gaussian_filter_c(imf, sigma), ndimage.gaussian_filter(imf, sigma))
    tiny = np.ascontiguousarray(im[:5, :7, :9])      # radius > length: repeated reflection
    assert np.array_equal(seed_oracle.gaussian_filter_c(tiny, sigma), ndimage.gaussian_filter(tiny, sigma))


@pytest.mark.parametrize("size", [3, 4, 5])
def test_c_rank_filters_are_scipy(golden_seeds, size):
    im = golden_seeds["im"]
    assert np.array_equal(seed_oracle.rank_filter_c(im, size, True), ndimage.maximum_filter(im, size))
    assert np.array_equal(seed_oracle.rank_filter_c(im, size, False), ndimage.minimum_filter(im, size))


def _rows(ps):
    return np.array([np.asarray(r, dtype=np.float64) for r in ps])


def test_iter_fit_oracle_v4_matches_golden(golden_fits):
    g = golden_fits
    o = fit_oracle.iter_fit(g["im"], g["seeds"].T, version=4)
    assert np.array_equal(_rows(o["first_ps"]), g["v4_first"], equal_nan=True)
    assert np.array_equal(_rows(o["ps"]), g["v4_final"], equal_nan=True)
    assert o["n_iter"] == int(g["v4_n_iter"]) and np.array_equal(o["converged"], g["v4_converged"])
    assert np.array_equal(o["dists"], g["v4_dists"])


@pytest.mark.parametrize("ws", [0, 1000])
def test_iter_fit_oracle_v3_matches_golden(golden_fits, ws):
    g = golden_fits
    o = fit_oracle.iter_fit(g["im"], g["seeds"].T, version=3, weight_sigma=ws)
    assert np.array_equal(_rows(o["first_ps"]), g[f"v3_ws{ws}_first"], equal_nan=True)
    assert np.array_equal(_rows(o["ps"]), g[f"v3_ws{ws}_final"], equal_nan=True)
    assert o["n_iter"] == int(g[f"v3_ws{ws}_n_iter"])


def test_iter_fit_oracle_edge_seeds(golden_fits):
    g = golden_fits
    o = fit_oracle.iter_fit(g["im"], g["edge_seeds"].T, version=4)
    got = _rows(o["ps"])
    assert np.array_equal(got, g["edge_final"], equal_nan=True)
    assert np.isnan(got).any(1).sum() == 1          # the seed whose clipped window has < 10 voxels


def test_fit_fov_image_oracle_matches_golden(golden_fits):
    g = golden_fits
    spots, _ = fit_oracle.fit_fov_image_oracle(g["im"], th_seed=300, max_num_seeds=None)
    assert spots.dtype == np.float32 and np.array_equal(spots, g["fov_spots"])
    spots, _ = fit_oracle.fit_fov_image_oracle(g["im"], th_seed=300, max_num_seeds=20)
    assert np.array_equal(spots, g["fov_spots_top20"])


def test_gaussian_fit_oracle_matches_golden(golden_fits):
    g = golden_fits
    X = g["gf_X"]
    r = fit_oracle.gaussian_fit(g["im"][X[0], X[1], X[2]], X, center=None, version=4, delta_center=2.5)
    assert np.array_equal(r["p"], g["gf_p"]) and np.array_equal(r["p_raw"], g["gf_p_raw"])
    assert np.array_equal(r["problem"].gauss(r["p_raw"]), g["gf_rec"])


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_oracle_against_live_reference():
    import warnings
    warnings.simplefilter("ignore")
    from imageanalysis3_b200.synth import synth
    ns = ref_loader.load()
    im = synth((16, 64, 72), 35, 21)
    s_ref = ns.fitting.get_seeds(im, th_seed=250, return_h=True)
    assert np.array_equal(seed_oracle.get_seeds_oracle(im, th_seed=250, return_h=True), s_ref)
    f = ns.Fitting_v4.iter_fit_seed_points(im, s_ref[:, :3].T)
    f.firstfit(); f.repeatfit()
    o = fit_oracle.iter_fit(im, s_ref[:, :3].T, version=4)
    assert np.array_equal(_rows(o["ps"]), _rows(f.ps), equal_nan=True)
    assert np.array_equal(o["im_add"], f.im_add)


def test_well_posed_statistics(golden_fits):
    """the cut-off of fit_oracle.comparable_mask is far from what real spots need: isolated and densely
    overlapping spots have cond(J D^-1) < 30 over the non-zero columns, while the junk seeds of the edge
    case end on 1e6 .. 1e8-conditioned Jacobians or run into maxfev."""
    from imageanalysis3_b200.synth import synth
    g = golden_fits
    o = fit_oracle.iter_fit(g["im"], g["seeds"].T, version=4)
    assert o["cond_max"].max() < 10 and o["nfev_max"].max() < 60
    assert o["well_posed"].all() and o["comparable"].all()
    assert np.array_equal(o["comparable"], g["v4_comparable"])
    e = fit_oracle.iter_fit(g["im"], g["edge_seeds"].T, version=4)
    assert np.array_equal(e["comparable"], g["edge_comparable"])
    junk = ~e["well_posed"] & np.array(e["success"])
    assert junk.sum() >= 3 and ((e["cond_max"][junk] > 1e5) | e["unstable"][junk]).all()
    assert (e["cond_max"] > 1e5).sum() >= 3
    im = synth((24, 96, 96), 220, 31, h_range=(500.0, 3000.0))     # crowded: merged / overlapping spots
    seeds = seed_oracle.get_seeds_oracle(im, th_seed=200, backend="c")
    d = fit_oracle.iter_fit(im, seeds.T, version=4)
    assert len(seeds) > 90 and d["cond_max"].max() < 50
    # 10x denser than BASELINE's dense config: merged blobs with widths on their bounds crawl, and a sixth of
    # the fits stop with a last step above the tolerance (0.8 % on the full C2 stack, 1.1 % at C4 density)
    assert d["well_posed"].mean() > 0.75


def test_comparable_mask_taints_window_overlap_components():
    cen = np.array([[10., 10, 10], [10, 10, 18], [10, 10, 26.5], [10, 40, 40], [2.2, 3, 3]])
    well = np.array([True, False, True, True, True])
    # 0 and 2 overlap the ill-posed seed 1 (|d| = 8 <= 2r-1), 3 and 4 are far away
    assert fit_oracle.comparable_mask(cen, well, 5).tolist() == [False, False, False, True, True]
    well = np.array([False, True, True, True, True])
    # taint travels through the chain 0 -> 1 -> 2
    assert fit_oracle.comparable_mask(cen, well, 5).tolist() == [False, False, False, True, True]
    assert fit_oracle.comparable_mask(cen[:0], well[:0], 5).shape == (0,)


def test_background_normalisation_oracle_matches_golden(golden_fits):
    g = golden_fits
    for tag, kw in (("local", dict(normalize_local=True)), ("global", dict(normalize_background=True)),
                    ("local_bin4", dict(normalize_local=True, background_args=dict(bin_size=4)))):
        sp, _ = fit_oracle.fit_fov_image_oracle(g["im"], th_seed=300, max_num_seeds=None, **kw)
        assert np.array_equal(sp, g[f"fov_spots_norm_{tag}"]), tag
    sp, _ = fit_oracle.fit_fov_image_oracle(g["im_ramp"], th_seed=300, max_num_seeds=None, normalize_local=True)
    assert np.array_equal(sp, g["ramp_spots_norm_local"])
    # the height loop's quirk: a peak found only at the 11th halving still falls back to the median
    flat = np.full((4, 5, 6), 7, dtype=np.uint16)             # single bin at the left edge: never a peak
    assert fit_oracle.image_background(flat) == 7.0


def test_reference_sensitivity_probe_flags_slow_crawls():
    """A spot whose width sits on its bound makes MINPACK crawl (> 100 evaluations on a well-conditioned
    Jacobian); its last accepted step is larger than the parity tolerance, and raising 1 % of its float32
    voxel values by one ulp moves scipy's own answer by more than the tolerance.  The oracle must flag it
    (and only a percent of the seeds of a dense stack)."""
    from imageanalysis3_b200.synth import synth
    im = synth((60, 256, 256), 780, 4, h_range=(400.0, 3000.0))
    seeds = seed_oracle.get_seeds_oracle(im, th_seed=300.0, backend="c")
    o = fit_oracle.iter_fit(im, seeds.T, version=4)
    assert (o["cond_max"] < fit_oracle.COND_WELL_POSED).all()
    assert o["unstable"][583] and o["nfev_max"][583] > 100
    assert 1 <= o["unstable"].sum() <= 0.02 * len(seeds)
    # a crawl takes the seeds whose windows overlap its own with it (comparable_mask): 1 % crawls -> 3 % of the rows
    assert o["comparable"].mean() > 0.95
    assert (~o["comparable"]).sum() > o["unstable"].sum()
    # the probe on its own: seed 583's slowest fit moves by > 1e-4 under the one-ulp perturbation
    sig = np.array([np.asarray(r, float) for r in o["ps"]])[583, 5:8]
    assert (np.abs(sig - 4.0) < 1e-3).any()


def test_parallel_oracle_equals_sequential():
    """oracle/parallel.py (independent groups of seeds over a process pool, used for the full-size GPU parity
    tests) gives exactly the sequential oracle's numbers"""
    from imageanalysis3_b200.synth import synth
    from oracle import parallel
    im = synth((30, 256, 256), 120, 17)
    seeds = seed_oracle.get_seeds_oracle(im, th_seed=300, backend="c")
    a = fit_oracle.iter_fit(im, seeds.T, version=4)
    b = parallel.iter_fit_parallel(im, seeds.T, procs=2)
    assert b["n_components"] > 20
    pa = np.array([np.asarray(r, dtype=np.float64) for r in a["ps"]])
    pb = np.array([np.asarray(r, dtype=np.float64) for r in b["ps"]])
    assert np.array_equal(pa, pb, equal_nan=True)
    assert a["n_iter"] == b["n_iter"]
    for k in ("converged", "comparable", "dists", "nfev_max"):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
