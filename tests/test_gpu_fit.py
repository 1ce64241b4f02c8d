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
    assert cmp_ok.mean() > 0.6       # 8x denser than BASELINE's dense config: a tenth of the blobs crawl and take their
                                     # direct window neighbours out of the tolerance comparison
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


def test_c1_config_v4_against_oracle(lib):
    """BASELINE config C1 (30 x 512 x 512, 500 planted spots): fit_fov_image against the oracle on every row"""
    from imageanalysis3_b200.spot_tools.fitting import fit_fov_image
    from imageanalysis3_b200.synth import synth
    im = synth((30, 512, 512), 500, 0)
    want, seeds = fit_oracle.fit_fov_image_oracle(im, th_seed=300, max_num_seeds=None, seed_backend="c")
    got = fit_fov_image(im, '647', th_seed=300, max_num_seeds=None, verbose=False)
    cmp_ok = fit_oracle.fit_fov_image_oracle.last_comparable
    assert len(want) > 400 and cmp_ok.mean() > 0.97
    assert got.dtype == want.dtype
    assert_spots_close(got, want, "C1 v4", cmp_ok)


def test_c1_config_v3_legacy_path_against_oracle(lib):
    """C1 through the legacy per-cell path: visual_tools seeder + Fitting_v3 with the width prior
    (classes/__init__.py:57-88, 3697-3698)"""
    from imageanalysis3_b200 import visual_tools as vt
    from imageanalysis3_b200.External import Fitting_v3
    from imageanalysis3_b200.synth import synth
    im = synth((30, 512, 512), 500, 0)
    seeds = vt.get_seed_in_distance(im, center=None, th_seed=300)
    want_seeds = seed_oracle.legacy_seed_in_distance(im, center=None, th_seed=300, backend="c")
    assert np.array_equal(seeds, want_seeds) and len(seeds) > 300
    cen = seeds[:, :3].T.astype(np.float64)        # (3, N), as _fit_single_image passes it
    o = fit_oracle.iter_fit(im, cen, version=3, weight_sigma=1000)
    f = Fitting_v3.iter_fit_seed_points(im, cen, 5, 1, 2.5, 10, 0.1, [1.35, 1.9, 1.9], weight_sigma=1000)
    f.firstfit()
    f.repeatfit()
    assert o["comparable"].mean() > 0.97
    assert_spots_close(f.ps, o["ps"], "C1 v3 ws=1000", o["comparable"])
    assert np.array_equal(f.converged[o["comparable"]], o["converged"][o["comparable"]])


def test_background_normalisation_matches_fixture(lib, golden_fits):
    """fit_fov_image(normalize_local / normalize_background): histogram-mode backgrounds on the device
    (spot_tools/fitting.py:240-258, io_tools/load.py:642-686) against the unmodified reference"""
    from imageanalysis3_b200.spot_tools.fitting import fit_fov_image
    g = golden_fits
    for tag, kw in (("local", dict(normalize_local=True)), ("global", dict(normalize_background=True)),
                    ("local_bin4", dict(normalize_local=True, background_args=dict(bin_size=4)))):
        got = fit_fov_image(g["im"], '647', th_seed=300, max_num_seeds=None, verbose=False, **kw)
        want = g[f"fov_spots_norm_{tag}"]
        assert got.dtype == want.dtype
        # the divisor must be identical: compare it exactly through the un-normalised fit of the same call
        plain = fit_fov_image(g["im"], '647', th_seed=300, max_num_seeds=None, verbose=False)
        assert np.array_equal(np.round(plain[:, 0] / got[:, 0], 3), np.round(g["fov_spots"][:, 0] / want[:, 0], 3)), tag
        assert_spots_close(got, want, f"normalised ({tag})", g["fov_spots_comparable"])
    got = fit_fov_image(g["im_ramp"], '647', th_seed=300, max_num_seeds=None, normalize_local=True, verbose=False)
    assert_spots_close(got, g["ramp_spots_norm_local"], "ramp, local background", g["ramp_comparable"])


def test_box_background_against_oracle(lib):
    """histogram / find_peaks / median-fallback semantics of ia3_box_background on awkward boxes"""
    rng = np.random.default_rng(11)
    im = rng.poisson(300, size=(24, 60, 64)).astype(np.uint16)
    im[:, :20, :20] = 5                    # all in the first bin: never a peak -> np.nanmedian
    im[:, 20:40, :20] = rng.integers(0, 65536, size=(24, 20, 20))       # flat histogram: tiny peaks
    im[:, 40:, :20] = np.where(rng.random((24, 20, 20)) < 0.5, 1000, 1010)   # plateau of two equal bins
    im[:, :20, 20:40] = 65533              # beyond the last edge: dropped from the histogram
    st = lib.Stack(im)
    boxes = [[0, 24, 0, 20, 0, 20], [0, 24, 20, 40, 0, 20], [0, 24, 40, 60, 0, 20], [0, 24, 0, 20, 20, 40],
             [0, 24, 0, 60, 0, 64], [3, 4, 10, 11, 30, 31], [0, 21, 39, 60, 43, 64], [2, 23, 0, 21, 0, 21],
             [5, 9, 18, 23, 18, 23]]
    for _ in range(40):
        lo = np.array([rng.integers(0, 23), rng.integers(0, 59), rng.integers(0, 63)])
        hi = np.minimum(lo + rng.integers(1, 22, size=3), im.shape)
        boxes.append([lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]])
    boxes = np.array(boxes, dtype=np.int32)
    for bin_size, max_iter in ((10, 10), (3, 10), (64, 2)):
        got = st.box_background(boxes, 0, 65535, bin_size, max_iter)
        want = np.array([fit_oracle.image_background(im[b[0]:b[1], b[2]:b[3], b[4]:b[5]], bin_size=bin_size, max_iter=max_iter)
                         for b in boxes])
        assert np.array_equal(got, want), (bin_size, np.nonzero(got != want)[0][:5], got[got != want][:5], want[got != want][:5])


def test_c4_density_crop_against_oracle(lib):
    """BASELINE config C4 (dense RNA-FISH: 50 000 spots in 60 x 2048 x 2048) at the same density on a
    60 x 256 x 256 crop: overlapping windows, several dependency levels, Voronoi ties"""
    from imageanalysis3_b200.External import Fitting_v4
    from imageanalysis3_b200.spot_tools.fitting import fit_fov_image, get_seeds
    from imageanalysis3_b200.synth import synth
    im = synth((60, 256, 256), 780, 4, h_range=(400.0, 3000.0))
    want, seeds = fit_oracle.fit_fov_image_oracle(im, th_seed=300, max_num_seeds=None, seed_backend="c")
    assert np.array_equal(get_seeds(im, th_seed=300.0), seeds) and len(seeds) > 600
    got = fit_fov_image(im, '647', th_seed=300, max_num_seeds=None, verbose=False)
    cmp_ok = fit_oracle.fit_fov_image_oracle.last_comparable
    assert cmp_ok.mean() > 0.9
    assert_spots_close(got, want, "C4-density crop", cmp_ok)
    f = Fitting_v4.iter_fit_seed_points(im, seeds.T)
    f.firstfit()
    assert f._h.num_levels >= 2 and f.n_tie_voxels > 0


def test_fast_fit_big_image_matches_fixture(lib, golden_fits):
    """Fitting_v4.fast_fit_big_image / gfit_fast (a13): weighted-moment fits on the device against the
    unmodified reference, incl. the uint16 wrap-around of the weights and the GaussianFit variant"""
    from imageanalysis3_b200.External import Fitting_v4
    g = golden_fits
    im = g["im"]
    imd = im.astype(np.float64)
    cases = {"f64": (imd, g["seeds"], {}, 1e-10), "f64_noavoid_r5": (imd, g["seeds"], dict(avoid_neigbors=False, radius_fit=5), 1e-10),
             "f64_close_recenter": (imd, g["fastfit_close"], dict(recenter=True), 1e-10),
             "f64_jitter": (imd, g["fastfit_jitter"], {}, 1e-10),
             "u16": (im, g["fastfit_close"], {}, 1e-10), "f32": (im.astype(np.float32), g["fastfit_close"], {}, 2e-5)}
    for tag, (arr, cen, kw, tol) in cases.items():
        got = Fitting_v4.fast_fit_big_image(arr, cen, verbose=False, **kw)
        want = g["fastfit_" + tag]
        assert got.shape == want.shape and got.dtype == want.dtype, tag
        assert np.array_equal(np.isnan(got), np.isnan(want)), tag
        ok = ~np.isnan(want)
        scale = np.maximum(np.abs(want[ok]), 1.0)
        assert (np.abs(got[ok] - want[ok]) / scale).max() <= tol, (tag, (np.abs(got[ok] - want[ok]) / scale).max())
    got = Fitting_v4.fast_fit_big_image(imd, g["fastfit_close"][:12], verbose=False, better_fit=True)
    assert got.shape == g["fastfit_better"].shape
    assert_spots_close(got, g["fastfit_better"], "fast_fit_big_image(better_fit=True)")
    assert Fitting_v4.fast_fit_big_image(imd, np.zeros((0, 3)), verbose=False).shape == (0,)


def test_remaining_argument_paths_match_fixture(lib, golden_fits):
    """fit_fov_image with given seeds / a seed mask / a float32 image / another radius without the boundary
    filter, get_centers with a crop and without the duplicate filter -- against the unmodified reference"""
    from imageanalysis3_b200.spot_tools.fitting import fit_fov_image, get_centers
    g = golden_fits
    im, seeds = g["im"], g["seeds"]
    got = fit_fov_image(im, '647', seeds=np.concatenate([seeds[:15], np.ones((15, 1))], axis=1), verbose=False)
    assert_spots_close(got, g["fov_given_seeds"], "given seeds")
    mask = np.zeros(im.shape, dtype=np.uint8)
    mask[:, :40, :] = 1
    got = fit_fov_image(im, '647', th_seed=300, max_num_seeds=None, seed_mask=mask, verbose=False)
    assert_spots_close(got, g["fov_seed_mask"], "seed mask")
    imf32 = im.astype(np.float32) / np.float32(2.5)
    got = fit_fov_image(imf32, '647', th_seed=120, max_num_seeds=None, verbose=False)
    assert got.dtype == g["fov_f32"].dtype
    assert_spots_close(got, g["fov_f32"], "float32 image")
    got = fit_fov_image(im, '647', th_seed=300, max_num_seeds=25, fit_radius=4, remove_boundary_points=False, verbose=False)
    assert_spots_close(got, g["fov_noboundary_r4"], "radius 4, no boundary filter")
    c = get_centers(im, th_seed=300, sel_center=[10, 36, 40], seed_radius=25)
    assert c.shape == g["centers_crop"].shape and np.abs(c - g["centers_crop"]).max() <= 1e-3
    c = get_centers(im, th_seed=300, remove_close_pts=False, max_num_seeds=12)
    assert c.shape == g["centers_noclose"].shape and np.abs(c - g["centers_noclose"]).max() <= 1e-3


_ENGINE_SCRIPT = r'''
import sys
import numpy as np
sys.path.insert(0, ".")
from imageanalysis3_b200.External import Fitting_v3, Fitting_v4
from imageanalysis3_b200.spot_tools import fitting
from imageanalysis3_b200.synth import synth
im = synth((30, 128, 128), 400, 31, h_range=(500.0, 3000.0))       # crowded: many long LM runs, overlapping windows
seeds = fitting.get_seeds(im, max_num_seeds=None, th_seed=200.0)
out = {}
for name, mod in (("v4", Fitting_v4), ("v3", Fitting_v3)):
    f = mod.iter_fit_seed_points(im, seeds.T)
    f.firstfit()
    out[name + "_first"] = np.array(f.ps)
    f.repeatfit()
    out[name + "_ps"] = np.array(f.ps)
    out[name + "_nfev"] = np.array(f.nfev)
    out[name + "_n_iter"] = f.n_iter
    out[name + "_converged"] = f.converged
    out[name + "_dists"] = f.dists
    g = mod.iter_fit_seed_points(im, seeds.T)
    g._fit_all()                                # one device run, with speculation
    out[name + "_all_ps"] = np.array(g.ps)
    out[name + "_all_n_iter"] = g.n_iter
    out[name + "_all_converged"] = g.converged
    st = g._h.engine_stats()
    out[name + "_stats"] = np.array([st["parked"], st["team_tasks"], st["memo_hits"], st["spec_hits"]])
np.savez(sys.argv[1], **out)
'''


def test_engine_schedule_does_not_change_a_bit(lib, tmp_path):
    """The fit engine may suspend a run after any number of evaluations, continue it with one warp or
    with a team of warps, skip a re-fit whose inputs are those of the seed's previous visit (memo) and
    start the first repeat visit of isolated seeds speculatively: none of that may change a single bit.
    Reference configuration: no suspension, no memo, no speculation, no merging of short rounds (every visit
    runs lmder from scratch on one warp).  Aggressive configuration: suspend every 3 evaluations, team continuation after
    9, memo and speculation on."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfgs = {"plain": dict(IA3_FIT_CAP="0", IA3_FIT_MEMO="0", IA3_FIT_SPEC="0", IA3_FIT_MERGE="0"),
            "busy": dict(IA3_FIT_CAP="3", IA3_FIT_TEAM_AFTER="9", IA3_FIT_TEAM_CAP="5", IA3_FIT_TEAM_CAP_LONG="7", IA3_FIT_MEMO="1",
                         IA3_FIT_SPEC="1", IA3_FIT_CHUNK="3"),
            "merged": dict(IA3_FIT_MERGE="100000", IA3_FIT_CAP="3", IA3_FIT_TEAM_CAP="4"),   # every task of every round runs on a team
            "merge_small": dict(IA3_FIT_MERGE="64"),
            "default": {}}
    res = {}
    for tag, env in cfgs.items():
        path = str(tmp_path / f"{tag}.npz")
        subprocess.run([sys.executable, "-c", _ENGINE_SCRIPT, path], cwd=root, check=True, timeout=600, env={**os.environ, **env})
        res[tag] = np.load(path)
    assert res["plain"]["v4_nfev"].max() > 50                       # the image does contain long runs
    assert res["plain"]["v4_stats"][:3].sum() == 0
    assert (res["busy"]["v4_stats"] > 0).all(), res["busy"]["v4_stats"]     # every mechanism was exercised
    for tag in ("busy", "merged", "merge_small", "default"):
        for key in res["plain"].files:
            if key.endswith("_stats"):
                continue
            a, b = res["plain"][key], res[tag][key]
            assert a.shape == b.shape and a.dtype == b.dtype, (tag, key)
            assert np.array_equal(a, b, equal_nan=True), (tag, key)
    for name in ("v4", "v3"):                                        # one run == firstfit(); repeatfit()
        r = res["plain"]
        assert np.array_equal(r[name + "_ps"], r[name + "_all_ps"], equal_nan=True)
        assert r[name + "_n_iter"] == r[name + "_all_n_iter"] and np.array_equal(r[name + "_converged"], r[name + "_all_converged"])


def test_device_convergence_rule_equals_host_loop(lib, golden_fits):
    """repeatfit's per-seed rule evaluated on the device (ia3_fit_run) against the reference's loop run on the
    host with one device sweep per iteration (ia3_fit_repeat_sweep): same rows, n_iter, converged, dists and
    the *_old attributes, also with a seed that never gets a fit (float64 distance arithmetic)."""
    from imageanalysis3_b200.External import Fitting_v4
    g = golden_fits
    for seeds in (g["seeds"], g["edge_seeds"]):
        a = Fitting_v4.iter_fit_seed_points(g["im"], seeds.T)
        a.firstfit(); a.repeatfit()
        b = Fitting_v4.iter_fit_seed_points(g["im"], seeds.T)
        b.firstfit(); b._repeatfit_host_loop()
        c = Fitting_v4.iter_fit_seed_points(g["im"], seeds.T)
        c._fit_all()
        for other in (b, c):
            assert np.array_equal(a._ps, other._ps, equal_nan=True)
            assert a.n_iter == other.n_iter and np.array_equal(a.converged, other.converged) and np.array_equal(a.dists, other.dists)
            assert np.array_equal(a.success_old, other.success_old)
            assert a.centers_fit_old.dtype == other.centers_fit_old.dtype
            assert np.array_equal(a.centers_fit_old, other.centers_fit_old, equal_nan=True)
            assert np.array(a.ps).dtype == np.array(other.ps).dtype
