"""Generates tests/golden/*.npz by running the UNMODIFIED reference (under the import shims of
oracle/ref_loader.py) in the build container.  Test infrastructure; run from the repo root:

    python -m oracle.make_golden

The fixtures carry the input stack itself, so they do not depend on the synthetic generator.
"""
import os
import sys
import warnings

import numpy as np
import scipy

from imageanalysis3_b200.synth import synth

from . import fit_oracle, ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SEED_CASES = {
    "default": dict(th_seed=300),
    "top40_h": dict(th_seed=300, max_num_seeds=40, return_h=True),
    "dynamic": dict(th_seed=6000, min_dynamic_seeds=25),
    "nodyn": dict(th_seed=500, use_dynamic_th=False),
    "crop": dict(th_seed=200, sel_center=[10, 40, 50], seed_radius=30),
    "nohot_noedge": dict(th_seed=300, remove_hot_pixel=False, min_edge_distance=0),
    "filt5": dict(th_seed=300, filt_size=5, min_edge_distance=4),
    "sigma1_bg5": dict(th_seed=250, gfilt_size=1.0, background_gfilt_size=5.0),
}


def _rows(ps):
    return np.array([np.asarray(r, dtype=np.float64) for r in ps])


def main():
    warnings.simplefilter("ignore")
    ns = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    meta = dict(numpy=np.__version__, scipy=scipy.__version__, python=sys.version.split()[0])

    # ---- seeds -------------------------------------------------------------------------------
    im = synth((20, 80, 96), 60, 11)
    im[:, 30, 40] = 5000          # a hot column (same x,y in every z)
    out = dict(im=im, meta=np.array(repr(meta)))
    for name, kw in SEED_CASES.items():
        out["seeds_" + name] = ns.fitting.get_seeds(im, **kw)
    out["legacy_all_h"] = ns.visual.get_seed_in_distance(im, center=None, th_seed=300, return_h=True)
    out["legacy_center"] = ns.visual.get_seed_in_distance(im, center=[10, 40, 50], th_seed=3000, num_seeds=6)
    out["legacy_base"] = ns.visual.get_seed_points_base(im, th_seed=200, hot_pix_th=3, return_h=True)
    imf = (im.astype(np.float32) / np.float32(301.7))
    out["seeds_f32"] = ns.fitting.get_seeds(imf, th_seed=1.0)
    # Fitting_v3's own seeder (a11): float input (as intended) and uint16 input (wrap-around heights)
    out["v3base_f32"] = ns.Fitting_v3.get_seed_points_base(imf, th_seed=0.4, hot_pix_th=3)
    out["v3base_f32_h"] = ns.Fitting_v3.get_seed_points_base(imf, th_seed=0.4, return_h=True, max_num=50)
    out["v3base_f32_snr"] = ns.Fitting_v3.get_seed_points_base(imf, th_seed=1.3, use_snr=True)
    out["v3base_u16"] = ns.Fitting_v3.get_seed_points_base(im, th_seed=200, max_num=300)
    np.savez_compressed(os.path.join(OUT, "seeds_small.npz"), **out)

    # ---- fits --------------------------------------------------------------------------------
    im = synth((20, 72, 80), 45, 12)
    seeds = ns.fitting.get_seeds(im, th_seed=300)
    out = dict(im=im, seeds=seeds, meta=np.array(repr(meta)))
    f = ns.Fitting_v4.iter_fit_seed_points(im, seeds.T)
    f.firstfit()
    out["v4_first"] = _rows(f.ps)
    f.repeatfit()
    out.update(v4_final=_rows(f.ps), v4_converged=f.converged, v4_n_iter=f.n_iter, v4_dists=f.dists,
               v4_success=np.array(f.success))
    # which rows are determined by their data (fit_oracle.comparable_mask), from the oracle after
    # checking that it reproduces the reference bit for bit
    o = fit_oracle.iter_fit(im, seeds.T, version=4)
    assert np.array_equal(_rows(o["ps"]), out["v4_final"], equal_nan=True)
    out["v4_comparable"] = o["comparable"]
    for ws in (0, 1000):
        f = ns.Fitting_v3.iter_fit_seed_points(im, seeds.T, weight_sigma=ws)
        f.firstfit()
        out[f"v3_ws{ws}_first"] = _rows(f.ps)
        f.repeatfit()
        out[f"v3_ws{ws}_final"] = _rows(f.ps)
        out[f"v3_ws{ws}_converged"] = f.converged
        out[f"v3_ws{ws}_n_iter"] = f.n_iter
        o = fit_oracle.iter_fit(im, seeds.T, version=3, weight_sigma=ws)
        assert np.array_equal(_rows(o["ps"]), out[f"v3_ws{ws}_final"], equal_nan=True)
        out[f"v3_ws{ws}_comparable"] = o["comparable"]
    out["fov_spots"] = ns.fitting.fit_fov_image(im, '647', th_seed=300, max_num_seeds=None, verbose=False)
    sp, _ = fit_oracle.fit_fov_image_oracle(im, th_seed=300, max_num_seeds=None)
    assert np.array_equal(sp, out["fov_spots"])
    out["fov_spots_comparable"] = fit_oracle.fit_fov_image_oracle.last_comparable
    out["fov_spots_top20"] = ns.fitting.fit_fov_image(im, '647', th_seed=300, max_num_seeds=20, verbose=False)
    sp, _ = fit_oracle.fit_fov_image_oracle(im, th_seed=300, max_num_seeds=20)
    assert np.array_equal(sp, out["fov_spots_top20"])
    out["fov_spots_top20_comparable"] = fit_oracle.fit_fov_image_oracle.last_comparable
    # intensity normalisation (find_image_background per spot crop / on the whole stack); a second stack
    # with a background ramp so that the local backgrounds differ from spot to spot
    for tag, kw in (("local", dict(normalize_local=True)), ("global", dict(normalize_background=True)),
                    ("local_bin4", dict(normalize_local=True, background_args=dict(bin_size=4)))):
        out[f"fov_spots_norm_{tag}"] = ns.fitting.fit_fov_image(im, '647', th_seed=300, max_num_seeds=None, verbose=False, **kw)
        sp, _ = fit_oracle.fit_fov_image_oracle(im, th_seed=300, max_num_seeds=None, **kw)
        assert np.array_equal(sp, out[f"fov_spots_norm_{tag}"])
    ramp = (im.astype(np.int64) + (np.arange(im.shape[2])[None, None, :] * 9) + (np.arange(im.shape[1])[None, :, None] * 5)).astype(np.uint16)
    out["im_ramp"] = ramp
    out["ramp_spots_norm_local"] = ns.fitting.fit_fov_image(ramp, '647', th_seed=300, max_num_seeds=None, normalize_local=True, verbose=False)
    sp, _ = fit_oracle.fit_fov_image_oracle(ramp, th_seed=300, max_num_seeds=None, normalize_local=True)
    assert np.array_equal(sp, out["ramp_spots_norm_local"])
    out["ramp_comparable"] = fit_oracle.fit_fov_image_oracle.last_comparable
    # the remaining argument paths of fit_fov_image / get_centers: given seeds, seed mask, float32 image, crop
    mask = np.zeros(im.shape, dtype=np.uint8)
    mask[:, :40, :] = 1
    out["fov_given_seeds"] = ns.fitting.fit_fov_image(im, '647', seeds=np.concatenate([seeds[:15], np.ones((15, 1))], axis=1), verbose=False)
    out["fov_seed_mask"] = ns.fitting.fit_fov_image(im, '647', th_seed=300, max_num_seeds=None, seed_mask=mask, verbose=False)
    imf32 = im.astype(np.float32) / np.float32(2.5)
    out["fov_f32"] = ns.fitting.fit_fov_image(imf32, '647', th_seed=120, max_num_seeds=None, verbose=False)
    out["fov_noboundary_r4"] = ns.fitting.fit_fov_image(im, '647', th_seed=300, max_num_seeds=25, fit_radius=4,
                                                        remove_boundary_points=False, verbose=False)
    out["centers_crop"] = ns.fitting.get_centers(im, th_seed=300, sel_center=[10, 36, 40], seed_radius=25)
    out["centers_noclose"] = ns.fitting.get_centers(im, th_seed=300, remove_close_pts=False, max_num_seeds=12)
    out["centers"] = ns.fitting.get_centers(im, th_seed=300)
    out["std_centers"] = ns.visual.get_STD_centers(im, th_seed=300)
    # seeds at the border: windows clipped by the image, one seed with < 10 voxels -> NaN row
    edge_seeds = np.array([[0., 0., 0.], [19., 71., 79.], [10., 0., 40.], [3., 36., 79.], [-2., -2., -3.],
                           [-2., -2., -2.], [8.6, 30.3, 41.7]])
    f = ns.Fitting_v4.iter_fit_seed_points(im, np.concatenate([seeds[:6], edge_seeds]).T)
    f.firstfit(); f.repeatfit()
    out["edge_seeds"] = np.concatenate([seeds[:6], edge_seeds])
    out["edge_final"] = _rows(f.ps)
    o = fit_oracle.iter_fit(im, out["edge_seeds"].T, version=4)
    assert np.array_equal(_rows(o["ps"]), out["edge_final"], equal_nan=True)
    out["edge_comparable"] = o["comparable"]
    out["edge_cond_max"], out["edge_nfev_max"] = o["cond_max"], o["nfev_max"]
    # Fitting_v4's moment fit (a13): float64 / uint16 / float32 images, neighbours avoided or not, recentring
    ff = ns.Fitting_v4.fast_fit_big_image
    imd = im.astype(np.float64)
    jit = seeds + np.random.default_rng(3).uniform(-0.4, 0.4, size=seeds.shape)
    close = np.concatenate([seeds, seeds[:5] + [1, 2, -2], edge_seeds[:4]])
    for tag, (arr, cen, kw) in {"f64": (imd, seeds, {}), "f64_noavoid_r5": (imd, seeds, dict(avoid_neigbors=False, radius_fit=5)),
                                "f64_close_recenter": (imd, close, dict(recenter=True)), "f64_jitter": (imd, jit, {}),
                                "u16": (im, close, {}), "f32": (im.astype(np.float32), close, {})}.items():
        out["fastfit_" + tag] = ff(arr, cen, verbose=False, **kw)
        o2 = fit_oracle.fast_fit_big_image_oracle(arr, cen, **kw)
        assert np.array_equal(o2, out["fastfit_" + tag], equal_nan=True), tag
    out["fastfit_close"], out["fastfit_jitter"] = close, jit
    out["fastfit_better"] = ff(imd, close[:12], verbose=False, better_fit=True)
    o2 = fit_oracle.fast_fit_big_image_oracle(imd, close[:12], better_fit=True)
    assert np.array_equal(o2, out["fastfit_better"], equal_nan=True)
    # single GaussianFit problems
    zb, xb, yb = f.zb, f.xb, f.yb
    c = seeds[0]
    X = np.array([int(c[0]) + zb, int(c[1]) + xb, int(c[2]) + yb])
    ok = ((X >= 0) & (X < np.array(im.shape)[:, None])).all(0)
    X = X[:, ok]
    g = ns.Fitting_v4.GaussianFit(im[X[0], X[1], X[2]], X, center=None, delta_center=2.5)
    g.fit()
    out.update(gf_X=X, gf_p=g.p, gf_p_raw=g.p_, gf_rec=g.get_im())
    np.savez_compressed(os.path.join(OUT, "fits_small.npz"), **out)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


def main_extra():
    """tests/golden/extra_r2.npz: argument paths and functions that had no fixture in round 1 -- percentile
    thresholds, find_matched_seeds, standalone Fitting_v3.GaussianFit, other input dtypes / 2-D images, and the
    alternative seeders of Fitting_v4 (get_seed_points_base, get_seed_points_base_v2, fft_gaussian_fast)."""
    warnings.simplefilter("ignore")
    ns = ref_loader.load()
    meta = dict(numpy=np.__version__, scipy=scipy.__version__, python=sys.version.split()[0])
    try:
        import cv2
        meta["cv2"] = cv2.__version__
    except ImportError:
        cv2 = None
    im = synth((20, 80, 96), 60, 11)
    im[:, 30, 40] = 5000
    out = dict(im=im, meta=np.array(repr(meta)))
    # percentile thresholds (spot_tools/fitting.py:75-76, visual_tools.py:1808-1811)
    out["seeds_percentile95"] = ns.fitting.get_seeds(im, use_percentile=True, th_seed_per=95)
    out["seeds_percentile99_h"] = ns.fitting.get_seeds(im, use_percentile=True, th_seed_per=99.5, return_h=True)
    out["fov_percentile"] = ns.fitting.fit_fov_image(im, '647', use_percentile=True, th_seed_per=99.5, max_num_seeds=30, verbose=False)
    sp, _ = fit_oracle.fit_fov_image_oracle(im, max_num_seeds=30, use_percentile=True, th_seed_per=99.5)
    assert np.array_equal(sp, out["fov_percentile"])
    out["fov_percentile_comparable"] = fit_oracle.fit_fov_image_oracle.last_comparable
    out["legacy_by_per"] = ns.visual.get_seed_in_distance(im, center=None, seed_by_per=True, th_seed_percentile=99.5, return_h=True)
    out["legacy_by_per_center"] = ns.visual.get_seed_in_distance(im, center=[10, 40, 50], seed_by_per=True, th_seed_percentile=99.5,
                                                                 num_seeds=5, return_h=True)
    # find_matched_seeds (visual_tools.py:3081-3139)
    ref = ns.visual.get_seed_in_distance(im, center=None, th_seed=300)[:25].astype(np.float64)
    ref = ref + np.random.default_rng(7).uniform(-1.2, 1.2, size=ref.shape)
    ref = np.concatenate([ref, [[3.0, 5.0, 5.0], [10.0, 30.0, 41.5]]])
    out["matched_ref"] = ref
    for tag, kw in (("default", {}), ("unique_d5", dict(keep_unique=True, search_distance=5)), ("th600", dict(th_seed=600, search_distance=2))):
        m, f = ns.visual.find_matched_seeds(im, ref, verbose=False, **kw)
        out[f"matched_{tag}"], out[f"matched_{tag}_found"] = m, f
    # other input types of get_seeds: the filters keep the input's dtype (spot_tools/fitting.py:91-106)
    out["seeds_i32"] = ns.fitting.get_seeds(im.astype(np.int32), th_seed=300)
    out["seeds_i64_h"] = ns.fitting.get_seeds(im.astype(np.int64), th_seed=300, return_h=True, max_num_seeds=30)
    imf64 = im.astype(np.float64) / 2.5
    out["seeds_f64"] = ns.fitting.get_seeds(imf64, th_seed=120)
    out["fov_f64"] = ns.fitting.fit_fov_image(imf64, '647', th_seed=120, max_num_seeds=20, verbose=False)
    out["seeds_2d"] = ns.fitting.get_seeds(im[10], th_seed=300, remove_hot_pixel=False)
    out["seeds_2d_f32_h"] = ns.fitting.get_seeds(im[10].astype(np.float32), th_seed=300, remove_hot_pixel=False, return_h=True)
    # standalone Fitting_v3.GaussianFit (External/Fitting_v3.py:50-257)
    seeds = ns.fitting.get_seeds(im, th_seed=300)
    f = ns.Fitting_v3.iter_fit_seed_points(im, seeds.T)
    c = seeds[0]
    X = np.array([int(c[0]) + f.zb, int(c[1]) + f.xb, int(c[2]) + f.yb])
    X = X[:, ((X >= 0) & (X < np.array(im.shape)[:, None])).all(0)]
    out["gf3_X"] = X
    for ws in (0, 1000):
        g = ns.Fitting_v3.GaussianFit(im[X[0], X[1], X[2]], X, center=list(c), delta_center=2.5, weight_sigma=ws)
        g.fit()
        out[f"gf3_ws{ws}_p"], out[f"gf3_ws{ws}_rec"] = g.p, g.get_im()
        o = fit_oracle.gaussian_fit(im[X[0], X[1], X[2]], X, center=list(c), version=3, delta_center=2.5, weight_sigma=ws)
        assert np.array_equal(o["p"], g.p)
    # alternative seeders of Fitting_v4 (External/Fitting_v4.py:66-126); pyfftw is stubbed with scipy.fft (ref_loader)
    imf = im.astype(np.float32)
    out["fftg_5"] = ns.Fitting_v4.fft_gaussian_fast(imf, gaus=[2.5, 5, 5])
    c1, s1 = ns.Fitting_v4.get_seed_points_base(imf, gfilt_size=2.5, th_seed=3.)
    out["lr_centers"], out["lr_std"] = c1, np.float64(s1)
    c1, s1 = ns.Fitting_v4.get_seed_points_base(imf, gfilt_size=2.5, th_seed=4., filt_size=5, max_num=12)
    out["lr_centers_f5_top12"] = c1
    if cv2 is not None:
        for tag, arr, kw in (("u16", im, dict(th_seed=3.)), ("f32_g7_f5", imf, dict(gfilt_size=7, filt_size=5, th_seed=2.5)),
                             ("u16_top10", im, dict(th_seed=3., max_num=10))):
            c2, s2 = ns.Fitting_v4.get_seed_points_base_v2(arr, **kw)
            out[f"v2_{tag}_centers"], out[f"v2_{tag}_std"] = c2, np.float32(s2)
        out["v2_norm_u16_g5"] = ns.Fitting_v4.normalzie_im(im, 5)[:3]          # first slices: pins the cv2.blur arithmetic
    np.savez_compressed(os.path.join(OUT, "extra_r2.npz"), **out)
    print("extra_r2.npz", os.path.getsize(os.path.join(OUT, "extra_r2.npz")), {k: np.shape(v) for k, v in out.items() if k != "meta"})


def corr_case(seed=5, Z=8, X=40, Y=48, nbuf=2):
    """a synthetic three-colour .dax movie with hot columns, its correction profiles and the frames array"""
    rng = np.random.default_rng(seed)
    chs = ['750', '647', '561']
    ims = [synth((Z, X, Y), 10, 20 + i) for i in range(3)]
    gain = 1.0 + 0.05 * np.arange(Z)[:, None, None]                        # planes of different brightness: Z_Shift_Correction has work to do
    ims = [np.clip(im * gain, 0, 65535).astype(np.uint16) for im in ims]
    for i, im in enumerate(ims):
        im[:, 10 + i, 12] = 30000                  # isolated hot columns
        im[:, 20, 20 + i] = 9000                   # two adjacent ones, both detected (the y + 1 neighbour is not looked at):
        im[:, 20, 21 + i] = 29000                  # the sequential fix of the second reads the replaced first
        im[:, 0, 5] = 31000                        # on the border: detected (np.roll wraps) but never replaced
    frames = np.zeros((2 * nbuf + Z * 3, X, Y), np.uint16)
    for c in range(3):
        s0 = nbuf + (c - nbuf) % 3
        frames[s0:s0 + Z * 3:3] = ims[c]
    illum = {ch: (1 + 0.2 * rng.random((X, Y))).astype(np.float32) for ch in chs}
    bleed = (np.eye(3)[:, :, None, None] + 0.05 * rng.random((3, 3, X, Y))).astype(np.float32)
    xx, yy = np.meshgrid(np.arange(X), np.arange(Y), indexing='ij')
    chrom = {ch: np.stack([0.2 * np.ones((1, X, Y)), (0.5 * np.sin(xx / 9.))[None], (0.4 * np.cos(yy / 7.))[None]]).astype(np.float32)
             if ch != '647' else None for ch in chs}
    return chs, ims, frames, illum, bleed, chrom, nbuf


def main_corr():
    """tests/golden/corr_r2.npz: io_tools/load.py correct_fov_image, the UNMODIFIED reference function (lifted with
    its helpers by ref_loader.load_corrections) run on a synthetic .dax movie; every case is also asserted equal to
    oracle/correct_oracle.py, which pins that restatement."""
    import tempfile
    from oracle import correct_oracle
    warnings.simplefilter("ignore")
    ns = ref_loader.load_corrections()
    ref_loader.load_chromatic()
    chs, ims, frames, illum, bleed, chrom, nbuf = corr_case()
    Z, X, Y = ims[0].shape
    d = tempfile.mkdtemp()
    fn = os.path.join(d, 'Conv_zscan_00.dax')
    frames.tofile(fn)
    with open(fn.replace('.dax', '.inf'), 'w') as fh:
        fh.write(f"frame dimensions = {Y} x {X}\nnumber of frames = {frames.shape[0]}\n little endian\n")
    kw = dict(single_im_size=[Z, X, Y], all_channels=chs, num_buffer_frames=nbuf, num_empty_frames=0, corr_channels=chs,
              illumination_profile=illum, bleed_profile=bleed, chromatic_profile=chrom, drift_channel='561')
    off = dict(hot_pixel_corr=False, bleed_corr=False, illumination_corr=False, chromatic_corr=False)
    drift = [0.4, -1.3, 2.2]
    cases = {
        "all_drift": (['750', '647'], drift, {}, True),
        "all_nodrift": (['561'], None, {}, True),
        "all_quiet": (['750', '647'], drift, {}, False),            # verbose=False: the reference does not warp
        "hot_only": (['750', '647', '561'], None, {**off, 'hot_pixel_corr': True}, True),
        "bleed_only": (['750', '647'], None, {**off, 'bleed_corr': True}, True),
        "illum_only": (['750', '561'], None, {**off, 'illumination_corr': True}, True),
        "chrom_only": (['750', '561'], None, {**off, 'chromatic_corr': True}, True),
        "drift_only": (['647'], [-2.6, 3.3, -0.7], off, True),
        "drift_big": (['647'], [11.0, -30.5, 60.25], off, True),    # far beyond the padding: mode='nearest' clamps
        "zshift_only": (['750', '561'], None, {**off, 'z_shift_corr': True}, True),
        "zshift_all": (['750', '647'], drift, {'z_shift_corr': True}, True),
        "highpass_only": (['750', '561'], None, {**off, 'gaussian_highpass': True}, True),
        "highpass_all": (['750', '647'], drift, {'gaussian_highpass': True, 'gauss_sigma': 2.2, 'gauss_truncate': 3}, True),
    }
    out = dict(frames=frames, bleed=bleed, meta=np.array(repr(dict(numpy=np.__version__, scipy=scipy.__version__))))
    for ch in chs:
        out[f"illum_{ch}"] = illum[ch]
        if chrom[ch] is not None:
            out[f"chrom_{ch}"] = chrom[ch]
    names = []
    for tag, (sel, dr, flags, verbose) in cases.items():
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            got, = ns.correct_fov_image(fn, sel, drift=dr, verbose=verbose, **{**kw, **flags})
        mine = correct_oracle.correct_stacks(ims, chs, sel, chs, drift=dr, illumination_profile=illum, bleed_profile=bleed,
                                             chromatic_profile=chrom, verbose=verbose, **flags)
        assert all(np.array_equal(a, b) for a, b in zip(got, mine)), tag
        for ch, a in zip(sel, got):
            assert a.dtype == np.uint16
            out[f"{tag}__{ch}"] = a
        names.append(f"{tag}|{','.join(sel)}|{repr(dr)}|{repr(flags)}|{int(verbose)}")
        print(tag, "reference == oracle on", sel)
    out["cases"] = np.array(names)
    # warp_image=False: unwarped images + one spot-coordinate function per channel (correction_tools/chromatic.py:41-114)
    consts = {'750': dict(constants=[np.array([0.2, 1e-3, -2e-3, 5e-4]), np.array([-0.1, 2e-3, 1e-3, 0., 1e-5, 0., 0., 2e-5, 0., -1e-5]), np.array([0.3])],
                          fitting_orders=np.array([1, 2, 0]), ref_center=np.array([4., 20., 24.])),
              '647': None, '561': dict(constants=[np.array([0.]), np.array([0.05]), np.array([-0.02, 1e-3, 0., 0.])],
                                       fitting_orders=np.array([0, 0, 1]), ref_center=np.array([4., 20., 24.]))}
    pts = np.random.default_rng(3).uniform([0, 0, 0], [Z, X, Y], size=(40, 3))
    table = np.concatenate([np.full((40, 1), 900.), pts, np.ones((40, 7))], axis=1).astype(np.float32)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        ims_nw, funcs = ns.correct_fov_image(fn, ['750', '647', '561'], drift=drift, warp_image=False, verbose=True,
                                             **{**kw, 'chromatic_profile': consts})
        ims_nw0, funcs0 = ns.correct_fov_image(fn, ['647'], drift=None, warp_image=False, verbose=True, **{**kw, 'chromatic_profile': consts})
    out["nowarp_pts"], out["nowarp_table"] = pts, table
    for ch, im, f in zip(['750', '647', '561'], ims_nw, funcs):
        assert np.array_equal(im, out[f"all_quiet__{ch}"]) if f"all_quiet__{ch}" in out else True
        out[f"nowarp_pts__{ch}"], out[f"nowarp_table__{ch}"] = f(pts), f(table)
    out["nowarp_identity"] = funcs0[0](pts)
    assert np.array_equal(out["nowarp_identity"], pts)
    out["nowarp_consts"] = np.array(repr({k: (None if v is None else {a: np.asarray(b).tolist() if a != 'constants' else [c.tolist() for c in b] for a, b in v.items()})
                                          for k, v in consts.items()}))
    np.savez_compressed(os.path.join(OUT, "corr_r2.npz"), **out)
    print("corr_r2.npz", os.path.getsize(os.path.join(OUT, "corr_r2.npz")))


ALIGN_CASE = dict(shape=(16, 256, 256), n=300, drift=(0.35, 2.4, -1.7), seed=41)


def write_dax(path, channels_ims, nbuf):
    """interleaved movie of the given channel stacks (in all_channels order) with nbuf buffer frames at both ends"""
    ncol, (Z, X, Y) = len(channels_ims), channels_ims[0].shape
    frames = np.zeros((2 * nbuf + Z * ncol, X, Y), np.uint16)
    for c, im in enumerate(channels_ims):
        s0 = nbuf + (c - nbuf) % ncol
        frames[s0:s0 + Z * ncol:ncol] = im
    frames.tofile(path)
    with open(path.replace('.dax', '.inf'), 'w') as fh:
        fh.write(f"frame dimensions = {Y} x {X}\nnumber of frames = {frames.shape[0]}\n little endian\n")


def align_files(folder, ref, src, other_ref, other_src, nbuf=2):
    """two .dax movies (channels '647' = a signal image, '488' = beads) and an illumination profile folder"""
    X, Y = ref.shape[1:]
    write_dax(os.path.join(folder, 'ref.dax'), [other_ref, ref], nbuf)
    write_dax(os.path.join(folder, 'src.dax'), [other_src, src], nbuf)
    xx, yy = np.meshgrid(np.arange(X), np.arange(Y), indexing='ij')
    for ch, amp in (('647', 0.15), ('488', 0.1)):
        pf = (1.0 + amp * np.cos((xx - X / 2) / X) * np.cos((yy - Y / 2) / Y)).astype(np.float32)
        np.save(os.path.join(folder, f'illumination_correction_{ch}_{X}x{Y}.npy'), pf)
    return os.path.join(folder, 'src.dax'), os.path.join(folder, 'ref.dax')


def main_align():
    """tests/golden/align_r2.npz: correction_tools/alignment.py align_image in its bead-fitting mode (use_autocorr=False; the
    phase-correlation mode needs scikit-image, absent here), the UNMODIFIED reference functions lifted by
    ref_loader.load_alignment, on a synthetic bead pair (imageanalysis3_b200.synth.bead_pair: images are regenerated in the
    tests from the stored parameters, a checksum guards the generator), and correct_fov_image(calculate_drift=True) on two
    .dax movies built from it."""
    import contextlib, io, tempfile
    from imageanalysis3_b200.synth import bead_pair
    warnings.simplefilter("ignore")
    ns = ref_loader.load_alignment()
    cor = ref_loader.load_corrections()
    c = ALIGN_CASE
    ref, src, centers = bead_pair(c["shape"], c["n"], c["drift"], c["seed"])
    out = dict(shape=np.array(c["shape"]), n=c["n"], drift_planted=np.array(c["drift"]), seed=c["seed"],
               checksum=np.array([int(ref.astype(np.uint64).sum()), int(src.astype(np.uint64).sum()), int((ref.astype(np.int64) * 3 + src).std() * 1e6)]),
               meta=np.array(repr(dict(numpy=np.__version__, scipy=scipy.__version__))))
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        out["crops"] = ns.generate_drift_crops(list(c["shape"]))
        d0, f0 = ns.align_image(src, ref, use_autocorr=False, correction_args=dict(single_im_size=list(c["shape"])), verbose=False)
        d1, f1 = ns.align_image(src, ref, use_autocorr=False, drift_diff_th=1e-4, correction_args=dict(single_im_size=list(c["shape"])), verbose=False)
        d2, f2 = ns.align_image(src, ref, use_autocorr=False, crop_list=out["crops"][[5, 2, 7]], min_good_drifts=2, match_distance_th=1.5,
                                fitting_args=dict(max_num_seeds=12), correction_args=dict(single_im_size=list(c["shape"])), verbose=False)
    out.update(drift_default=d0, flag_default=f0, drift_suboptimal=d1, flag_suboptimal=f1, drift_custom=d2, flag_custom=f2)
    print("planted", c["drift"], "reference finds", -d0, f0, "| sub-optimal path", -d1, f1, "| custom", -d2, f2)
    assert f0 == 0 and f1 == 1 and np.abs(d0 + np.array(c["drift"])).max() < 0.05
    # one crop's pairing, step by step (inputs of the host functions)
    s = tuple(slice(*r) for r in out["crops"][0])
    with contextlib.redirect_stdout(sink):
        sp_src = ns_fit_centers(src[s])
        sp_ref = ns_fit_centers(ref[s])
        rough = ns.fft3d_from2d(src[s], ref[s], gb=0, max_disp=np.max(src[s].shape) / 2)
        dft, p_t, p_r = ns.find_paired_centers(sp_src, sp_ref, rough, cutoff=2., return_paired_cts=True)
        dft2, k_t, k_r = ns.check_paired_centers(p_t, p_r, outlier_sigma=1.5, return_paired_cts=True)
    out.update(crop0_src_cts=sp_src, crop0_ref_cts=sp_ref, crop0_rough=rough, crop0_drift_paired=dft, crop0_paired_tar=p_t, crop0_paired_ref=p_r,
               crop0_drift_checked=dft2, crop0_kept_tar=k_t, crop0_kept_ref=k_r)
    # through files: correct_fov_image(calculate_drift=True) of a two-colour movie
    other_ref = synth(c["shape"], 40, 77)
    other_src = synth(c["shape"], 40, 78)
    folder = tempfile.mkdtemp()
    src_dax, ref_dax = align_files(folder, ref, src, other_ref, other_src)
    with contextlib.redirect_stdout(sink):
        ims, drift, flag = cor.correct_fov_image(src_dax, ['647'], single_im_size=list(c["shape"]), all_channels=['647', '488'], num_buffer_frames=2,
                                                 num_empty_frames=0, calculate_drift=True, drift_channel='488', ref_filename=ref_dax, use_autocorr=False,
                                                 corr_channels=['647'], correction_folder=folder, bleed_corr=False, chromatic_corr=False,
                                                 return_drift=True, verbose=True)
    out.update(file_drift=drift, file_flag=flag, file_im_647_slab=ims[0][4:12, 64:192, 64:192], file_im_647_sum=int(ims[0].astype(np.uint64).sum()),
               file_other_seeds=np.array([77, 78]))
    print("through files: drift", drift, flag, "corrected image", ims[0].shape, ims[0].dtype)
    np.savez_compressed(os.path.join(OUT, "align_r2.npz"), **out)
    print("align_r2.npz", os.path.getsize(os.path.join(OUT, "align_r2.npz")))


def ns_fit_centers(im):
    ns = ref_loader.load()
    al = ref_loader.load_alignment()
    spots = ns.fitting.fit_fov_image(im, '488', verbose=False, **al.defaults[1])
    return ns.fitting.select_sparse_centers(spots[:, 1:4], 2.)


if __name__ == "__main__":
    if "align" in sys.argv[1:]:
        main_align()
    elif "corr" in sys.argv[1:]:
        main_corr()
    elif "extra" in sys.argv[1:]:
        main_extra()
    else:
        main()
