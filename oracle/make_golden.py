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


if __name__ == "__main__":
    main()
