"""GPU parity of correct_fov_image (io_tools/load.py mirror over ia3_corr_*) with the unmodified reference's outputs
(tests/golden/corr_r2.npz) and with the oracle on fresh seeded inputs; size-independent properties at full size."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_corr_host import parse_case, profiles, write_movie

pytestmark = pytest.mark.gpu

# the spline coefficients are floating point: ours and scipy's agree to ~1e-11 of a count, so the rounded uint16
# output may differ by one count where a value lies that close to a half-integer -- nowhere in practice
WARP_MAX_DIFF = 1
WARP_MAX_FRAC = 1e-5


def same_warp(a, b, what):
    d = np.abs(a.astype(np.int64) - b.astype(np.int64))
    assert d.max() <= WARP_MAX_DIFF and (d > 0).mean() <= WARP_MAX_FRAC, f"{what}: max diff {d.max()}, {(d > 0).sum()} voxels differ"


@pytest.fixture(scope="module")
def corr():
    return np.load(os.path.join(GOLDEN, "corr_r2.npz"))


def test_correct_fov_image_matches_reference_fixture(lib, corr, tmp_path):
    from imageanalysis3_b200.io_tools import load
    chs, illum, bleed, chrom = profiles(corr)
    fn = write_movie(tmp_path, corr)
    kw = dict(single_im_size=[8, 40, 48], all_channels=chs, num_buffer_frames=2, num_empty_frames=0, corr_channels=chs,
              illumination_profile=illum, bleed_profile=bleed, chromatic_profile=chrom, drift_channel='561')
    n_exact = n_total = 0
    for line in corr["cases"]:
        tag, sel, drift, flags, verbose = parse_case(str(line))
        got, = load.correct_fov_image(fn, sel, drift=drift, verbose=verbose, **{**kw, **flags})
        warped = verbose and (drift is not None or flags.get('chromatic_corr', True))
        for ch, a in zip(sel, got):
            want = corr[f"{tag}__{ch}"]
            assert a.dtype == np.uint16 and a.shape == want.shape
            if warped:
                same_warp(a, want, f"{tag} {ch}")
            else:
                assert np.array_equal(a, want), (tag, ch)          # integer work: bit-exact
            n_exact += int(np.array_equal(a, want))
            n_total += 1
    assert n_exact == n_total, f"{n_total - n_exact} warped stacks differ from the reference by one count somewhere"
    # return_drift and resident outputs
    (st,), d, flag = load.correct_fov_image(fn, ['647'], drift=[0.4, -1.3, 2.2], return_drift=True, return_stacks=True, verbose=False, force_warp=True, **kw)
    assert flag == 0 and np.array_equal(d, np.array([0.4, -1.3, 2.2], dtype=np.float32))
    assert np.array_equal(st.fetch(), corr["all_drift__647"])


@pytest.mark.parametrize("shape,seed", [((7, 33, 130), 1), ((12, 96, 64), 2)])
def test_each_step_matches_oracle_on_fresh_inputs(lib, shape, seed):
    from imageanalysis3_b200.io_tools import load
    from imageanalysis3_b200.synth import synth
    from oracle import correct_oracle
    rng = np.random.default_rng(seed)
    Z, X, Y = shape
    chs = ['750', '647', '561']
    ims = [synth(shape, 12, 40 + seed * 3 + i) for i in range(3)]
    for i, im in enumerate(ims):
        for _ in range(6):
            x, y = rng.integers(0, X), rng.integers(0, Y)
            im[:, x, y] = 15000
            if rng.random() < 0.5:
                im[:, x, min(y + 1, Y - 1)] = 50000          # also detected; its fix reads the replaced (x, y)
    for dt in (np.float32, np.float64):
        illum = {ch: (0.7 + 0.6 * rng.random((X, Y))).astype(dt) for ch in chs}
        bleed = (np.eye(3)[:, :, None, None] * 1.3 + 0.1 * rng.standard_normal((3, 3, X, Y))).astype(dt)   # reaches both clip ends
        chrom = {ch: (rng.standard_normal((3, 1, X, Y)) * 0.8).astype(np.float32) if ch != '647' else None for ch in chs}
        for sel, drift in ((['750', '561'], None), (['647', '561'], [0.25, -3.5, 1.75])):
            want = correct_oracle.correct_stacks(ims, chs, sel, chs, drift=drift, illumination_profile=illum, bleed_profile=bleed, chromatic_profile=chrom)
            got = load.correct_image_stacks(ims, chs, sel, chs, drift=drift, illumination_profile=illum, bleed_profile=bleed, chromatic_profile=chrom)
            for ch, a, b in zip(sel, got, want):
                same_warp(a, b, f"{shape} {dt.__name__} {ch}")
            # without the warp everything is integer work
            want = correct_oracle.correct_stacks(ims, chs, sel, chs, illumination_profile=illum, bleed_profile=bleed, chromatic_profile=chrom, warp_image=False)
            got = load.correct_image_stacks(ims, chs, sel, chs, illumination_profile=illum, bleed_profile=bleed, chromatic_profile=chrom, warp=False)
            assert all(np.array_equal(a, b) for a, b in zip(got, want))
    # z-shift correction: planes of different brightness scaled to the stack's median (exact medians from device histograms)
    dim = [np.clip(im * (0.6 + 0.05 * np.arange(Z))[:, None, None], 0, 65535).astype(np.uint16) for im in ims]
    st = lib.Stack(dim[0])
    st.z_shift_correct()
    assert np.array_equal(st.fetch(), correct_oracle.z_shift_correction(dim[0].astype(np.float32)))
    want = correct_oracle.correct_stacks(dim, chs, ['750', '561'], chs, drift=[0.5, 0.5, -0.5], z_shift_corr=True, illumination_profile=illum,
                                         bleed_profile=bleed, chromatic_profile=chrom)
    got = load.correct_image_stacks(dim, chs, ['750', '561'], chs, drift=[0.5, 0.5, -0.5], z_shift_corr=True, illumination_profile=illum,
                                    bleed_profile=bleed, chromatic_profile=chrom)
    for ch, a, b in zip(['750', '561'], got, want):
        same_warp(a, b, f"z-shift + everything else, {ch}")
    # Gaussian high pass: exact uint16 Gaussian with mode='nearest' (several sigmas / truncations, radius up to the plane count and beyond)
    for sigma, trunc in ((3, 2), (5, 2), (1.3, 4), (4.5, 3)):
        st = lib.Stack(ims[1])
        st.gaussian_highpass(sigma, trunc)
        assert np.array_equal(st.fetch(), correct_oracle.gaussian_high_pass(ims[1], sigma, trunc)), (sigma, trunc)
    want = correct_oracle.correct_stacks(ims, chs, ['647'], chs, drift=[0.5, 0.5, -0.5], gaussian_highpass=True, illumination_profile=illum,
                                         bleed_profile=bleed, chromatic_profile=chrom)
    got = load.correct_image_stacks(ims, chs, ['647'], chs, drift=[0.5, 0.5, -0.5], gaussian_highpass=True, illumination_profile=illum,
                                    bleed_profile=bleed, chromatic_profile=chrom)
    same_warp(got[0], want[0], "high pass after the warp")
    # a per-plane chromatic profile (3, Z, X, Y)
    chrom_z = {ch: (rng.standard_normal((3, Z, X, Y)) * 0.5).astype(np.float32) if ch != '647' else None for ch in chs}
    off = dict(hot_pixel_corr=False, bleed_corr=False, illumination_corr=False)
    want = correct_oracle.correct_stacks(ims, chs, ['750'], chs, chromatic_profile=chrom_z, **off)
    got = load.correct_image_stacks(ims, chs, ['750'], chs, chromatic_profile=chrom_z, **off)
    same_warp(got[0], want[0], "per-plane chromatic profile")
    # what correction_tools/chromatic.py saves: float64, one plane per z; uploaded once, reused by the second call
    chrom_64 = {ch: None if v is None else v.astype(np.float64) for ch, v in chrom_z.items()}
    want = correct_oracle.correct_stacks(ims, chs, ['561'], chs, drift=[-0.5, 0.25, 3.0], chromatic_profile=chrom_64, **off)
    for _ in range(2):
        got = load.correct_image_stacks(ims, chs, ['561'], chs, drift=[-0.5, 0.25, 3.0], chromatic_profile=chrom_64, **off)
        same_warp(got[0], want[0], "float64 per-plane chromatic profile")
    assert any(ref() is chrom_64['561'] for ref, _, _ in load._RESIDENT.values())
    chrom_64['561'][1] += 0.75                                       # edited in place: must not be served from the device copy
    want = correct_oracle.correct_stacks(ims, chs, ['561'], chs, chromatic_profile=chrom_64, **off)
    got = load.correct_image_stacks(ims, chs, ['561'], chs, chromatic_profile=chrom_64, **off)
    same_warp(got[0], want[0], "edited chromatic profile")


def test_full_size_properties(lib):
    """at the reference's stack size (30 x 2048 x 2048): an integer drift is an exact shift with edge replication,
    identity profiles leave the stack unchanged, and a stack without hot columns is untouched"""
    rng = np.random.default_rng(3)
    shape = (30, 2048, 2048)
    im = rng.integers(100, 3000, size=shape, dtype=np.uint16)
    st = lib.Stack(im)
    assert st.remove_hot_pixels() == 0 and np.array_equal(st.fetch(), im)
    out = lib.Stack.mix([st], illum=np.ones(shape[1:], dtype=np.float32)).fetch()
    assert np.array_equal(out, im)
    eye = np.zeros((2,) + shape[1:], dtype=np.float32)
    eye[0] = 1
    assert np.array_equal(lib.Stack.mix([st, st], bleed=eye).fetch(), im)
    dz, dx, dy = 2, -3, 5
    got = st.warp(drift=[dz, dx, dy]).fetch()
    zi = np.clip(np.arange(shape[0]) - dz, 0, shape[0] - 1)
    xi = np.clip(np.arange(shape[1]) - dx, 0, shape[1] - 1)
    yi = np.clip(np.arange(shape[2]) - dy, 0, shape[2] - 1)
    assert np.array_equal(got, im[zi][:, xi][:, :, yi])
    # hot columns at full size: count and replacement against the oracle's rule on the touched columns only
    im2 = im.copy()
    cols = [(5, 7), (1000, 1000), (1000, 1001), (2046, 2046), (0, 9)]
    for x, y in cols:
        im2[:, x, y] = 65535
    im2[:, 1000, 1000] = 20000                  # (1000, 1001) is still detected: only the y - 1 neighbour counts, twice
    st2 = lib.Stack(im2)
    n = st2.remove_hot_pixels()
    got = st2.fetch()
    from oracle import correct_oracle
    crop = im2[:, 990:1012, 990:1012]
    want = correct_oracle.remove_hot_pixels(crop.astype(np.float32))
    assert n == len(cols)
    assert np.array_equal(got[:, 992:1010, 992:1010], want[:, 2:-2, 2:-2].astype(np.uint16))
    assert np.array_equal(got[:, 0, 9], im2[:, 0, 9])                        # border column: found, never replaced


def test_corrected_stacks_feed_fit_fov_image_without_leaving_the_device(lib, corr, tmp_path):
    """correct_fov_image(return_stacks=True) -> fit_fov_image(stack, ...) equals the route through numpy arrays"""
    from imageanalysis3_b200.io_tools import load
    from imageanalysis3_b200.spot_tools.fitting import fit_fov_image
    chs, illum, bleed, chrom = profiles(corr)
    fn = write_movie(tmp_path, corr)
    kw = dict(single_im_size=[8, 40, 48], all_channels=chs, num_buffer_frames=2, corr_channels=chs, illumination_profile=illum,
              bleed_profile=bleed, chromatic_profile=chrom, drift_channel='561', drift=[0.4, -1.3, 2.2], verbose=False, force_warp=True)
    (arrays,) = load.correct_fov_image(fn, ['750', '647'], **kw)
    (stacks,) = load.correct_fov_image(fn, ['750', '647'], return_stacks=True, **kw)
    fkw = dict(th_seed=50., max_num_seeds=20, verbose=False)
    for ch, a, st in zip(['750', '647'], arrays, stacks):
        want = fit_fov_image(a, ch, **fkw)
        got = fit_fov_image(st, ch, **fkw)
        assert len(want) > 0 and np.array_equal(got, want)
        assert np.array_equal(fit_fov_image(st, ch, normalize_local=True, **fkw), fit_fov_image(a, ch, normalize_local=True, **fkw))
        assert np.array_equal(st.fetch(), a)                      # the caller's stack keeps its image
        with pytest.raises(NotImplementedError):
            fit_fov_image(st, ch, seeding_kwargs=dict(sel_center=[4, 20, 20]), **fkw)


def test_thousands_of_hot_columns(lib):
    """a camera with many hot columns (or a low hot_th): the sequential replacement still follows np.where order"""
    from oracle import correct_oracle
    rng = np.random.default_rng(9)
    Z, X, Y = 6, 240, 250
    im = rng.integers(200, 400, size=(Z, X, Y), dtype=np.uint16)
    xs, ys = rng.integers(0, X, 3000), rng.integers(0, Y - 1, 3000)
    im[:, xs, ys] = 15000
    half = rng.random(3000) < 0.5
    im[:, xs[half], ys[half] + 1] = 60000            # runs of adjacent hot columns: later fixes read earlier ones
    want = correct_oracle.remove_hot_pixels(im.astype(np.float32))
    st = lib.Stack(im)
    n = st.remove_hot_pixels()
    f = im.astype(np.float32)
    conv = (np.roll(f, 1, 1) + np.roll(f, -1, 1) + np.roll(f, 1, 2) + np.roll(f, 1, 2)) / 4
    assert n == int((np.sum(f > 4 * conv, 0) > 0.5 * Z).sum()) and n > 2500
    assert np.array_equal(st.fetch(), want.astype(np.uint16))


def test_warp_image_false_returns_images_and_coordinate_functions(lib, corr, tmp_path):
    from imageanalysis3_b200.io_tools import load
    from test_corr_host import nowarp_consts
    chs, illum, bleed, chrom = profiles(corr)
    fn = write_movie(tmp_path, corr)
    kw = dict(single_im_size=[8, 40, 48], all_channels=chs, num_buffer_frames=2, num_empty_frames=0, corr_channels=chs,
              illumination_profile=illum, bleed_profile=bleed, chromatic_profile=nowarp_consts(corr), drift_channel='561')
    ims, funcs, drift, flag = load.correct_fov_image(fn, ['750', '647', '561'], drift=[0.4, -1.3, 2.2], warp_image=False, return_drift=True, verbose=True, **kw)
    assert flag == 0 and drift.dtype == np.float32 and len(funcs) == 3
    for ch, im, f in zip(['750', '647', '561'], ims, funcs):
        if f"all_quiet__{ch}" in corr.files:
            assert np.array_equal(im, corr[f"all_quiet__{ch}"])            # corrected, not warped
        assert np.array_equal(f(corr["nowarp_pts"]), corr[f"nowarp_pts__{ch}"]) and np.array_equal(f(corr["nowarp_table"]), corr[f"nowarp_table__{ch}"])
    (ims0, funcs0) = load.correct_fov_image(fn, ['647'], drift=None, warp_image=False, verbose=False, **kw)
    assert funcs0[0](corr["nowarp_pts"]) is corr["nowarp_pts"] or np.array_equal(funcs0[0](corr["nowarp_pts"]), corr["nowarp_identity"])
