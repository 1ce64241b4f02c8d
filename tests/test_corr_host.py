"""CPU: the correct_fov_image oracle against the reference-made fixture (tests/golden/corr_r2.npz, written by
``python -m oracle.make_golden corr`` from the unmodified reference function), and the host side of the mirror
(io_tools/load.py: .dax reading, channel splitting, argument checks)."""
import ast
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def corr():
    return np.load(os.path.join(GOLDEN, "corr_r2.npz"))


def parse_case(line):
    tag, sel, drift, flags, verbose = line.split("|")
    return tag, sel.split(","), ast.literal_eval(drift), ast.literal_eval(flags), bool(int(verbose))


def write_movie(tmp_path, g):
    frames = g["frames"]
    fn = str(tmp_path / "Conv_zscan_00.dax")
    frames.tofile(fn)
    with open(fn.replace(".dax", ".inf"), "w") as fh:
        fh.write(f"frame dimensions = {frames.shape[2]} x {frames.shape[1]}\nnumber of frames = {frames.shape[0]}\n little endian\n")
    return fn


def profiles(g):
    chs = ['750', '647', '561']
    illum = {ch: g[f"illum_{ch}"] for ch in chs}
    chrom = {ch: (g[f"chrom_{ch}"] if f"chrom_{ch}" in g.files else None) for ch in chs}
    return chs, illum, g["bleed"], chrom


def test_oracle_reproduces_every_reference_case(corr, tmp_path):
    from imageanalysis3_b200.io_tools import load
    from oracle import correct_oracle
    chs, illum, bleed, chrom = profiles(corr)
    fn = write_movie(tmp_path, corr)
    raw = load.read_dax(fn)
    assert np.array_equal(raw, corr["frames"])
    shape, n_color = load.get_num_frame(fn, frame_per_color=8, buffer_frame=2, empty_frame=0)
    assert shape == [raw.shape[0], raw.shape[2], raw.shape[1]] and n_color == 3     # the .inf order: height x width
    ims = load.split_im_by_channels(raw, chs, chs, single_im_size=[8, 40, 48], num_buffer_frames=2)
    assert all(im.shape == (8, 40, 48) and im.flags.c_contiguous for im in ims)
    assert len(corr["cases"]) >= 9
    for line in corr["cases"]:
        tag, sel, drift, flags, verbose = parse_case(str(line))
        got = correct_oracle.correct_stacks(ims, chs, sel, chs, drift=drift, illumination_profile=illum, bleed_profile=bleed,
                                            chromatic_profile=chrom, verbose=verbose, **flags)
        for ch, a in zip(sel, got):
            assert np.array_equal(a, corr[f"{tag}__{ch}"]), (tag, ch)


def test_reference_warps_only_when_verbose(corr):
    """the quirk the mirror keeps: io_tools/load.py:436-459 sit under ``if verbose:``"""
    assert not np.array_equal(corr["all_drift__750"], corr["all_quiet__750"])


def test_split_handles_buffer_offsets():
    from imageanalysis3_b200.io_tools import load
    n_col, Z, nbuf = 4, 5, 3
    movie = np.arange((2 * nbuf + Z * n_col) * 2 * 2, dtype=np.uint16).reshape(-1, 2, 2)
    chs = ['750', '647', '561', '488']
    out = load.split_im_by_channels(movie, ['561', '750'], chs, single_im_size=[Z, 2, 2], num_buffer_frames=nbuf)
    for ch, im in zip(['561', '750'], out):
        start = nbuf + (chs.index(ch) - nbuf) % n_col
        assert np.array_equal(im, movie[start:start + Z * n_col:n_col])
    with pytest.raises(ValueError):
        load.split_im_by_channels(movie, ['405'], chs, single_im_size=[Z, 2, 2], num_buffer_frames=nbuf)


def test_big_endian_movie(tmp_path):
    from imageanalysis3_b200.io_tools import load
    frames = (np.arange(3 * 4 * 5, dtype=np.uint16) * 257).reshape(3, 4, 5)
    fn = str(tmp_path / "m.dax")
    frames.astype('>u2').tofile(fn)
    with open(fn.replace(".dax", ".inf"), "w") as fh:
        fh.write("frame dimensions = 5 x 4\nnumber of frames = 3\ndata type = 16 bit integers (binary, big endian)\n")
    assert np.array_equal(load.read_dax(fn), frames)


def test_argument_errors_follow_the_reference(corr, tmp_path):
    from imageanalysis3_b200.io_tools import load
    chs, illum, bleed, chrom = profiles(corr)
    fn = write_movie(tmp_path, corr)
    kw = dict(single_im_size=[8, 40, 48], all_channels=chs, num_buffer_frames=2, corr_channels=chs, illumination_profile=illum,
              bleed_profile=bleed, chromatic_profile=chrom, drift_channel='561', verbose=False)
    with pytest.raises(IOError):
        load.correct_fov_image(str(tmp_path / "missing.dax"), ['750'], **kw)
    with pytest.raises(IndexError):
        load.correct_fov_image(fn, ['750'], drift=[1, 2], **kw)
    with pytest.raises(ValueError):
        load.correct_fov_image(fn, ['750'], **{**kw, 'drift_channel': '405'})
    with pytest.raises(TypeError):
        load.correct_fov_image(fn, ['750'], **{**kw, 'illumination_profile': [1]})
    with pytest.raises(KeyError):
        load.correct_fov_image(fn, ['750'], **{**kw, 'illumination_profile': {'750': illum['750']}})
    with pytest.raises(IndexError):
        load.correct_fov_image(fn, ['750'], **{**kw, 'bleed_profile': bleed[:2, :2]})
    with pytest.raises(KeyError):
        load.correct_fov_image(fn, ['750'], **{**kw, 'chromatic_profile': {'647': None}})
    for flag in ('calculate_drift', 'normalization'):      # calculate_drift: with the default use_autocorr=True
        with pytest.raises(NotImplementedError):
            load.correct_fov_image(fn, ['750'], **{**kw, flag: True})


def nowarp_consts(g):
    raw = ast.literal_eval(str(g["nowarp_consts"]))
    return {ch: None if v is None else dict(constants=[np.array(c) for c in v['constants']], fitting_orders=np.array(v['fitting_orders']),
                                            ref_center=np.array(v['ref_center'])) for ch, v in raw.items()}


def test_chromatic_coordinate_functions_match_reference(corr, tmp_path):
    """generate_chromatic_function (correction_tools/chromatic.py:41-114): what correct_fov_image(warp_image=False) hands back"""
    import pickle
    from imageanalysis3_b200.correction_tools import chromatic
    consts = nowarp_consts(corr)
    drift = np.array([0.4, -1.3, 2.2], dtype=np.float32)
    pts, table = corr["nowarp_pts"], corr["nowarp_table"]
    for ch in ('750', '647', '561'):
        f = chromatic.generate_chromatic_function(consts[ch], drift)
        assert np.array_equal(f(pts), corr[f"nowarp_pts__{ch}"]) and np.array_equal(f(table), corr[f"nowarp_table__{ch}"]), ch
        assert f(table).dtype == corr[f"nowarp_table__{ch}"].dtype
    with open(tmp_path / "c.pkl", "wb") as fh:
        pickle.dump(consts['750'], fh)
    assert np.array_equal(chromatic.generate_chromatic_function(str(tmp_path / "c.pkl"), drift)(pts), corr["nowarp_pts__750"])
    ident = chromatic.generate_chromatic_function(None, None)
    assert ident(pts) is pts and len(chromatic.generate_chromatic_function(consts['750'])([])) == 0
    with pytest.raises(TypeError):
        chromatic.generate_chromatic_function(3)
    with pytest.raises(ValueError):
        chromatic.generate_chromatic_function(consts['750'])(np.zeros((4, 5)))
    X = chromatic.generate_polynomial_data(np.array([[1., 2., 3.], [2., 0., -1.]]), 2)
    assert X.shape == (2, 10) and np.array_equal(X[0], [1, 1, 2, 3, 1, 2, 3, 4, 6, 9])
