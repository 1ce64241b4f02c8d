import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box: pytest -m gpu)")


@pytest.fixture(scope="session")
def golden_seeds():
    return np.load(os.path.join(GOLDEN, "seeds_small.npz"))


@pytest.fixture(scope="session")
def golden_fits():
    return np.load(os.path.join(GOLDEN, "fits_small.npz"))


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library, with CUDA initialised (GPU tests only)."""
    from imageanalysis3_b200 import _lib
    _lib.init()
    return _lib


# tolerances of BASELINE.json's north_star for fitted parameters
TOL_CENTER_PX = 1e-3
TOL_REL = 1e-4


def assert_spots_close(got, want, what="", comparable=None):
    """got / want: (n, 11) rows [h, z, x, y, bk, sz, sx, sy, sin_t, sin_p, eps]; NaN rows must coincide
    on every row.  ``comparable`` (bool per row, from oracle.fit_oracle.comparable_mask) restricts the
    tolerance checks to rows whose REFERENCE fit is determined by its data: junk seeds on which
    MINPACK gives up at maxfev or ends on a (numerically) rank-deficient Jacobian, and slow crawls that
    scipy itself does not reproduce when 1 % of the voxels move by one float32 ulp, have no reproducible
    answer (tests/test_lm_core.py::test_ill_posed_reference_fits_are_not_reproducible_by_minpack_itself,
    tests/test_oracle_pinned.py::test_reference_sensitivity_probe_flags_slow_crawls)."""
    got = np.asarray([np.asarray(r, dtype=np.float64) for r in got])
    want = np.asarray([np.asarray(r, dtype=np.float64) for r in want])
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if got.size == 0:
        return
    gn, wn = np.isnan(got).any(1), np.isnan(want).any(1)
    assert np.array_equal(gn, wn), f"{what}: accept/reject (NaN rows) differ"
    if comparable is not None:
        comparable = np.asarray(comparable, dtype=bool)
        assert comparable.shape == gn.shape, (what, comparable.shape, gn.shape)
        gn, wn = gn | ~comparable, wn | ~comparable
    g, w = got[~gn], want[~wn]
    dc = np.abs(g[:, 1:4] - w[:, 1:4]).max() if len(g) else 0.0
    rel = lambda a, b: (np.abs(a - b) / np.maximum(np.abs(b), 1e-12)).max() if len(a) else 0.0
    assert dc <= TOL_CENTER_PX, f"{what}: centre differs by {dc} px"
    assert rel(g[:, 5:8], w[:, 5:8]) <= TOL_REL, f"{what}: sigma rel err {rel(g[:, 5:8], w[:, 5:8])}"
    assert rel(g[:, 0], w[:, 0]) <= TOL_REL, f"{what}: height rel err {rel(g[:, 0], w[:, 0])}"
    assert rel(g[:, 4], w[:, 4]) <= TOL_REL, f"{what}: background rel err {rel(g[:, 4], w[:, 4])}"
    assert rel(g[:, 10], w[:, 10]) <= TOL_REL, f"{what}: eps rel err {rel(g[:, 10], w[:, 10])}"
    assert np.abs(g[:, 8:10] - w[:, 8:10]).max() <= 1e-3 if len(g) else True, f"{what}: sin_t/sin_p differ"
