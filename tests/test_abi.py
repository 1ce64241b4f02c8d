"""CPU: the C-ABI library builds, loads, and exports every symbol include/ia3b200.h declares.
No compute call is made here (no GPU in the build container)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "ia3b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ia3_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from imageanalysis3_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names
    assert lib.ia3_version() >= 100


def test_no_cpu_fallback_without_gpu():
    """Without a visible GPU every compute entry point must fail loudly."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from imageanalysis3_b200 import _lib
    from imageanalysis3_b200.spot_tools import fitting
    with pytest.raises(_lib.IA3Error):
        fitting.get_seeds(np.zeros((8, 16, 16), dtype=np.uint16))
    with pytest.raises(_lib.IA3Error):
        fitting.fit_fov_image(np.zeros((8, 16, 16), dtype=np.uint16), '647', verbose=False)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "imageanalysis3_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


def test_argument_errors_mirror_reference():
    import numpy as np
    from imageanalysis3_b200.spot_tools import fitting
    with pytest.raises(TypeError):
        fitting.get_seeds([[1, 2], [3, 4]])
    with pytest.raises(IndexError):
        fitting.get_seeds(np.zeros((4, 8, 8), dtype=np.uint16), sel_center=[1, 2])
    assert fitting.remove_edge_points(np.zeros((10, 10, 10)), (np.array([2, 1, 8]), np.array([2, 5, 8]), np.array([8, 5, 9]))).tolist() == [True, False, False]
    c = fitting.select_sparse_centers(np.array([[0., 0, 0], [1, 1, 1], [20, 0, 0]]), distance_th=9)
    assert c.tolist() == [[0, 0, 0], [20, 0, 0]]
