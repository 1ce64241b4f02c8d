"""GPU: cProfile of one fit_fov_image call (host-side cost per stack; the GIL serialises it across stacks)."""
import cProfile
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from imageanalysis3_b200 import _lib
from imageanalysis3_b200.spot_tools import fitting
from imageanalysis3_b200.synth import synth_torch

SHAPE = (50, 2048, 2048)
_lib.init(0)
d = synth_torch(SHAPE, 5000, 1, torch.device("cuda", 0))
h = torch.empty(SHAPE, dtype=torch.int16, pin_memory=True)
h.copy_(d)
host = h.numpy().view(np.uint16)
for _ in range(2):
    fitting.fit_fov_image(host, '647', th_seed=300., max_num_seeds=None, verbose=False)
pr = cProfile.Profile()
pr.enable()
fitting.fit_fov_image(host, '647', th_seed=300., max_num_seeds=None, verbose=False)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
