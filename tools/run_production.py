"""GPU: the production call of the reference's pipeline (classes/field_of_view.py:1001-1008,
classes/batch_functions.py:269-285): fit_fov_image(im, ch, th_seed=600, max_num_seeds=4000,
min_dynamic_seeds=50, remove_hot_pixel=True, normalize_local=True) on 30 x 2048 x 2048 stacks
(BASELINE configs C3 / C5), `inflight` stacks at a time.
    python tools/run_production.py [n_stacks] [inflight]
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from imageanalysis3_b200 import _lib, sharding
from imageanalysis3_b200.spot_tools import fitting
from imageanalysis3_b200.synth import synth_torch

n_stacks = int(sys.argv[1]) if len(sys.argv) > 1 else 48
inflight = int(sys.argv[2]) if len(sys.argv) > 2 else 32
SHAPE = (30, 2048, 2048)
_lib.init(0)
dev = torch.device("cuda", 0)
hosts = []
for i in range(4):
    d = synth_torch(SHAPE, 2000, 1000 + i, dev)
    h = torch.empty(SHAPE, dtype=torch.int16, pin_memory=True)
    h.copy_(d)
    hosts.append(h.numpy().view(np.uint16))
    del d
KW = dict(th_seed=600, max_num_seeds=4000, min_dynamic_seeds=50, remove_hot_pixel=True, normalize_local=True, verbose=False)
run = lambda i: fitting.fit_fov_image(hosts[i % 4], '647', **KW)
sharding.map_stacks(run, range(inflight), inflight)            # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
res = sharding.map_stacks(run, range(n_stacks), inflight)
dt = time.perf_counter() - t0
n = sum(len(r) for r in res)
print(f"production kwargs, {n_stacks} stacks of {SHAPE}, {inflight} in flight: {1e3 * dt / n_stacks:.1f} ms per stack, "
      f"{n / dt:.0f} spots/s, {n / n_stacks:.0f} spots per stack, height/background median {np.median(res[0][:, 0]):.2f}")
t0 = time.perf_counter()
one = run(0)
print(f"one stack alone: {1e3 * (time.perf_counter() - t0):.1f} ms, {len(one)} spots")
