"""GPU: where one fit_fov_image call spends HOST time -- cProfile of the Python side plus, for every libia3b200 call,
wall time and the calling thread's CPU time (a blocked wait costs wall but no CPU; CPU is what limits several ranks
sharing a box, Python-side time is what the GIL serialises across a rank's stacks in flight)."""
import cProfile
import pstats
import sys
import time
from collections import defaultdict

import numpy as np
import torch

sys.path.insert(0, ".")
from imageanalysis3_b200 import _lib
from imageanalysis3_b200.spot_tools import fitting
from imageanalysis3_b200.synth import synth_torch

SHAPE = (50, 2048, 2048)
_lib.init(0)
lib = _lib.load()
acc = defaultdict(lambda: [0, 0.0, 0.0])


def wrap(name):
    fn = getattr(lib, name)

    def w(*a):
        t0, c0 = time.perf_counter(), time.thread_time()
        r = fn(*a)
        e = acc[name]
        e[0] += 1
        e[1] += time.perf_counter() - t0
        e[2] += time.thread_time() - c0
        return r
    setattr(lib, name, w)


for n in _lib.EXPORTS:
    if n not in ("ia3_last_error",):
        wrap(n)
d = synth_torch(SHAPE, 5000, 1, torch.device("cuda", 0))
h = torch.empty(SHAPE, dtype=torch.int16, pin_memory=True)
h.copy_(d)
host = h.numpy().view(np.uint16)
for _ in range(3):
    fitting.fit_fov_image(host, '647', th_seed=300., max_num_seeds=None, verbose=False)
acc.clear()
N = 5
pr = cProfile.Profile()
t0, c0, p0 = time.perf_counter(), time.thread_time(), time.process_time()
pr.enable()
for _ in range(N):
    fitting.fit_fov_image(host, '647', th_seed=300., max_num_seeds=None, verbose=False)
pr.disable()
wall, cpu, pcpu = time.perf_counter() - t0, time.thread_time() - c0, time.process_time() - p0
print(f"per stack: wall {1e3*wall/N:.2f} ms, calling thread CPU {1e3*cpu/N:.2f} ms, process CPU {1e3*pcpu/N:.2f} ms")
tot_w = sum(v[1] for v in acc.values())
tot_c = sum(v[2] for v in acc.values())
print(f"inside libia3b200: wall {1e3*tot_w/N:.2f} ms, CPU {1e3*tot_c/N:.2f} ms  ->  Python side: CPU {1e3*(cpu-tot_c)/N:.2f} ms per stack")
print(f"{'entry point':28s} {'calls':>6s} {'wall ms':>9s} {'cpu ms':>9s}   (per stack)")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][2]):
    print(f"{k:28s} {v[0]/N:6.1f} {1e3*v[1]/N:9.3f} {1e3*v[2]/N:9.3f}")
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
