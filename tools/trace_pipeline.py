"""GPU: where the wall time of D concurrent fit_fov_image-like steps goes (host-side phase trace).
    python tools/trace_pipeline.py [D] [K]
"""
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, ".")
from imageanalysis3_b200 import _lib
from imageanalysis3_b200.External import Fitting_v4
from imageanalysis3_b200.spot_tools import fitting
from imageanalysis3_b200.synth import synth_torch

D = int(sys.argv[1]) if len(sys.argv) > 1 else 4
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
MODE = sys.argv[3] if len(sys.argv) > 3 else "both"      # both | seed | fit | first | e2e | upload
QUIET = len(sys.argv) > 4
SHAPE = (50, 2048, 2048)
_lib.init(0)
if os.environ.get("TRACE_MAXFEV"):         # experiment only (not reference semantics): how much of the pipeline's
    _orig_cfg = _lib.make_fit_cfg          # latency is the few junk seeds that run MINPACK to maxfev = 1000
    def _cfg(*a, **kw):
        kw["maxfev"] = int(os.environ["TRACE_MAXFEV"])
        return _orig_cfg(*a, **kw)
    _lib.make_fit_cfg = _cfg
dev = torch.device("cuda", 0)
stacks = []
for i in range(min(D, 4)):
    d = synth_torch(SHAPE, 5000, 1 + i, dev)
    h = torch.empty(SHAPE, dtype=torch.int16, pin_memory=True)
    h.copy_(d)
    stacks.append((d, h.numpy().view(np.uint16)))
torch.cuda.synchronize()
T0 = time.perf_counter()
trace = []


def step(i):
    d, host = stacks[i % len(stacks)]
    ev = []
    def mark(name, t0):
        ev.append((name, 1e3 * (t0 - T0), 1e3 * (time.perf_counter() - T0)))
    t = time.perf_counter()
    if MODE == "e2e":                      # the public call on a pinned host stack (bench.py's e2e leg)
        spots = fitting.fit_fov_image(host, '647', th_seed=300.0, max_num_seeds=None, verbose=False)
        mark("e2e", t)
        trace.append((i, threading.get_ident() % 1000, ev))
        return len(spots)
    if MODE == "upload":                   # host -> device copy of the stack alone
        _lib.Stack(host).close()
        mark("upload", t)
        trace.append((i, threading.get_ident() % 1000, ev))
        return 0
    st = _lib.Stack(device_ptr=d.data_ptr(), shape=SHAPE, dtype=np.uint16)
    if MODE in ("fit", "first"):
        seeds = SEEDS[i % len(stacks)]
    else:
        seeds = fitting.get_seeds(host, max_num_seeds=None, th_seed=300.0, _stack=st)
        st.trim(1)
    mark("seed", t); t = time.perf_counter()
    if MODE == "seed":
        trace.append((i, threading.get_ident() % 1000, ev))
        return len(seeds)
    f = Fitting_v4.iter_fit_seed_points(host, seeds.T, _stack=st)
    f.firstfit()
    mark("first", t); t = time.perf_counter()
    if MODE == "first":
        trace.append((i, threading.get_ident() % 1000, ev))
        return len(seeds)
    f.repeatfit()
    mark(f"repeat x{f.n_iter}", t)
    trace.append((i, threading.get_ident() % 1000, ev))
    return len(seeds)


SEEDS = [fitting.get_seeds(h, max_num_seeds=None, th_seed=300.0) for _, h in stacks]
pool = ThreadPoolExecutor(D)
list(pool.map(step, range(D)))       # warm-up
trace.clear()
_lib.debug_stats(reset=True)
CPU0 = time.process_time()
T0 = time.perf_counter()
list(pool.map(step, range(K)))
total = 1e3 * (time.perf_counter() - T0)
for i, tid, ev in sorted(trace):
    if QUIET:
        break
    print(f"step {i} thr {tid:3d}: " + "  ".join(f"{n} {a:7.1f}-{b:7.1f}" for n, a, b in ev))
print(_lib.debug_stats())
print(f"process CPU {1e3 * (time.process_time() - CPU0) / K:.2f} ms per step")
print(f"mode={MODE} D={D} K={K} total {total:.1f} ms -> {total / K:.1f} ms/step")
