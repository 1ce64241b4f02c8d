"""GPU: stage timings of the seed stage on one synthetic stack (also the ncu target for the seed kernels).
    python tools/profile_seed.py [Z X Y] [reps]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from imageanalysis3_b200 import _lib
from imageanalysis3_b200.spot_tools import fitting

shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (50, 2048, 2048)
reps = int(sys.argv[4]) if len(sys.argv) >= 5 else 3
_lib.init(0)
g = torch.Generator(device="cuda").manual_seed(0)
d = (300 + 20 * torch.randn(shape, device="cuda", generator=g)).clamp_(0, 65535).to(torch.int16)   # uint16 bit pattern
torch.cuda.synchronize()
for r in range(reps):
    st = _lib.Stack(device_ptr=d.data_ptr(), shape=shape, dtype=np.uint16)
    zxy, h, t = st.seed_candidates(fitting._gauss_half_kernel(0.75), fitting._gauss_half_kernel(7.5), 3, 0, 2.0, 30.0)
    print(f"rep {r}: fg {t.ms_gauss_fg:.3f} bg {t.ms_gauss_bg:.3f} rank {t.ms_rank:.3f} compact {t.ms_compact:.3f} "
          f"total {t.ms_total:.3f} ms, {len(zxy)} candidates")
    st.close()
