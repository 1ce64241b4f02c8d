"""GPU: pinned host -> device copy bandwidth of this box (the floor under bench.py's e2e number).
    python tools/h2d_peak.py
One 50x2048x2048 uint16 stack (419 MB) per copy; 1, 2 and 4 copies in flight on separate streams.
"""
import torch

N = 50 * 2048 * 2048
dev = torch.device("cuda", 0)
for k in (1, 2, 4):
    hs = [torch.empty(N, dtype=torch.int16, pin_memory=True) for _ in range(k)]
    ds = [torch.empty(N, dtype=torch.int16, device=dev) for _ in range(k)]
    ss = [torch.cuda.Stream() for _ in range(k)]
    for h in hs:
        h.zero_()
    best = 1e9
    for rep in range(6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in ss:
            s.wait_stream(torch.cuda.current_stream())
        for h, d, s in zip(hs, ds, ss):
            with torch.cuda.stream(s):
                d.copy_(h, non_blocking=True)
        for s in ss:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1))
    print(f"{k} copies in flight: {k * N * 2 / best / 1e6:.1f} GB/s  ({best / k:.2f} ms per 419 MB stack)")
