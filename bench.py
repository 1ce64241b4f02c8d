#!/usr/bin/env python
"""bench.py -- spot-finding hot path (seed + firstfit + repeatfit) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = fit_fov_image over one synthetic FOV stack (BASELINE.json configs[1]: 50 x 2048 x
2048 uint16, 5000 planted spots, th_seed=300, max_num_seeds=None).  Steps are independent stacks, as in
the reference's mp.Pool over (round, channel) images (classes/field_of_view.py:1129); --inflight D of them
are in flight at once per GPU (D host threads, one CUDA stream per stack), so the copy of one stack
overlaps the seed kernels of the next and the long tail of a third one's fit sweeps.  Prints ONE JSON
line on rank 0.
  value      spots fitted per second, stack already resident in HBM when the timed region starts
  e2e        same metric through the public API (fit_fov_image on a pinned host stack): the H2D copy
             of the stack and the D2H reads of candidates / results are inside the timed region
  roofline   the seed stage's dominant kernel, one 61-tap exact Gaussian axis pass (k_gauss_strided<30>),
             against the measured HBM copy peak; roofline_fit = the fit stage against the FP pipes
  cpu_baseline  the oracle port (same scipy calls as the reference) on a bounded sample, 1 core
Multi-GPU: one process per GPU (torchrun), every rank processes its own stacks (weak scaling), no
data-path collective; the barrier / max-over-ranks reduction uses torch.distributed (NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# one hardware queue per in-flight stack (the default 8 makes streams share queues, and a stack's long
# fit tail then blocks unrelated stacks); must be set before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

SHAPE = (50, 2048, 2048)
N_PLANTED = 5000
TH_SEED = 300.0
FIT_KW = dict(th_seed=TH_SEED, max_num_seeds=None, verbose=False)
CPU_CROP = (50, 448, 448)      # bounded CPU sample: a crop of the same stack (same spot density)


def _traffic():
    """dram__bytes_read + dram__bytes_write of the 61-tap pass from the committed ncu capture (per launch)"""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
            t = json.load(fh)["k_gauss_61tap_pass"]
        return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except Exception:
        return None


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def _dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
def cpu_sample(im_crop):
    """The reference's CPU path (oracle port: same scipy/numpy calls) on one crop -> (#spots, seconds)."""
    from oracle import fit_oracle
    t0 = time.perf_counter()
    spots, seeds = fit_oracle.fit_fov_image_oracle(im_crop, th_seed=TH_SEED, max_num_seeds=None)
    return len(spots), time.perf_counter() - t0


def _cpu_worker(args):
    seed, crop = args
    from imageanalysis3_b200.synth import synth
    n = max(4, int(round(N_PLANTED * np.prod(crop) / np.prod(SHAPE))))
    im = synth(crop, n, seed)
    return cpu_sample(im)


def _reference_crop(steps, warmup, budget_s=150.0):
    """Crop of the C2 stack (same spot density) one worker handles per step, sized so that the whole
    --steps K --warmup W run takes about ``budget_s``: a 50x448x448 crop costs ~3 s with the pool busy."""
    per_step = budget_s / max(1, steps + warmup)
    edge = 448.0 * min(1.0, per_step / 3.0) ** 0.5
    edge = int(max(96, min(448, round(edge / 32.0) * 32)))
    return (SHAPE[0], edge, edge)


def run_reference(args):
    rank, world, local = _dist_setup(args.gpus)
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    ctx = mp.get_context("fork")
    crop = _reference_crop(args.steps, args.warmup)
    times, spots = [], []
    with ctx.Pool(procs) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(1000 + it * procs + i, crop) for i in range(procs)], chunksize=1)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
                spots.append(sum(r[0] for r in res))
    T = float(np.sum(times))
    value = float(np.sum(spots)) / T
    vox = float(np.prod(crop)) * procs * args.steps
    line = {
        "impl": "reference", "metric": "spots_fitted_per_s", "value": value, "unit": "spots/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "stacks_per_s": vox / float(np.prod(SHAPE)) / T,
        "config": {"workload": "C2: fit_fov_image on 50x2048x2048 uint16 FOV, 5000 planted spots, th_seed=300 "
                               "(reference CPU path timed on crops of that workload)"},
        "cpu_baseline": {"value": value, "unit": "spots/s", "cores": procs, "kind": "port",
                         "sample": f"{procs} process(es) x one {crop[0]}x{crop[1]}x{crop[2]} crop of the C2 stack per step "
                                   f"(same spot density), multiprocessing.Pool like classes/field_of_view.py:1129"},
        "e2e": {"value": value, "unit": "spots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    rank, world, local = _dist_setup(args.gpus)
    os.environ.setdefault("IA3_DEVICE", str(local))
    from imageanalysis3_b200 import _lib
    _lib.init(local)          # before torch touches the device: the library asks for blocking-sync waits
    import torch
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        backend = os.environ.get("IA3_BENCH_BACKEND", "nccl")
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    from imageanalysis3_b200 import _lib
    from imageanalysis3_b200.External import Fitting_v4
    from imageanalysis3_b200.spot_tools import fitting
    from imageanalysis3_b200.synth import synth_torch
    torch.cuda.set_device(local)
    _lib.init(local)
    dev = torch.device("cuda", local)

    from imageanalysis3_b200 import sharding
    D = max(1, args.inflight)
    n_stacks = max(2, min(D, 4))   # distinct stacks, cycled (each 419 MB > 126 MB L2)

    def run_steps(fn, first, count):
        return sum(sharding.map_stacks(fn, range(first, first + count), inflight=D))
    host, devt = [], []
    for i in range(n_stacks):
        d = synth_torch(SHAPE, N_PLANTED, 1 + rank * 16 + i, dev)
        h = torch.empty(SHAPE, dtype=torch.int16, pin_memory=True)
        h.copy_(d)
        host.append(h.numpy().view(np.uint16))
        devt.append(d)
    torch.cuda.synchronize()

    def step_resident(i):
        """hot path with the stack already in HBM: seed stage + host replay + firstfit + repeatfit"""
        d = devt[i % n_stacks]
        st = _lib.Stack(device_ptr=d.data_ptr(), shape=SHAPE, dtype=np.uint16)
        spots = fitting.fit_fov_image(host[i % n_stacks], '647', _stack=st, **FIT_KW)
        return len(spots)

    def step_e2e(i):
        spots = fitting.fit_fov_image(host[i % n_stacks], '647', **FIT_KW)
        return len(spots)

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: HBM-resident -------------------------------------------------------------------
    run_steps(step_resident, 0, max(args.warmup, D))
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launch_count()
    _lib.timer_start()
    t0 = time.perf_counter()
    cpu0 = time.process_time()
    n_spots = run_steps(step_resident, args.warmup, args.steps)
    ms_dev = _lib.timer_stop()
    wall = time.perf_counter() - t0
    cpu_ms_per_step = 1e3 * (time.process_time() - cpu0) / args.steps
    launches = _lib.launch_count() - l0
    barrier()
    clocks = sampler.stop()

    # stage timings (CUDA events on the library stream) for the roofline: separate short loop
    stage = {"gauss_fg": [], "gauss_bg": [], "rank": [], "compact": [], "seed_total": [], "n_cand": []}
    for i in range(max(3, args.steps)):
        d = devt[i % n_stacks]
        st = _lib.Stack(device_ptr=d.data_ptr(), shape=SHAPE, dtype=np.uint16)
        zxy, h, t = st.seed_candidates(fitting._gauss_half_kernel(0.75), fitting._gauss_half_kernel(7.5), 3, 0, 2.0, TH_SEED * 0.1)
        stage["gauss_fg"].append(t.ms_gauss_fg); stage["gauss_bg"].append(t.ms_gauss_bg); stage["rank"].append(t.ms_rank)
        stage["compact"].append(t.ms_compact); stage["seed_total"].append(t.ms_total); stage["n_cand"].append(len(zxy))
        st.close()
    # fit-stage device time
    d = devt[0]
    st = _lib.Stack(device_ptr=d.data_ptr(), shape=SHAPE, dtype=np.uint16)
    seeds = fitting.get_seeds(host[0], max_num_seeds=None, th_seed=TH_SEED, _stack=st)
    f = Fitting_v4.iter_fit_seed_points(host[0], seeds.T, _stack=st)
    f.firstfit(); first_ms = f._h.last_ms
    t1 = time.perf_counter(); f.repeatfit(); repeat_wall = time.perf_counter() - t1
    n_levels, n_iter = f._h.num_levels, f.n_iter
    del f, st

    # ---- e2e: public API, host buffers -----------------------------------------------------------
    run_steps(step_e2e, 0, max(min(args.warmup, 3), D))
    barrier()
    c0 = dict(_lib.COPIED)
    _lib.timer_start()
    n_e2e = run_steps(step_e2e, args.warmup, args.steps)
    ms_e2e = _lib.timer_stop()
    barrier()
    c1 = dict(_lib.COPIED)
    # latency of one stack with nothing else in flight (same public call)
    _lib.timer_start()
    step_e2e(0)
    ms_latency = _lib.timer_stop()

    t_val = torch.tensor([ms_dev, ms_e2e, float(n_spots), float(n_e2e), float(launches)], dtype=torch.float64,
                         device=dev if (not use_dist or dist.get_backend() == "nccl") else "cpu")
    if use_dist and os.environ.get("IA3_BENCH_VERBOSE"):
        print(f"rank {rank}: {ms_dev / args.steps:.2f} ms/step resident, {ms_e2e / args.steps:.2f} ms/step e2e", file=sys.stderr, flush=True)
    if use_dist:
        tmax = t_val.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t_val.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_dev, ms_e2e = float(tmax[0]), float(tmax[1])
        n_spots, n_e2e, launches = float(tsum[2]), float(tsum[3]), float(tsum[4])
    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    vox = float(np.prod(SHAPE))
    peak, peak_src = _peaks()
    bg_launch_ms = float(np.median(stage["gauss_bg"])) / 3.0
    alg_bytes = 4.0 * vox                                  # one axis pass: read u16 + write u16 per voxel
    achieved = alg_bytes / (bg_launch_ms * 1e-3) / 1e9
    seed_ms_med = float(np.median(stage["seed_total"]))
    stage_bytes = 2.0 * vox + 16.0 * float(np.median(stage["n_cand"]))
    fp64_inst = 32.0 * vox                                  # 1 DMUL + 30 DFMA + 1 DADD (offset + guard) per voxel
    fit_flops_per_spot = 1.67e6                             # model / Jacobian / normal equations per spot
    line = {
        "metric": "spots_fitted_per_s", "value": n_spots / (ms_dev * 1e-3), "unit": "spots/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "stacks_per_s": world * args.steps / (ms_dev * 1e-3),
        "config": {"workload": "C2: fit_fov_image(th_seed=300, max_num_seeds=None) on one 50x2048x2048 uint16 FOV per step per GPU, "
                               "5000 planted spots; seed stage + Fitting_v4 firstfit + repeatfit",
                   "l2": f"inputs larger than L2 (419 MB per stack, {n_stacks} stacks cycled)",
                   "inflight": D, "latency_ms_one_stack_alone": ms_latency,
                   "host_cpu_ms_per_step": cpu_ms_per_step, "host_cores": os.cpu_count(),
                   "spots_per_stack": n_spots / (world * args.steps), "fit_levels": n_levels, "repeat_sweeps": n_iter},
        "clocks": clocks,
        "e2e": {"value": n_e2e / (ms_e2e * 1e-3), "unit": "spots/s",
                "h2d_bytes_per_step": (c1["h2d"] - c0["h2d"]) / args.steps, "d2h_bytes_per_step": (c1["d2h"] - c0["d2h"]) / args.steps,
                "ms_per_step": ms_e2e / args.steps, "stacks_per_s": world * args.steps / (ms_e2e * 1e-3)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": _traffic(),
                     "kernel": "k_gauss_short<30,50> / k_gauss_strided<30> / k_gauss_contig<30> (one 61-tap bit-exact axis pass; 3 launches per stack, mean)",
                     "launch_ms": bg_launch_ms, "peak_source": peak_src,
                     "note": "exact uint16 semantics make this pass FP64-pipe / issue bound, not HBM bound: 32 FP64 + ~60 integer instructions per voxel; "
                             "by time the dominant kernel of a step is k_fit<double> (latency bound, see roofline_fit and profiles/r01_ncu_launch_list_final.csv)",
                     "fp64_inst_per_s": fp64_inst / (bg_launch_ms * 1e-3),
                     "fp64_pipe_frac": fp64_inst / (bg_launch_ms * 1e-3) / (148 * 64 * 1.965e9),
                     "seed_stage": {"ms": seed_ms_med, "algorithmic_bytes": stage_bytes,
                                    "achieved_GBps": stage_bytes / (seed_ms_med * 1e-3) / 1e9,
                                    "frac": stage_bytes / (seed_ms_med * 1e-3) / 1e9 / peak,
                                    "ms_gauss_fg": float(np.median(stage["gauss_fg"])), "ms_gauss_bg": float(np.median(stage["gauss_bg"])),
                                    "ms_rank": float(np.median(stage["rank"])), "ms_compact": float(np.median(stage["compact"]))},
                     "fit_stage": {"firstfit_ms": first_ms, "repeatfit_wall_ms": repeat_wall * 1e3}},
        "roofline_fit": {"bound": "fp64/fp32 pipes (latency bound: one warp per spot runs MINPACK's serial iterations)",
                         "achieved": fit_flops_per_spot * n_spots / (ms_dev * 1e-3) / 1e12, "unit": "TFLOP/s",
                         "peak_fp32": 148 * 128 * 2 * 1.965e9 / 1e12, "peak_fp64": 148 * 64 * 2 * 1.965e9 / 1e12,
                         "algorithmic_flops_per_spot": fit_flops_per_spot},
        "wall_ms_per_step": 1e3 * wall / args.steps,
    }
    if not args.no_cpu:
        crop = np.ascontiguousarray(host[0][:CPU_CROP[0], 800:800 + CPU_CROP[1], 800:800 + CPU_CROP[2]])
        n_cpu, t_cpu = cpu_sample(crop)
        line["cpu_baseline"] = {"value": n_cpu / t_cpu, "unit": "spots/s", "cores": 1, "kind": "port",
                                "sample": f"one {CPU_CROP[0]}x{CPU_CROP[1]}x{CPU_CROP[2]} crop of the same stack: {n_cpu} spots in {t_cpu:.1f} s "
                                          "(oracle port = the reference's scipy/numpy calls, single thread as in the reference)"}
    print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--inflight", type=int, default=64, help="stacks in flight per GPU (host threads / CUDA streams)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
