"""Sharding of independent image stacks over the GPUs of one box, and over the stacks in flight on
each GPU.

The reference parallelises this path with a process pool over (hyb round, channel) images
(``mp.Pool(num_threads).starmap(batch_process_image_to_spots, ..., chunksize=1)``,
classes/field_of_view.py:1129-1142; every image is independent).  Here:

* across GPUs: one process per GPU (``torchrun``), rank r takes stacks r, r + N, r + 2N, ...  There is
  no collective on the data path; the per-stack spot tables (a few hundred kB) are gathered on the
  host with ``torch.distributed.gather_object`` (NCCL is used by bench.py only for its barrier);
* on one GPU: ``inflight`` host threads, each driving one stack on its own CUDA stream, so that the
  host->device copy of one stack overlaps the seed kernels of the next and the long serial tail of a
  third one's fit sweeps (a handful of junk seeds run MINPACK to maxfev) -- the GPU would otherwise
  idle for ~95 % of a stack's latency.
"""
from concurrent.futures import ThreadPoolExecutor


def assign_stacks(n_stacks, world_size, rank):
    """indices of the stacks rank ``rank`` of ``world_size`` processes (round robin)"""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size: {rank}/{world_size}")
    return list(range(rank, n_stacks, world_size))


def map_stacks(fn, items, inflight=16):
    """``[fn(item) for item in items]`` with ``inflight`` items in flight (threads; the C ABI releases
    the GIL and gives every stack its own stream).  Results keep the order of ``items``; the first
    exception is re-raised."""
    items = list(items)
    if inflight <= 1 or len(items) <= 1:
        return [fn(it) for it in items]
    with ThreadPoolExecutor(max_workers=min(inflight, len(items))) as pool:
        return list(pool.map(fn, items))


def gather_tables(local, dst=0):
    """``local``: {stack index: ndarray} of this rank.  Returns the merged dict on rank ``dst`` (None
    elsewhere).  Without an initialised process group (single GPU) it returns ``local``."""
    try:
        import torch.distributed as dist
    except Exception:                                       # pragma: no cover
        dist = None
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return dict(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    out = [None] * world if rank == dst else None
    dist.gather_object(dict(local), out, dst=dst)
    if rank != dst:
        return None
    merged = {}
    for part in out:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError(f"stack {k} was processed by two ranks")
            merged[k] = v
    return merged


def process_stacks(fn, n_stacks, load, inflight=16, dst=0):
    """Run ``fn(load(i))`` for every stack i of this rank's shard and gather {i: result} on ``dst``.
    ``load(i)`` produces the i-th stack (e.g. reads a DAX/HDF5 image); ``fn`` is e.g.
    ``lambda im: fit_fov_image(im, '647', ...)``."""
    try:
        import torch.distributed as dist
        ok = dist.is_available() and dist.is_initialized()
    except Exception:                                       # pragma: no cover
        ok = False
    world, rank = (dist.get_world_size(), dist.get_rank()) if ok else (1, 0)
    mine = assign_stacks(n_stacks, world, rank)
    res = map_stacks(lambda i: fn(load(i)), mine, inflight)
    return gather_tables(dict(zip(mine, res)), dst=dst)
