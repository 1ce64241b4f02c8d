"""CPU: the multi-GPU sharding logic (one process per GPU, stacks sharded round robin, host gather)
exercised with two gloo ranks; the per-stack work is a stand-in (no GPU here)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from imageanalysis3_b200 import sharding


def test_assign_stacks_partitions():
    for n in (0, 1, 7, 100):
        for world in (1, 2, 3, 8):
            parts = [sharding.assign_stacks(n, world, r) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        sharding.assign_stacks(4, 2, 2)


def test_map_stacks_keeps_order_and_raises():
    out = sharding.map_stacks(lambda i: i * i, range(20), inflight=4)
    assert out == [i * i for i in range(20)]
    assert sharding.map_stacks(lambda i: i + 1, [], inflight=4) == []

    def boom(i):
        if i == 3:
            raise KeyError("stack 3")
        return i
    with pytest.raises(KeyError):
        sharding.map_stacks(boom, range(6), inflight=3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_stacks, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def load(i):
            return np.full((2, 3), i, dtype=np.float32)

        def fn(im):                                   # stand-in for fit_fov_image: (M, 11) table
            m = int(im[0, 0]) % 4
            return np.full((m, 11), im[0, 0], dtype=np.float32)
        merged = sharding.process_stacks(fn, n_stacks, load, inflight=3, dst=0)
        if rank == 0:
            q.put({k: (v.shape, float(v[0, 0]) if len(v) else None) for k, v in merged.items()})
        else:
            assert merged is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_and_gathers():
    world, n_stacks = 2, 9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_stacks, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(got) == list(range(n_stacks))
    for i, (shape, v) in got.items():
        assert shape == (i % 4, 11)
        assert v is None or v == float(i)


def test_fit_fov_image_gates_are_released_on_errors():
    """fit_fov_image holds an admission slot (stacks between upload and the end of firstfit) and, while
    seeding, a seed-stage slot: a call that fails must give them back, or the 25th failing caller of a
    `map_stacks` run would wait forever.  No GPU involved: the reference's TypeError for a non-array
    image is raised before any device work."""
    import inspect
    from imageanalysis3_b200.spot_tools import fitting
    n_slots = fitting._ADMIT_GATE._initial_value

    def bad(_):
        with pytest.raises(TypeError):
            fitting.fit_fov_image("not an image", '647', verbose=False)
        return 1
    assert sum(sharding.map_stacks(bad, range(3 * n_slots + 5), inflight=8)) == 3 * n_slots + 5
    for gate in (fitting._ADMIT_GATE, fitting._SEED_GATE):
        assert gate._value == gate._initial_value
    # the wrapper keeps the reference's parameters and defaults visible
    params = inspect.signature(fitting.fit_fov_image).parameters
    assert list(params)[:4] == ["im", "channel", "seeds", "seed_mask"] and params["max_num_seeds"].default == 500
    assert "_front_done" not in params
