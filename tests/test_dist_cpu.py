"""N > 1 host logic on CPU: world_size-2 gloo processes scatter a ragged batch, run a per-image
stand-in for the forward, gather, and the result must equal the single-process computation."""
import importlib
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dmod = importlib.import_module("deepcut-cnn_b200.dist")


def test_shard_counts_and_lpt():
    assert dmod.shard_counts(128, 8) == [16] * 8                   # BASELINE configs[3]
    assert dmod.shard_counts(5, 2) == [3, 2] and dmod.shard_counts(1, 4) == [1, 0, 0, 0]
    assert [(s.start, s.stop) for s in dmod.shard_slices(7, 3)] == [(0, 3), (3, 5), (5, 7)]
    # configs[4]: 8 images x scales {0.5, 1, 1.5} of 720p on 8 GPUs -> cost 0.25 : 1 : 2.25, perfectly balanced
    costs = [c for _ in range(8) for c in (0.25, 1.0, 2.25)]
    bins = dmod.lpt_assign(costs, 8)
    loads = [sum(costs[i] for i in b) for b in bins]
    assert sorted(i for b in bins for i in b) == list(range(24))
    assert max(loads) - min(loads) < 1e-9 and abs(loads[0] - 3.5) < 1e-9


def _per_image(x):            # stand-in for Net::Forward: any function that treats images independently
    return torch.stack([x.mean(dim=(1, 2, 3)), x.amax(dim=(1, 2, 3))], dim=1)


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cap = max(dmod.shard_counts(n_items, world))
    full = None
    if rank == 0:
        full = torch.arange(n_items * 3 * 4 * 5, dtype=torch.float32).reshape(n_items, 3, 4, 5) % 17
    local = torch.zeros(cap, 3, 4, 5)
    n = dmod.scatter_batch(dist, rank, world, full, local)
    assert n == dmod.shard_counts(n_items, world)[rank]
    out = torch.zeros(cap, 2)
    out[:n] = _per_image(local[:n])
    g = dmod.gather_batch(dist, rank, world, out, n_items)
    if rank == 0:
        q.put(g.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_forward_gather_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    for n_items in (5, 4):
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
        for p in procs:
            p.start()
        got = q.get(timeout=120)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        full = torch.arange(n_items * 3 * 4 * 5, dtype=torch.float32).reshape(n_items, 3, 4, 5) % 17
        assert np.array_equal(got, _per_image(full).numpy())
