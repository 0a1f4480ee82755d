"""CPU: scene sharding and the loss-gradient all_gather over gloo (world_size 2)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lcp_physics_b200.sharding import shard_range


def test_shard_range_covers_batch_exactly():
    for B in (0, 1, 7, 8, 4096, 32768 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from lcp_physics_b200.sharding import gather_loss_gradients, gather_scene_outputs, shard_inputs
    from lcp_physics_b200.scenes import make_scenes
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B = 7
        inp = make_scenes(B, 4, 4, fd=2, e=0, dtype=torch.float64, seed=3)
        mine = shard_inputs(inp, rank, world)
        lo, hi = shard_range(B, rank, world)
        assert mine[0].shape[0] == hi - lo and torch.equal(mine[2], inp[2][lo:hi])
        assert mine[4].numel() == 0           # the 1-D empty A passes through
        # a per-scene "output" and a per-rank parameter gradient
        local_out = mine[1] * 2.0
        full = gather_scene_outputs(local_out, B)
        ok_out = torch.equal(full, inp[1] * 2.0)
        local_grad = mine[1].sum(0)
        allg = gather_loss_gradients(local_grad)
        ok_grad = allg.shape == (world, inp[1].shape[1]) and torch.allclose(allg.sum(0), inp[1].sum(0))
        q.put((rank, bool(ok_out), bool(ok_grad)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res)
