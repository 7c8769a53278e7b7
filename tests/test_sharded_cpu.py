"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo processes, ragged shards.
The per-rank compute is a stand-in (the CUDA path needs a GPU); what is checked is the sharding
arithmetic and the gather order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from egotap_b200.sharded import ShardedLifter, gather_job_poses, gather_poses, shard_bounds


def test_shard_bounds_cover_exactly():
    for total in (0, 1, 5, 16, 255, 256, 1024, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class _FakeNet:
    num_joints = 16

    def predict_pose(self, x):  # frame-wise function of the input, like the real net in eval mode
        s = x.flatten(1).sum(1)
        return torch.stack([s + j for j in range(48)], 1).view(-1, 16, 3)


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        frames = torch.rand(total, 6, 4, 4, generator=g)
        want = _FakeNet().predict_pose(frames)
        got = ShardedLifter(_FakeNet()).predict_pose(frames)
        lo, hi = shard_bounds(total, world, rank)
        got2 = ShardedLifter(_FakeNet()).predict_pose(frames[lo:hi], total=total, presharded=True)
        # a sharded JOB: K batches per rank with no collective in between, one gather at its end (bench.py --gpus N)
        K, Bl = 3, 2
        mine = torch.stack([_FakeNet().predict_pose(frames[:Bl] + 10 * rank + k) for k in range(K)])
        job = gather_job_poses(mine)
        job_ok = tuple(job.shape) == (world, K, Bl, 16, 3) and all(
            torch.equal(job[r, k], _FakeNet().predict_pose(frames[:Bl] + 10 * r + k)) for r in range(world) for k in range(K))
        q.put((rank, bool(torch.equal(got, want)), bool(torch.equal(got2, want)) and job_ok))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,total", [(2, 8), (2, 7), (3, 4)])
def test_gather_order_over_gloo(world, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] and r[2] for r in res), res
