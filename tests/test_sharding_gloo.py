"""N>1 host logic on CPU: two ranks over gloo shard a frame batch with dmc_shard_frames, agree on coverage with an
all_gather, and reduce their timings with MAX -- the same plumbing bench.py uses over NCCL (no data-path collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from depthmapcompression_b200 import shard_frames
    begin, count = shard_frames(n_frames, rank, world)
    mine = torch.zeros(n_frames, dtype=torch.int32); mine[begin:begin + count] = 1       # frames this rank would filter
    owners = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(owners, mine)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)                                  # per-rank elapsed ms
    dist.barrier(); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cover = torch.stack(owners).sum(0)
    q.put((rank, begin, count, bool((cover == 1).all()), float(t[0])))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [1000, 1001, 3])
def test_two_rank_frame_sharding(n_frames):
    import __graft_entry__ as g
    g.build()
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs: p.join(timeout=60)
    assert [r[0] for r in res] == [0, 1]
    assert res[0][1] == 0 and res[0][1] + res[0][2] == res[1][1] and res[1][1] + res[1][2] == n_frames
    assert abs(res[0][2] - res[1][2]) <= 1
    assert all(r[3] for r in res), "every frame owned by exactly one rank"
    assert all(r[4] == 11.0 for r in res), "timing is the MAX over ranks"
