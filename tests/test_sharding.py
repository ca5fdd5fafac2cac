"""CPU: the N>1 path (chunk sharding + final gather) on gloo with world size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from faster_whisper_b200.sharding import shard_indices, transcribe_sharded


def test_shard_indices_partition():
    for n in (0, 1, 7, 16, 33):
        for world in (1, 2, 4, 8):
            for block in (1, 4, 16):
                seen = sorted(i for r in range(world) for i in shard_indices(n, r, world, block))
                assert seen == list(range(n))
    assert shard_indices(10, 1, 2, block=2) == [2, 3, 6, 7]
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chunks = [f"chunk{i}" for i in range(7)]
    res = transcribe_sharded(lambda cs: [(c, rank) for c in cs], chunks, rank, world, block=2)
    dist.barrier()
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the max-over-ranks timing reduction bench.py uses
    q.put((rank, res, float(t.item())))
    dist.destroy_process_group()


def test_gloo_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res, tmax in got:
        assert [c for c, _ in res] == [f"chunk{i}" for i in range(7)]  # every rank sees all chunks, in order
        assert [r for _, r in res] == [0, 0, 1, 1, 0, 0, 1]  # blocks of 2 dealt round-robin
        assert tmax == 2.0
