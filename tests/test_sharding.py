"""CPU test of the N>1 path: world_size-2 gloo run of the stream partition + timing reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cmix_b200.sharding import reduce_timing, stream_block


def test_stream_block_partitions_exactly():
    for total in (0, 1, 5, 8, 13, 26):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                blk = stream_block(total, world, r)
                assert blk == sorted(blk)
                seen += blk
            assert seen == list(range(total))
            sizes = [len(stream_block(total, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = stream_block(5, world, rank)
    t, b = reduce_timing(dist, torch.device("cpu"), 1.0 + rank, 100 * len(mine))
    q.put((rank, mine, t, b))
    dist.destroy_process_group()


def test_two_rank_gloo_run():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4]
    for _, _, t, b in res:
        assert t == 2.0          # max over ranks
        assert b == 500.0        # sum over ranks
