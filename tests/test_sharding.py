"""Multi-GPU path is request sharding with no data-path collective.  The host logic is
checked with a world_size-2 gloo group on CPU."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pocket_tts_b200.tts_model import shard_requests


def test_shards_partition():
    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                seen += list(shard_requests(n, world, r))
            assert seen == list(range(n))
            sizes = [len(shard_requests(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_req, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_requests(n_req, world, rank)
    frames = torch.tensor([len(mine) * 125], dtype=torch.int64)  # each request emits 125 frames
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)        # per-rank device time
    dist.all_reduce(frames, op=dist.ReduceOp.SUM)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                    # bench.py contract: max over ranks
    if rank == 0:
        q.put((int(frames), float(ms)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_aggregate_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 129, q)) for r in range(2)]
    for p in procs:
        p.start()
    frames, ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert frames == 129 * 125 and ms == 11.0
