"""CPU, world_size 2, gloo: the multi-process host logic of the N>1 path (batch sharding, max-over-ranks timing,
output gathering).  No CUDA compute."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stp3_b200 import parallel


def _worker(rank, world_size, port, global_batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        assert parallel.world() == (rank, world_size)
        start, count = parallel.shard_batch(global_batch, rank, world_size)
        # each rank "computes" logits for its own samples: value = global sample index
        logits = torch.arange(start, start + count, dtype=torch.float32).view(count, 1, 1).expand(count, 2, 3).contiguous()
        out = parallel.gather_outputs({"segmentation": logits, "hdmap": None}, global_batch)
        slowest = parallel.max_over_ranks(10.0 + rank)
        parallel.barrier()
        q.put((rank, start, count, out["segmentation"][:, 0, 0].tolist(), out["hdmap"], slowest))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("global_batch", [8, 5])
def test_two_rank_sharding_gather_and_max(global_batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + global_batch
    procs = [ctx.Process(target=_worker, args=(r, 2, port, global_batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, start, count, gathered, hdmap, slowest in res:
        covered += list(range(start, start + count))
        assert gathered == [float(i) for i in range(global_batch)]      # global-batch order on every rank
        assert hdmap is None and slowest == 11.0                         # max over ranks
    assert covered == list(range(global_batch))                          # disjoint, complete


def test_shard_batch_properties():
    for gb in range(0, 40):
        for ws in (1, 2, 3, 4, 8):
            spans = [parallel.shard_batch(gb, r, ws) for r in range(ws)]
            assert sum(c for _, c in spans) == gb
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        parallel.shard_batch(4, 2, 2)


def _gather_worker(rank, world_size, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        f0, fc = parallel.shard_batch(n_frames, rank, world_size)
        local = torch.arange(f0, f0 + fc, dtype=torch.float32).view(fc, 1, 1, 1).expand(fc, 2, 2, 3).contiguous()
        allf = parallel.all_gather_frames(local, n_frames)
        q.put((rank, allf[:, 0, 0, 0].tolist(), tuple(allf.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size,n_frames", [(2, 1), (2, 3), (2, 6), (4, 3)])
def test_all_gather_frames_uneven_and_even(world_size, n_frames):
    """Frame-sharded mode: flat frames b*S+t split over the ranks (3 frames on 2 ranks -> 2 + 1; on 4 ranks -> 1 + 1 + 1
    + 0: ranks without a frame still take part in the exchange) and gathered back in order."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29620 + 10 * world_size + n_frames
    procs = [ctx.Process(target=_gather_worker, args=(r, world_size, port, n_frames, q)) for r in range(world_size)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, order, shape in res:
        assert order == [float(i) for i in range(n_frames)] and shape == (n_frames, 2, 2, 3)
