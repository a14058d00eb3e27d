"""N>1 host logic on CPU: world_size-2 gloo processes shard segments and re-assemble results in order; thread pool of
replicas; the reference's long-segment cutting rule."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sonicscribe_b200.pool import ReplicaPool, cut_long_segments, gather_in_order, shard_indices


def _worker(rank, world, port, n_items, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_items, world, rank)
    local = [f"seg{idx}-by{rank}" for idx in mine]          # stand-in for ASRModel.transcribe_batch on this rank's GPU
    full = gather_in_order(local, n_items, world, rank)
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t)                                      # the bench's max/sum-over-ranks plumbing
    if rank == 0:
        ret["full"] = full
        ret["total"] = float(t[0])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [1, 7, 180])
def test_two_rank_sharding_restores_order(n_items):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 1000) + n_items % 7
    mp.spawn(_worker, args=(world, port, n_items, ret), nprocs=world, join=True)
    full = ret["full"]
    assert len(full) == n_items and ret["total"] == n_items
    for i, s in enumerate(full):
        assert s == f"seg{i}-by{i % world}"


def test_cut_long_segments_matches_reference_rule():
    sr = 16000
    one_hour = 3600 * sr
    cuts = cut_long_segments(0, one_hour, sr, 20.0)
    assert len(cuts) == 180 and all(e - s == 320000 for s, e in cuts)        # BASELINE config 4: 180 x 20 s
    assert cut_long_segments(0, 10 * sr, sr, 20.0) == [(0, 10 * sr)]
    cuts = cut_long_segments(100, 100 + 41 * sr, sr, 20.0)
    assert cuts == [(100, 100 + 20 * sr), (100 + 20 * sr, 100 + 40 * sr), (100 + 40 * sr, 100 + 41 * sr)]
    cuts = cut_long_segments(0, 40 * sr + 800, sr, 20.0)                     # 50 ms tail is dropped (<= 0.1 s)
    assert len(cuts) == 2


def test_replica_pool_orders_results():
    class Fake:
        def __init__(self, dev):
            self.dev = dev

        def transcribe_batch(self, segs, **kw):
            return [f"{s}@{self.dev}" for s in segs]

    pool = ReplicaPool(lambda d: Fake(d), n_gpus=3, batch=4)
    out = pool.transcribe_segments(list(range(23)))
    assert [int(o.split("@")[0]) for o in out] == list(range(23))
    assert {o.split("@")[1] for o in out} == {"0", "1", "2"}
    pool.close()
