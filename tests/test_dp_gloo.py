"""The N>1 path on CPU: world_size-2 gloo process group exercising the data-parallel plumbing (dp.py)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from medical_image_analysis_b200.dp import allreduce_param_grads, max_over_ranks, shard_batch
    try:
        dA = torch.full((3, 2), float(rank + 1))
        dD = torch.arange(3, dtype=torch.float32) * (rank + 1)
        db = None
        out = allreduce_param_grads([dA, dD, db], average=False)
        assert out[2] is None
        assert torch.equal(dA, torch.full((3, 2), 3.0)), dA
        assert torch.equal(dD, torch.arange(3, dtype=torch.float32) * 3)
        g = torch.ones(4, dtype=torch.bfloat16) * (rank + 1)
        allreduce_param_grads([g], average=True)
        assert torch.equal(g, torch.full((4,), 1.5, dtype=torch.bfloat16))
        assert max_over_ranks(10.0 + rank) == 11.0
        # the bucketed exchange bench.py runs at N > 1: 3 buckets of 4 floats (+1 short one), async, summed over ranks
        from medical_image_analysis_b200.dp import BucketedGradExchange
        ex = BucketedGradExchange(13, "cpu", bucket_bytes=16)
        assert len(ex.buckets) == 4
        ex.step([torch.full((3,), float(rank + 1)), None, torch.full((2,), 10.0 * (rank + 1))])
        ex.step([torch.full((3,), 2.0 * (rank + 1)), None, torch.full((2,), 1.0)], wait=True)     # waits for the first, completes the second
        assert torch.equal(ex.flat[:5], torch.tensor([6.0, 6.0, 6.0, 2.0, 2.0])), ex.flat
        assert ex.report()["bytes_per_step"] == 52 and ex.report()["world"] == 2
        # the layer cadence: one bucket per step, round robin (a gradient set leaves over 4 steps); only the bucket whose turn it
        # is gets summed over the ranks
        rr = BucketedGradExchange(13, "cpu", bucket_bytes=16, buckets_per_step=1)
        rr.flat.fill_(float(rank + 1))
        for k in range(3):
            rr.step([], wait=(k == 2))
        want = torch.tensor([3.0] * 12 + [float(rank + 1)])
        assert torch.equal(rr.flat, want), rr.flat
        rep = rr.report()
        assert rep["steps_per_gradient_set"] == 4 and rep["bytes_per_step"] == 13.0 and rr.bytes_sent == 48 and rr.cursor == 3
        rr.step([], wait=True)                                                       # the short tail bucket, then wrap around
        assert rr.flat[12].item() == 3.0 and rr.cursor == 0
        mine = shard_batch(7, rank, world)
        counts = [torch.zeros(1) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([float(len(mine))]))
        assert sum(int(c.item()) for c in counts) == 7
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_dp_plumbing_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_is_a_noop():
    from medical_image_analysis_b200.dp import allreduce_param_grads, max_over_ranks, shard_batch
    g = torch.ones(3)
    allreduce_param_grads([g, None])
    assert torch.equal(g, torch.ones(3)) and max_over_ranks(2.5) == 2.5
    assert list(shard_batch(5, 1, 2)) == [3, 4] and list(shard_batch(5, 0, 2)) == [0, 1, 2]
