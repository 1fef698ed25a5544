"""N > 1 host logic on CPU: two gloo ranks run the bucketed gradient reducer over a ParamStore and must end
with the mean of the per-rank gradients in every bucket (gradients = mean over ranks), with
contiguous buckets covering the used arena exactly once.  Also checks the batch sharding helper."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q, model="retina"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import zsg_b200  # noqa: F401
    from zsg_b200 import ddp, engine, spec
    store = engine.ParamStore(torch.device("cpu"), model)
    g = torch.Generator().manual_seed(100 + rank)
    store.grad_arena.copy_(torch.randn(store.used, generator=g))
    mine = store.grad_arena.clone()
    red = ddp.GradReducer(store, min_bucket_elems=2 << 20)
    # replay the engine's bucket marks: stage ranges in arena order
    names = [n for n, _, _ in reversed(spec.trainable_specs(model))]
    marks, lo = [], 0
    for i, n in enumerate(names):
        hi = store.offsets[n] + engine._align(store.numel(n))
        if i % 9 == 8 or i == len(names) - 1:
            marks.append((lo, hi))
            lo = hi
    for lo_, hi_ in marks:
        red.on_bucket(lo_, hi_)
    red.finish()
    other = torch.randn(store.used, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    expect = (mine + other) / world               # the reducer leaves the MEAN in the arena (what DDP leaves in .grad)
    got = store.grad_arena
    assert red.grad_scale == 1.0
    ok = torch.allclose(got, expect, rtol=1e-6, atol=1e-6)
    a, b = ddp.shard_range(128, rank, world)
    out_q.put((rank, ok, red.calls, red.bytes_reduced, (a, b)))
    dist.destroy_process_group()


@pytest.mark.parametrize("model", ["retina", "ssd_vgg"])
def test_bucketed_allreduce_two_ranks_gloo(model):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, model)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, calls, nbytes, shard in res:
        assert ok, f"rank {rank}: reduced gradients differ from the mean of the per-rank gradients"
        assert 2 <= calls <= 12                       # coalesced buckets, not one call per stage
        assert shard == (rank * 64, rank * 64 + 64)
    assert res[0][3] == res[1][3] and res[0][3] % 4 == 0


def test_single_rank_reducer_is_a_noop():
    import zsg_b200  # noqa: F401
    from zsg_b200 import ddp, engine
    store = engine.ParamStore(torch.device("cpu"))
    red = ddp.GradReducer(store)
    red.on_bucket(0, store.used)
    red.finish()
    assert red.world == 1 and red.calls == 0 and red.grad_scale == 1.0
