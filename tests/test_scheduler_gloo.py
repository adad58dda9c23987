"""Host-side logic of the data-parallel scheduler on CPU: world_size-2 gloo processes
(shard bounds, padded all-gather of id columns)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_partition():
    from multimodalspectraltransformer_b200.scheduler import shard_bounds
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            per = (n + world - 1) // world
            assert all(hi - lo <= per for lo, hi in spans)


def _worker(rank, world, port, B, k, T, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimodalspectraltransformer_b200.scheduler import shard_bounds, gather_columns, gather_rows, PendingTokens
    full = (torch.arange(T * B * k) % 43).reshape(T, B * k).to(torch.uint8)
    lo, hi = shard_bounds(B, world, rank)
    local = full[:, lo * k:hi * k].contiguous()
    per = (B + world - 1) // world
    out = gather_columns(local, B * k, per * k)
    ok = bool(torch.equal(out, full))
    # the scheduler's own payload: sequence-major (n_local, T) blocks land in the final order, no re-layout
    rows = gather_rows(local.t().contiguous(), per * k)
    ok = ok and tuple(rows.shape) == (world * per * k, T) and bool(torch.equal(rows[:B * k], full.t()))
    pend = PendingTokens(None, rows, None, B * k, T)
    ok = ok and bool(torch.equal(pend.tokens(), full.to(torch.int64))) and bool(torch.equal(pend.packed(), full.t()))
    pr = gather_columns(local.float(), B * k, per * k)
    ok = ok and bool(torch.equal(pr, full.float()))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B,k", [(8, 4), (7, 3), (1, 5)])
def test_gather_columns_world2_gloo(B, k):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + B * 7 + k) % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, k, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == [0, 1] and all(ok for _, ok in res)
