"""The data-parallel scheduler on real GPUs: 2 ranks over NCCL (skipped on a 1-GPU box) -- the gathered ids of
scheduler.generate_sharded equal the single-GPU run's, for multinomial sampling (shard-invariant Philox draws) with an
uneven split, and for greedy; the asynchronous form returns the same bytes."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STOI = {"<PAD>": 0, "<UNK>": 1, "<EOS>": 2, "<SOS>": 3, "<MASK>": 4}
B, K, T = 7, 5, 12


def _single(dev):
    import multimodalspectraltransformer_b200 as M
    from multimodalspectraltransformer_b200 import synthetic
    cfg = M.default_config(device=str(dev), max_len=T, precision="bf16")
    torch.manual_seed(0)
    model = M.MultimodalTransformer(cfg).eval()
    data = synthetic.make_spectra(B, seed=909)
    return M, cfg, model, data


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from multimodalspectraltransformer_b200 import scheduler
    M, cfg, model, data = _single(dev)
    gen = torch.cuda.default_generators[rank]
    torch.manual_seed(321)
    off0 = gen.get_offset()
    tok, pr, span = scheduler.generate_sharded(model, data, cfg, STOI, n_candidates=K, sampling="multinomial", gather_probs=True)
    off1 = gen.get_offset()
    torch.manual_seed(321)
    pend = scheduler.generate_sharded(model, data, cfg, STOI, n_candidates=K, sampling="multinomial", async_gather=True)
    same_async = bool(torch.equal(pend.tokens(), tok)) and bool(torch.equal(pend.packed().t().long(), tok))
    gtok, _, _ = scheduler.generate_sharded(model, data, cfg, STOI, n_candidates=1, sampling="greedy")
    torch.cuda.synchronize()
    q.put((rank, tok.cpu(), pr.cpu(), gtok.cpu(), off1 - off0, same_async, span))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_generate_sharded_two_ranks_equals_single_gpu():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
    M, cfg, model, data = _single(torch.device("cuda", 0))
    memory, mask, *_ = M.run_model(model, data, cfg)
    gen = torch.cuda.default_generators[0]
    torch.manual_seed(321)
    off0 = gen.get_offset()
    tok1, pr1 = M.multinomial_sequence_multi(model, memory, mask, STOI, cfg, n_candidates=K)
    inc1 = gen.get_offset() - off0
    g1, _ = M.greedy_sequence(model, STOI, None, memory, mask, cfg)
    assert [r[6] for r in res] == [(0, 4), (4, 7)]
    for rank, tok, pr, gtok, inc, same_async, _ in res:
        assert torch.equal(tok, tok1.cpu()), rank
        assert torch.equal(pr, pr1.cpu()), rank
        assert torch.equal(gtok, g1.cpu()), rank
        assert inc == inc1 and same_async, rank
