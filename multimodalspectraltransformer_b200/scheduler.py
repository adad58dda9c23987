"""Data-parallel batch scheduler: shard spectra across ranks, decode locally,
all-gather only the generated token ids.

Every spectrum (and every candidate of a spectrum) is independent on this path
(SURVEY.md 8e), so ranks own contiguous blocks of spectra, keep a full replica of
the 25.6 M-parameter model, and exchange nothing until the end: one NCCL
all-gather of the ids as bytes (vocab 43 < 256), optionally the chosen-token
probabilities.  The reference has no multi-GPU inference; the oracle for the
sharded run is the single-GPU run (shard-invariant Philox indexing, see
generate.multinomial_sequence_multi's seq_index_base / n_total).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int, rank: int):
    """Contiguous block [lo, hi) of ``n_items`` for ``rank``: ceil-sized blocks, the tail ranks may be short or empty."""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    hi = min(lo + per, n_items)
    return lo, hi


def shard_dict(data_dict, world: int, rank: int):
    n = next(iter(data_dict.values())).shape[0]
    lo, hi = shard_bounds(n, world, rank)
    return {k: v[lo:hi] for k, v in data_dict.items()}, lo, hi


def gather_columns(local: torch.Tensor, n_cols_total: int, cols_per_rank: int, group=None) -> torch.Tensor:
    """All-gather column blocks: every rank holds ``local`` (T, <=cols_per_rank) and
    receives (T, n_cols_total).  Short tail shards are padded for the collective and
    trimmed afterwards.  Works for any dtype / backend (NCCL on GPUs, gloo in the CPU tests)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    T = local.shape[0]
    send = local
    if local.shape[1] != cols_per_rank:
        send = torch.zeros(T, cols_per_rank, dtype=local.dtype, device=local.device)
        send[:, :local.shape[1]] = local
    send = send.contiguous()
    recv = torch.empty(world * T, cols_per_rank, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)       # rank r's block lands in rows [r*T, (r+1)*T)
    out = recv.view(world, T, cols_per_rank).permute(1, 0, 2).reshape(T, world * cols_per_rank)
    return out[:, :n_cols_total].contiguous()


def generate_sharded(model, data_dict, config, stoi, *, n_candidates=1, sampling="multinomial", gather_probs=False,
                     group=None):
    """Encode + decode this rank's spectra and all-gather the ids.

    Returns (tokens (T, B*n_candidates) i64 -- identical on every rank and identical to
    the single-GPU run --, probs or None, local (lo, hi))."""
    from . import generate as G
    from .engine import engine_for
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = next(iter(data_dict.values())).shape[0]
    local, lo, hi = shard_dict(data_dict, world, rank)
    per = (B + world - 1) // world
    T = int(config.max_len)
    eng = engine_for(model, config)
    if hi > lo:
        memory, mask, *_ = G.run_model(model, local, config)
        if sampling == "greedy":
            # stop_on_all_pad is a whole-batch property: decide it after the gather
            tok, pr, _ = G._decode(model, memory, mask, config, "greedy", False, n_candidates, 0, 0)
        else:
            tok, pr, _ = G._decode(model, memory, mask, config, "multinomial", False, n_candidates,
                                   lo * n_candidates, B * n_candidates)
    else:
        tok = torch.zeros(T, 0, dtype=torch.int64, device=eng.device)
        pr = torch.zeros(T, 0, dtype=torch.float32, device=eng.device)
        if sampling != "greedy":
            # an empty shard draws nothing, but the unsharded run's T multinomial calls would have advanced this device's
            # generator: keep every rank's Philox offset in step so later sharded calls still reproduce the 1-GPU draws
            gen = torch.cuda.default_generators[eng.dev_index]
            gen.set_offset(int(gen.get_offset()) + eng.philox_increment(B * n_candidates) * T)
    packed = eng.pack_tokens(tok) if tok.numel() else torch.zeros(T, 0, dtype=torch.uint8, device=eng.device)
    all_u8 = gather_columns(packed, B * n_candidates, per * n_candidates, group)
    tokens = eng.unpack_tokens(all_u8)
    probs = gather_columns(pr, B * n_candidates, per * n_candidates, group) if gather_probs else None
    if sampling == "greedy":
        allpad = (tokens == 0).all(dim=1)
        if bool(allpad.any()):
            steps = int(torch.nonzero(allpad)[0]) + 1
            tokens = tokens[:steps]
            probs = probs[:steps] if probs is not None else None
    return tokens, probs, (lo, hi)
