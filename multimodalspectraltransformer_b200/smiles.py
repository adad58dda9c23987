"""ids -> SMILES strings + truncated token probabilities, the reference's post-processing of the
generated tensors (reference utils_MMT/helper_functions_pl_v15_4.py):

    tensor_to_smiles              :247-269
    tensor_to_smiles_and_prob     :272-301
    tensor_to_smiles_and_prob_2   :390-419

The reference walks the tensors element by element with ``.item()`` (one device-to-host sync per
token).  Here the first-<EOS> scan runs on the device (``mmt_first_eos``), the ids cross to the
host once as bytes, and only the string join and the probability slicing remain in Python.  Same
arguments, same return layouts.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _eos_id(itos) -> int:
    for k, v in itos.items():
        if v == "<EOS>":
            return int(k)
    raise KeyError("<EOS> not in itos")


def _lengths_and_ids(tensor, eos):
    """(T,N) ids on a CUDA device -> (first-EOS position per column as list, ids as a CPU uint8/int64 array)."""
    t = tensor if tensor.dim() == 2 else tensor.unsqueeze(1)
    t = t.to(torch.int64).contiguous()
    T, N = t.shape
    if not t.is_cuda:
        # ids the caller already moved to the host (the reference's helpers take either, helper_functions_pl_v15_4.py:247-301):
        # the scan is a host-side index computation on data that is already there -- no kernel, nothing copied back and forth
        is_eos = (t == eos)
        first = torch.where(is_eos.any(dim=0), is_eos.to(torch.uint8).argmax(dim=0), torch.full((N,), T, dtype=torch.int64))
        return first.tolist(), t.numpy()
    lens = torch.empty(N, dtype=torch.int32, device=t.device)
    stream = C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    _lib.check(_lib.lib().mmt_first_eos(t.data_ptr(), T, N, eos, lens.data_ptr(), stream))
    ids = t.to(torch.uint8) if int(T) and eos < 256 else t
    return lens.cpu().tolist(), ids.cpu().numpy()


def tensor_to_smiles(tensor, itos):
    """(T,N) -> list of N strings; (T,) -> one string."""
    lens, ids = _lengths_and_ids(tensor, _eos_id(itos))
    out = ["".join(itos[str(int(v))] for v in ids[:lens[i], i]) for i in range(ids.shape[1])]
    return out if tensor.dim() > 1 else out[0]


def tensor_to_smiles_and_prob(tensor, token_prob, itos):
    """tensor (T,N), token_prob (N,T') -> (strings, [token_prob[i, :eos_i]]); 1-D input -> (string, prob[:eos])."""
    lens, ids = _lengths_and_ids(tensor, _eos_id(itos))
    if tensor.dim() > 1:
        seqs = ["".join(itos[str(int(v))] for v in ids[:lens[i], i]) for i in range(ids.shape[1])]
        return seqs, [token_prob[i, :lens[i]] for i in range(ids.shape[1])]
    smi = "".join(itos[str(int(v))] for v in ids[:lens[0], 0])
    return smi, torch.stack(list(token_prob[:lens[0]]))


def tensor_to_smiles_and_prob_2(tensor, token_prob, itos):
    """tensor (T,N), token_prob (T',N) -> (strings, [token_prob[:len_i, i]]); 1-D input -> (string, prob[:len])."""
    lens, ids = _lengths_and_ids(tensor, _eos_id(itos))
    if tensor.dim() > 1 and token_prob.dim() > 1:
        seqs = ["".join(itos[str(int(v))] for v in ids[:lens[i], i]) for i in range(ids.shape[1])]
        return seqs, [token_prob[:lens[i], i] for i in range(ids.shape[1])]
    smi = "".join(itos[str(int(v))] for v in ids[:lens[0], 0])
    return smi, torch.stack(list(token_prob[:lens[0]]))
