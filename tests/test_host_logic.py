"""Host-side logic that needs no GPU: the ids -> SMILES helpers on host tensors, the inference-only guard of forward(),
the locked library build."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_tensor_to_smiles_and_prob_2(tensor, token_prob, itos):
    """The reference's element-by-element loop (helper_functions_pl_v15_4.py:390-410), restated for the check."""
    seqs, cut = [], []
    for i in range(tensor.shape[1]):
        seq = []
        for j in range(tensor.shape[0]):
            tok = itos[str(int(tensor[j, i]))]
            if tok == "<EOS>":
                break
            seq.append(tok)
        seqs.append("".join(seq))
        cut.append(token_prob[:len(seq), i])
    return seqs, cut


def test_tensor_to_smiles_accepts_host_tensors_like_the_reference():
    """Callers that .cpu() the ids first (helper_functions_pl_v15_4.py:247-301 take either): same strings / slices."""
    import multimodalspectraltransformer_b200 as M
    itos = {str(i): f"t{i}|" for i in range(43)}
    itos.update({"0": "<PAD>", "2": "<EOS>", "3": "<SOS>"})
    g = torch.Generator().manual_seed(3)
    tok = torch.randint(0, 43, (64, 50), generator=g)
    tok[:, 5] = 7                        # never emits <EOS>
    tok[0, 6] = 2                        # <EOS> first -> empty string
    pr = torch.rand(64, 50, generator=g)
    ref_s, ref_p = _ref_tensor_to_smiles_and_prob_2(tok, pr, itos)
    got_s, got_p = M.tensor_to_smiles_and_prob_2(tok, pr, itos)
    assert got_s == ref_s and got_s[6] == "" and all(torch.equal(a, b) for a, b in zip(got_p, ref_p))
    assert M.tensor_to_smiles(tok, itos) == ref_s
    got_s1, got_p1 = M.tensor_to_smiles_and_prob(tok, pr.t().contiguous(), itos)
    assert got_s1 == ref_s and all(torch.equal(a, b) for a, b in zip(got_p1, ref_p))
    one_s, one_p = M.tensor_to_smiles_and_prob_2(tok[:, 9], pr[:, 9], itos)
    assert one_s == ref_s[9] and torch.equal(one_p, ref_p[9])


def test_forward_refuses_training_mode():
    """nn.Transformer layers apply dropout 0.1 in train() whatever config.drop_out says and the engine's outputs carry no
    autograd graph: forward() must refuse instead of returning eval-mode tensors (both branches)."""
    import multimodalspectraltransformer_b200 as M
    from multimodalspectraltransformer_b200 import synthetic
    cfg = M.default_config(device="cuda", num_encoder_layers=1, num_decoder_layers=1, drop_out=0.0)
    model = M.MultimodalTransformer(cfg)
    model.train()
    d = synthetic.make_spectra(2)
    args = [d[k] for k in ("src_1H", "mask_1H", "src_13C", "mask_13C", "src_HSQC", "mask_HSQC", "src_COSY", "mask_COSY",
                           "src_IR", "mask_IR", "src_MF", "mask_MF", "src_MS", "mask_MS", "trg_MW")]
    with pytest.raises(RuntimeError, match="inference-only"):
        model(*args)
    with pytest.raises(RuntimeError, match="inference-only"):
        model(*args, d["trg_enc_SMI"].t()[:4])


def test_concurrent_builds_never_expose_a_partial_library():
    """Two processes calling _lib.build() at once (torchrun ranks on a box with a compiler): both return the same loadable
    library; the build is serialised by a file lock and moved into place atomically."""
    code = ("import sys; sys.path.insert(0, %r); from multimodalspectraltransformer_b200 import _lib; import ctypes; "
            "p = _lib.build(); L = ctypes.CDLL(p); L.mmt_abi_version.restype = ctypes.c_int32; print(L.mmt_abi_version())" % ROOT)
    procs = [subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for _ in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert [o[0].strip() for o in outs] == ["2", "2"]
    from multimodalspectraltransformer_b200 import _lib
    assert not [f for f in os.listdir(os.path.dirname(_lib.LIB_PATH)) if f.endswith(".so.tmp")]
