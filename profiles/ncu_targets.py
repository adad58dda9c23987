"""Short workload for the ncu captures of the HBM-bound glue kernels and the tcgen05 GEMM (north_star: every kernel
choice evidenced by an ncu capture): one bf16 encode of 256 realistic spectra (embed_tokens, gemm_bf16_tc at the encoder
shapes, the K/V projection GEMM of decode_prepare_wave) and two greedy positions (sample_tokens).

    python profiles/ncu_targets.py [n_spectra] [positions] [n_candidates]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = M.default_config(device="cuda", precision="bf16", max_len=T)
torch.manual_seed(0)
model = M.MultimodalTransformer(cfg).eval()
data = {k: v.cuda() for k, v in synthetic.make_spectra(B, seed=1000).items()}
for _ in range(2):
    memory, mask, *_ = M.run_model(model, data, cfg)
    if K == 1:
        tok, pr = M.greedy_sequence(model, {"<SOS>": 3}, None, memory, mask, cfg)
    else:
        tok, pr = M.multinomial_sequence_multi(model, memory, mask, {"<SOS>": 3}, cfg, n_candidates=K)
torch.cuda.synchronize()
print("ok", tuple(memory.shape), tuple(tok.shape))
