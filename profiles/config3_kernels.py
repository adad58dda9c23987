"""Per-kernel device time of one config-3 wave (128 spectra x 128 candidates, un-fused decode kernels), event-timed by
the engine's profile mode (launches serialised): python profiles/config3_kernels.py [T] [spectra]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic
from multimodalspectraltransformer_b200.engine import engine_for
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cfg = M.default_config(device="cuda", precision="bf16", max_len=T)
torch.manual_seed(0)
model = M.MultimodalTransformer(cfg).eval()
data = {k: v.cuda() for k, v in synthetic.make_spectra(B, seed=1000).items()}
memory, mask, *_ = M.run_model(model, data, cfg)
eng = engine_for(model, cfg)
M.multinomial_sequence_multi(model, memory, mask, {"<SOS>": 3}, cfg, n_candidates=128)
torch.cuda.synchronize()
eng.profile(True)
M.multinomial_sequence_multi(model, memory, mask, {"<SOS>": 3}, cfg, n_candidates=128)
rep = eng.profile_report()
eng.profile(False)
tot = sum(v["ms"] for v in rep.values())
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:28s} {v['launches']:5d} launches {v['ms']:9.3f} ms {100*v['ms']/tot:5.1f} %  {1e3*v['ms']/v['launches']:8.1f} us each")
print(f"total {tot:.2f} ms for {T} positions of {B * 128} sequences")
