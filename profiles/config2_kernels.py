"""Per-kernel device time of one config-2 step (256 spectra, greedy, 128 tokens, bf16), event-timed by the engine's profile
mode (launches serialised, single lane, no graph / PDL): python profiles/config2_kernels.py [B] [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic
from multimodalspectraltransformer_b200.engine import engine_for
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
peaks = sys.argv[3] if len(sys.argv) > 3 else "realistic"
cfg = M.default_config(device="cuda", precision="bf16", max_len=T)
torch.manual_seed(0)
model = M.MultimodalTransformer(cfg).eval()
data = {k: v.cuda() for k, v in synthetic.make_spectra(B, seed=1000, peaks=peaks).items()}
eng = engine_for(model, cfg)
for _ in range(2):
    memory, mask, *_ = M.run_model(model, data, cfg)
    M.greedy_sequence(model, {"<SOS>": 3}, None, memory, mask, cfg)
torch.cuda.synchronize()
eng.profile(True)
memory, mask, *_ = M.run_model(model, data, cfg)
enc = eng.profile_report()
eng.profile(False); eng.profile(True)
M.greedy_sequence(model, {"<SOS>": 3}, None, memory, mask, cfg)
dec = eng.profile_report()
eng.profile(False)
for name, rep in (("encode", enc), ("decode", dec)):
    tot = sum(v["ms"] for v in rep.values())
    print(f"--- {name}: {tot:.3f} ms of kernel time ({B} spectra, {peaks} peaks)")
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
        tf = f"  {v['flops'] / (v['ms'] * 1e-3) / 1e12:7.1f} TFLOP/s" if v.get("flops", 0) > 0 else ""
        print(f"{k:28s} {v['launches']:5d} launches {v['ms']:9.3f} ms {100*v['ms']/tot:5.1f} %  {1e3*v['ms']/v['launches']:8.1f} us each{tf}")
