"""Device time of the batched beam search (mmt_beam_search): python profiles/beam_timing.py [B] [beam] [gen_len] [precision]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic
from multimodalspectraltransformer_b200.engine import engine_for
from multimodalspectraltransformer_b200.generate import _mask_to_bias

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T = int(sys.argv[3]) if len(sys.argv) > 3 else 64
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
cfg = M.default_config(device="cuda", precision=prec)
torch.manual_seed(0)
model = M.MultimodalTransformer(cfg).eval()
data = synthetic.make_spectra(B, seed=5)
memory, mask, *_ = M.run_model(model, data, cfg)
eng = engine_for(model, cfg)
bias = _mask_to_bias(mask)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    seq, ln, score, probs, steps = eng.beam_search(memory, bias, beam_size=K, gen_len=T, eos=2, precision=prec)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"{prec}: {B} spectra x {K} beams x {steps} steps: {dt*1e3:.1f} ms ({dt*1e6/steps:.0f} us/step, {B*K*steps/dt/1e6:.2f} M beam-tokens/s)")
if len(sys.argv) > 5:      # CPU oracle (the reference's algorithm: one decoder call per beam per item per step) on a sample
    from oracle import mmt_oracle as O
    P = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    n = int(sys.argv[5])
    t0 = time.perf_counter()
    with torch.no_grad():
        O.beam_search(P, memory[:, :n].cpu(), mask[:n].cpu(), O.default_config(), K, T)
    dt = time.perf_counter() - t0
    print(f"oracle port of vgmmt.beam_search on the host cores: {n} spectra x {K} beams x {T} steps: {dt:.1f} s ({n*K*T/dt:.0f} beam-tokens/s)")
