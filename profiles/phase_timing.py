"""Device time of the two phases of one bench step (CUDA events on the current stream):
encode (256 spectra) and greedy decode (128 positions), per precision, graph replay on/off."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic
from multimodalspectraltransformer_b200.engine import engine_for

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
max_len = int(sys.argv[2]) if len(sys.argv) > 2 else 128
for prec in sys.argv[3:] or ["bf16", "fp32"]:
    cfg = M.default_config(device="cuda", precision=prec, max_len=max_len)
    torch.manual_seed(0)
    model = M.MultimodalTransformer(cfg).eval()
    data = {k: v.cuda() for k, v in synthetic.make_spectra(B, seed=1000).items()}
    eng = engine_for(model, cfg)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    enc, dec = [], []
    for it in range(6):
        a, b, c = ev(), ev(), ev()
        a.record()
        memory, mask, *_ = M.run_model(model, data, cfg)
        b.record()
        tok, pr = M.greedy_sequence(model, {"<SOS>": 3}, None, memory, mask, cfg)
        c.record()
        torch.cuda.synchronize()
        if it >= 3:
            enc.append(a.elapsed_time(b)); dec.append(b.elapsed_time(c))
    print(f"{prec} B={B} T={max_len} graph={'off' if os.environ.get('MMT_NO_GRAPH') else 'on'}: "
          f"encode {sum(enc)/len(enc):.2f} ms, decode {sum(dec)/len(dec):.2f} ms ({1e3*sum(dec)/len(dec)/max_len:.1f} us/step)", flush=True)
