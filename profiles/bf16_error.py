"""bf16-mode logit error against the reference's fp32 golden logits, per golden case:
max |logit - ref| / max|ref row|  (the north_star tolerance is 1e-2) and greedy agreement."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from golden_util import CASES, load_case
from test_gpu_parity import cfg_for, setup

s = setup()
for name in CASES:
    case, data, z = load_case(name)
    for prec in ("fp32", "bf16"):
        cfg = cfg_for(case, precision=prec)
        memory, mask, *_ = s["M"].run_model(s["model"], data, cfg)
        logits = s["M"].teacher_forced_logits(s["model"], memory, mask, torch.from_numpy(z["tf_tokens"]), cfg).cpu().numpy()
        ref = z["tf_logits"]
        scale = np.abs(ref).max(axis=-1, keepdims=True)
        rel = np.abs(logits - ref) / scale
        stride = int(z["memory_stride"])
        mem_err = np.abs(memory[::stride].cpu().numpy() - z["memory_sample"])
        rms = float(np.sqrt(np.mean(z["memory_sample"] ** 2)))
        agree = float((logits.argmax(-1) == ref.argmax(-1)).mean())
        print(f"{name:22s} {prec}: logits max rel {rel.max():.2e} mean rel {rel.mean():.2e} argmax agree {agree:.4f} | "
              f"memory rms err {np.sqrt(np.mean(mem_err**2))/rms:.2e} max {mem_err.max()/rms:.2e}", flush=True)
