"""Config handling of the path: the reference reads a JSON whose values are
one-element lists into an argparse.Namespace that callers then mutate
(reference utils_MMT/execution_function_v15_4.py:20-23; config_V8.json)."""
from __future__ import annotations

import argparse
import json

# hot-path subset of utils_MMT/config_V8.json
_V8 = dict(hidden_size=128, num_heads=16, num_encoder_layers=6, num_decoder_layers=6, in_size=43, out_size=43,
           max_len=128, drop_out=0.1, fingerprint_size=512, input_dim_1H=2, input_dim_13C=1, input_dim_HSQC=2,
           input_dim_COSY=2, input_dim_IR=1000, MF_vocab_size=212, MS_vocab_size=43, padding_points_number=64,
           training_mode="1H_13C_HSQC_COSY_IR_MF_MW", temperature=1, use_real_data=False, batch_size=64,
           multinom_runs=1, device="cuda", gpu_num=1)

STOI = {"<PAD>": 0, "<UNK>": 1, "<EOS>": 2, "<SOS>": 3, "<MASK>": 4}


def default_config(**over):
    """config_V8.json's hot-path hyper-parameters as a mutable Namespace."""
    c = dict(_V8)
    c.update(over)
    return argparse.Namespace(**c)


def load_config(path, **over):
    """Parse a reference-style JSON config ({key: [value]}) exactly as the reference does."""
    with open(path) as f:
        raw = json.load(f)
    c = {k: (v[0] if isinstance(v, list) and len(v) == 1 else v) for k, v in raw.items()}
    c.update(over)
    return argparse.Namespace(**c)
