"""Helpers shared by the parity tests: load a golden fixture and regenerate its inputs."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLDEN, "weights_meta.json")))["cases"]


def load_case(name):
    from multimodalspectraltransformer_b200 import synthetic
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = json.loads(str(z["case"]))
    data = synthetic.make_spectra(case["B"], seed=case["seed"], peaks=case["peaks"], blank=tuple(case["blank"]))
    return case, data, z


def golden_mask(z):
    m = torch.from_numpy(z["mask"])
    return m
