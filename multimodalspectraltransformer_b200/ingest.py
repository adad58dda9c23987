"""Ragged ingest: per-spectrum peak lists -> the collated ``data_dict`` the path consumes.

The reference builds every padded tensor on the host, one spectrum at a time
(``MultimodalData.__getitem__`` / ``_zero_pad`` / ``_load_IR_data`` / ``collate_fn``,
utils_MMT/dataloaders_pl_v15_4.py:267-299, 324-346, 352-365, 440-560, 665-712).  Here the raw lists are
packed into CSR arrays once, copied to the GPU, and two kernels (``mmt_ingest_peaks``, ``mmt_ingest_ir``)
emit the tensors of the collate contract directly in device memory.  The molecular
formula tokens, MW and target SMILES are tiny integer/string work and stay on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

# divisors of the reference: 1H [ppm/10, integral], 13C ppm/200, HSQC [ppm_H/10, ppm_C/200], COSY [ppm/10, ppm/10]
DIVISORS = {"1H": (10.0, 1.0), "13C": (200.0, 1.0), "HSQC": (10.0, 200.0), "COSY": (10.0, 10.0)}
COLS = {"1H": 2, "13C": 1, "HSQC": 2, "COSY": 2}


def _csr(lists, cols):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    for i, x in enumerate(lists):
        off[i + 1] = off[i] + len(x)
    vals = np.zeros((int(off[-1]), cols), dtype=np.float64)
    for i, x in enumerate(lists):
        if len(x):
            vals[off[i]:off[i + 1]] = np.asarray(x, dtype=np.float64).reshape(len(x), cols)
    return vals, off


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def peaks_to_padded(lists, modality, device="cuda", pad_points=64):
    """lists[b] = peaks of spectrum b (1H / HSQC / COSY: [[x, y], ...]; 13C: [x, ...]) -> (src, mask) on ``device``."""
    cols = COLS[modality]
    vals, off = _csr(lists, cols)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("mmt_b200 ingest runs on the CUDA device (no CPU path)")
    B = len(lists)
    d_vals = torch.from_numpy(vals).to(dev) if vals.size else torch.zeros(1, cols, dtype=torch.float64, device=dev)
    d_off = torch.from_numpy(off).to(dev)
    src = torch.empty((B, pad_points, cols) if cols == 2 else (B, pad_points), dtype=torch.float32, device=dev)
    mask = torch.empty(B, pad_points, dtype=torch.float32, device=dev)
    d0, d1 = DIVISORS[modality]
    _lib.check(_lib.lib().mmt_ingest_peaks(d_vals.data_ptr(), d_off.data_ptr(), B, cols, d0, d1, pad_points,
                                           src.data_ptr(), mask.data_ptr(), _stream(dev)))
    for t in (d_vals, d_off):
        t.record_stream(torch.cuda.current_stream(dev))
    return src, mask


def ir_to_binned(spectra, device="cuda", bins=1000):
    """spectra[b] = raw IR absorbances of any length -> (B, bins) f32 mean-binned / max-normalised; mask zeros (B, bins)."""
    vals, off = _csr(spectra, 1)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("mmt_b200 ingest runs on the CUDA device (no CPU path)")
    B = len(spectra)
    d_vals = torch.from_numpy(vals.reshape(-1)).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    out = torch.empty(B, bins, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().mmt_ingest_ir(d_vals.data_ptr(), d_off.data_ptr(), B, bins, out.data_ptr(), _stream(dev)))
    for t in (d_vals, d_off):
        t.record_stream(torch.cuda.current_stream(dev))
    return out, torch.zeros(B, bins, dtype=torch.float32, device=dev)


def collate_ragged(peaks, src_MF, mask_MF, trg_MW, trg_enc_SMI=None, device="cuda", pad_points=64, ir_bins=1000):
    """peaks: {"1H": [...], "13C": [...], "HSQC": [...], "COSY": [...], "IR": [...]} (lists per spectrum; a missing
    key gives the reference's blank modality: zeros + all-ones mask, dataloaders_pl_v15_4.py:369-392).
    src_MF / mask_MF (B,64) and trg_MW (B,) are host tensors in the collate contract.  Returns the ``data_dict``."""
    B = int(trg_MW.shape[0])
    d = {}
    for m in ("1H", "13C", "HSQC", "COSY"):
        if m in peaks:
            d[f"src_{m}"], d[f"mask_{m}"] = peaks_to_padded(peaks[m], m, device, pad_points)
        else:
            shape = (B, pad_points, 2) if COLS[m] == 2 else (B, pad_points)
            d[f"src_{m}"] = torch.zeros(shape, dtype=torch.float32, device=device)
            d[f"mask_{m}"] = torch.ones(B, pad_points, dtype=torch.float32, device=device)
    if "IR" in peaks:
        d["src_IR"], d["mask_IR"] = ir_to_binned(peaks["IR"], device, ir_bins)
    else:
        d["src_IR"] = torch.zeros(B, ir_bins, dtype=torch.float32, device=device)
        d["mask_IR"] = torch.zeros(B, ir_bins, dtype=torch.float32, device=device)
    d["src_MF"], d["mask_MF"] = src_MF.to(device), mask_MF.to(device)
    d["src_MS"] = torch.zeros(B, pad_points, dtype=torch.int64, device=device)
    d["mask_MS"] = torch.ones(B, pad_points, dtype=torch.bool, device=device)
    d["trg_MW"] = trg_MW.to(device, torch.float32)
    d["trg_enc_SMI"] = (trg_enc_SMI if trg_enc_SMI is not None else torch.zeros(B, pad_points, dtype=torch.int64)).to(device)
    d["src_HSQC_"], d["src_COSY_"] = d["src_HSQC"], d["src_COSY"]
    return d
