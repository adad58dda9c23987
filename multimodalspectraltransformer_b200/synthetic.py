"""Synthetic spectra in the reference's collated tensor contract.

Shapes, dtypes, normalisation and mask polarity follow what the reference's
``MultimodalData.__getitem__`` + ``collate_fn`` emit
(reference utils_MMT/dataloaders_pl_v15_4.py:426-661, 665-712):

    src_1H   (B,64,2) f32  rows [ppm/10, integral], zero rows after k peaks
    src_13C  (B,64)   f32  ppm/200
    src_HSQC (B,64,2) f32  [ppm_H/10, ppm_C/200]
    src_COSY (B,64,2) f32  [ppm/10, ppm/10]
    mask_X   (B,64)   f32  0 = valid peak, 1 = padding
    src_IR   (B,1000) f32  max-normalised absorbance, mask_IR zeros (unused)
    src_MF   (B,64)   i64  [<SOS>=3, formula tokens, <EOS>=2, 0...]
    mask_MF  (B,64)        0 valid / 1 pad.  The reference collates this as int64,
                           which torch>=2 rejects as a key-padding mask; the only
                           runnable semantics is bool (True = pad), SURVEY.md B.1.
    trg_MW   (B,)     f32  molecular weight in Da (raw)
    trg_enc_SMI (B,64) i64 [3, tokens, 2, 0...]

There is no dataset on the box (no network), so the bench and the tests draw
these from a seeded CPU generator.  ``peaks="realistic"`` draws peak counts
around the means of the reference's shipped example CSVs (1H 36.7, 13C 15.4,
HSQC 12.5, COSY 27.2, ~10 formula tokens); ``peaks="max"`` fills every slot
(BASELINE.json config 5).
"""
from __future__ import annotations

import torch

PAD_POINTS = 64
IR_BINS = 1000
SOS, EOS = 3, 2

_COUNT_RANGES = {            # inclusive [lo, hi] for peaks="realistic"
    "1H": (10, 64),
    "13C": (4, 27),
    "HSQC": (3, 22),
    "COSY": (6, 48),
    "MF": (6, 14),           # formula tokens between <SOS> and <EOS>
}


def _counts(gen, B, key, peaks):
    if peaks == "max":
        n = PAD_POINTS if key != "MF" else PAD_POINTS - 2
        return torch.full((B,), n, dtype=torch.int64)
    lo, hi = _COUNT_RANGES[key]
    return torch.randint(lo, hi + 1, (B,), generator=gen)


def _pad_mask(counts):
    ar = torch.arange(PAD_POINTS).unsqueeze(0)
    return (ar >= counts.unsqueeze(1))


def make_spectra(B: int, seed: int = 0, peaks: str = "realistic", blank=(), mask_mf_dtype=torch.bool):
    """Return the collated ``data_dict`` for ``B`` synthetic spectra (CPU tensors).

    ``blank`` lists modalities ("1H","13C","HSQC","COSY") to blank the way the
    reference's data loader does when a spectrum is missing: zeros + all-ones
    mask (dataloaders_pl_v15_4.py:369-392, 468-470).
    """
    gen = torch.Generator().manual_seed(seed)
    d = {}

    def u(*shape, scale=1.0):
        return torch.rand(*shape, generator=gen) * scale

    for key, cols, scales in (("1H", 2, (1.0, 2.0)), ("13C", 1, (1.0,)),
                              ("HSQC", 2, (1.0, 1.0)), ("COSY", 2, (1.0, 1.0))):
        n = _counts(gen, B, key, peaks)
        pad = _pad_mask(n)
        x = torch.stack([u(B, PAD_POINTS, scale=s) for s in scales], dim=-1)
        x = x.masked_fill(pad.unsqueeze(-1), 0.0)
        if key in blank:
            x = torch.zeros_like(x)
            pad = torch.ones_like(pad)
        if cols == 1:
            x = x.squeeze(-1)
        d[f"src_{key}"] = x.float().contiguous()
        d[f"mask_{key}"] = pad.float()
    d["src_IR"] = u(B, IR_BINS).float()
    d["mask_IR"] = torch.zeros(B, IR_BINS)

    n_mf = _counts(gen, B, "MF", peaks)
    mf = torch.randint(5, 212, (B, PAD_POINTS), generator=gen)
    ar = torch.arange(PAD_POINTS).unsqueeze(0)
    mf = torch.where(ar == 0, torch.full_like(mf, SOS), mf)
    mf = torch.where(ar == (n_mf + 1).unsqueeze(1), torch.full_like(mf, EOS), mf)
    mf_pad = ar > (n_mf + 1).unsqueeze(1)
    mf = mf.masked_fill(mf_pad, 0)
    d["src_MF"] = mf
    d["mask_MF"] = mf_pad.to(mask_mf_dtype)

    d["trg_MW"] = (100.0 + 400.0 * u(B)).float()

    n_smi = torch.randint(8, 40, (B,), generator=gen)
    smi = torch.randint(5, 43, (B, PAD_POINTS), generator=gen)
    smi = torch.where(ar == 0, torch.full_like(smi, SOS), smi)
    smi = torch.where(ar == (n_smi + 1).unsqueeze(1), torch.full_like(smi, EOS), smi)
    smi = smi.masked_fill(ar > (n_smi + 1).unsqueeze(1), 0)
    d["trg_enc_SMI"] = smi
    d["src_MS"] = smi.clone()
    d["mask_MS"] = (ar > (n_smi + 1).unsqueeze(1))
    d["src_HSQC_"] = d["src_HSQC"].clone()
    d["src_COSY_"] = d["src_COSY"].clone()
    return d


def slice_spectra(d, lo, hi):
    """Rows [lo, hi) of every entry (contiguous shard of spectra)."""
    return {k: v[lo:hi] for k, v in d.items()}
