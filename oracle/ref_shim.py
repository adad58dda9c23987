"""Shim loader for the *unmodified* reference modules under /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (in the build
container, where /root/reference is mounted) to execute the reference's own
Python code so that the CPU restatement in ``oracle/mmt_oracle.py`` can be
pinned against it.  Nothing under the product package imports this file and
the GPU box never has /root/reference, so nothing on a ``-m gpu`` path may
call :func:`load_reference`.

The reference imports a number of third-party packages that are absent from
this image (pytorch_lightning, rdkit, dgl, matplotlib, IPython, ...; see
SURVEY.md section 8c).  None of them is touched by the hot path
(``MultimodalTransformer``, ``run_model``, ``greedy_sequence``,
``multinomial_sequence``, ``multinomial_sequence_multi``), so they are
replaced by inert stand-ins injected into ``sys.modules`` before import.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import importlib.abc
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("MMT_REFERENCE_ROOT", "/root/reference")

_STUB_ROOTS = (
    "rdkit", "dgl", "dgllife", "matplotlib", "IPython", "wandb", "umap",
    "plotly", "molvs", "cairosvg", "tensorboardX", "seaborn", "sklearn_extra",
    "pytorch_lightning", "flask", "flask_socketio", "chemprop", "nmr_sgnn_norm",
    "rdkit_contrib", "PIL", "networkx", "selfies", "mordred", "openbabel",
)


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []  # behave as a package so sub-imports resolve
        return m

    def exec_module(self, module):
        pass


class _StubModule(types.ModuleType):
    """Module whose every attribute is a MagicMock (created on demand)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        v = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, v)
        return v


class _StubFinder(importlib.abc.MetaPathFinder):
    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in self.roots:
            return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)
        return None


def _install_stubs():
    import torch.nn as nn

    missing = []
    for root in _STUB_ROOTS:
        try:
            if root in sys.modules:
                continue
            importlib.util.find_spec(root)
            if importlib.util.find_spec(root) is None:
                missing.append(root)
        except (ImportError, ValueError):
            missing.append(root)
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder(missing))
    else:
        for f in sys.meta_path:
            if isinstance(f, _StubFinder):
                f.roots.update(missing)

    # pytorch_lightning.LightningModule must be a real class (it is subclassed).
    if "pytorch_lightning" in missing:
        pl = importlib.import_module("pytorch_lightning")
        pl.LightningModule = nn.Module
        pl.Trainer = mock.MagicMock(name="pl.Trainer")

    # The SGNN glue loads normalisation statistics from disk at import time
    # (validate_generate_MMT_v15_4.py:1200-1207); give it a fake.
    name = "utils_MMT.sgnn_code_pl_v15_4"
    if name not in sys.modules:
        sg = _StubModule(name)
        sg.load_std_mean = lambda *a, **k: (0.0, 1.0)
        sys.modules[name] = sg
    # Heavy side modules imported by run_batch_gen_val (CLIP, train/test fns,
    # SMILES augmenter) are not on the hot path.
    for extra in ("utils_MMT.models_CLIP_v15_4", "utils_MMT.train_test_functions_pl_v15_4",
                  "utils_MMT.smi_augmenter_v15_4", "utils_MMT.clip_functions_v15_4"):
        if extra not in sys.modules:
            sys.modules[extra] = _StubModule(extra)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "utils_MMT"))


def load_reference():
    """Import the reference's hot-path modules; returns a namespace.

    Attributes: ``models`` (models_MMT_v15_4), ``vgmmt``
    (validate_generate_MMT_v15_4), ``rbgvm`` (run_batch_gen_val_MMT_v15_4).
    NB importing vgmmt/rbgvm reseeds every RNG with a random seed
    (validate_generate_MMT_v15_4.py:44-51): re-seed AFTER calling this.
    """
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        models = importlib.import_module("utils_MMT.models_MMT_v15_4")
        vgmmt = importlib.import_module("utils_MMT.validate_generate_MMT_v15_4")
        rbgvm = importlib.import_module("utils_MMT.run_batch_gen_val_MMT_v15_4")
    return types.SimpleNamespace(models=models, vgmmt=vgmmt, rbgvm=rbgvm)


def load_reference_config(device: str = "cpu"):
    """config_V8.json -> Namespace, un-listing each value the way the reference
    does (execution_function_v15_4.py:20-23)."""
    import argparse
    import json
    with open(os.path.join(REFERENCE_ROOT, "utils_MMT", "config_V8.json")) as f:
        raw = json.load(f)
    cfg = argparse.Namespace(**{k: v[0] for k, v in raw.items()})
    cfg.device = device
    return cfg
