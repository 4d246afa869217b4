"""Import the UNMODIFIED reference (MB-iSTFT-VITS models.py) as test / benchmark infrastructure.

Looks for the staged copy under baseline/_ref/ (see stage_ref.py; this is what exists on the GPU box) and falls back
to /root/reference (the build container).  Three import shims, none of which touches arithmetic:
  1. `monotonic_align` (unbuilt Cython extension, models.py:11; used only by the training forward) is stubbed;
  2. `librosa` / `librosa.util` (stft.py:32-33; used only by the legacy STFT class, never by TorchSTFT) are stubbed;
  3. on CPU only, `Tensor.cuda(cpu_device)` is a no-op (pqmf.py:78,79,86 call .cuda(device) unconditionally).
Only tests/, bench.py's reference legs and tools/make_golden.py import this module; the product never does.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.path.join(HERE, "_ref"), "/root/reference"]
_models = None
_path = None


def ref_path():
    for p in CANDIDATES:
        if os.path.isfile(os.path.join(p, "models.py")):
            return p
    return None


def available() -> bool:
    return ref_path() is not None


def import_reference():
    """-> the reference's `models` module (imported once)."""
    global _models, _path
    if _models is not None:
        return _models
    p = ref_path()
    if p is None:
        raise ImportError("the reference is neither staged under baseline/_ref (python baseline/stage_ref.py) nor at /root/reference")
    ma = types.ModuleType("monotonic_align")
    ma.maximum_path = None
    sys.modules.setdefault("monotonic_align", ma)
    if "librosa" not in sys.modules:
        lib, libu = types.ModuleType("librosa"), types.ModuleType("librosa.util")
        libu.pad_center = lambda d, size, axis=-1, **k: d
        libu.tiny = lambda x: np.finfo(np.float32).tiny
        libu.normalize = lambda S, norm=None, **k: S
        lib.util = libu
        sys.modules["librosa"] = lib
        sys.modules["librosa.util"] = libu
    _cuda = torch.Tensor.cuda

    def cuda(self, device=None, *a, **k):
        if device is not None and torch.device(device).type == "cpu":
            return self
        return _cuda(self, device, *a, **k)
    torch.Tensor.cuda = cuda
    sys.path.insert(0, p)
    import models  # noqa: the reference's models.py
    _models, _path = models, p
    return models


def reference_json(name):
    """The reference's own JSON config (configs/<name>.json)."""
    return json.load(open(os.path.join(ref_path(), "configs", name + ".json")))


def model_kwargs(cfg):
    """Constructor keywords of SynthesizerTrn for one of mb_istft_vits_b200.configs' geometry dicts (the values the
    reference JSON `model` sections carry; synthesis_module.py:106-112 passes **hps.model)."""
    return dict(
        inter_channels=cfg["inter_channels"], hidden_channels=cfg["hidden_channels"],
        filter_channels=768 if cfg["hidden_channels"] == 192 else 384, n_heads=2,
        n_layers=3 if cfg["hidden_channels"] == 96 else 6,
        kernel_size=3, p_dropout=0.1, resblock=cfg["resblock"],
        resblock_kernel_sizes=cfg["resblock_kernel_sizes"],
        resblock_dilation_sizes=cfg["resblock_dilation_sizes"],
        upsample_rates=cfg["upsample_rates"], upsample_initial_channel=cfg["upsample_initial_channel"],
        upsample_kernel_sizes=cfg["upsample_kernel_sizes"],
        gen_istft_n_fft=cfg["gen_istft_n_fft"], gen_istft_hop_size=cfg["gen_istft_hop_size"],
        n_speakers=cfg["n_speakers"], gin_channels=cfg["gin_channels"], use_sdp=False,
        ms_istft_vits=cfg["variant"] == "ms", mb_istft_vits=cfg["variant"] == "mb",
        istft_vits=cfg["variant"] == "istft",
        subbands=cfg["subbands"] if cfg["variant"] != "istft" else False,
    )


def build_synthesizer(cfg, sd, device="cpu"):
    """A real `SynthesizerTrn` (models.py:568) carrying the seeded dec.* / flow.* / emb_g.* weights of
    mb_istft_vits_b200.synth (strict key and shape check); enc_p / dp / enc_q keep their own (torch-seeded) init."""
    models = import_reference()
    net = models.SynthesizerTrn(59, 513, 32, **model_kwargs(cfg)).eval()
    ref_sd = net.state_dict()
    pre = ("dec.", "flow.", "emb_g.") + tuple(p for p in ("enc_q.", "enc_p.") if any(k.startswith(p) for k in sd))
    want = {k for k in ref_sd if k.startswith(pre)}
    have = {k for k in sd if k.startswith(pre)}
    assert want == have, f"key inventory mismatch: missing {sorted(want - have)[:5]} extra {sorted(have - want)[:5]}"
    for k in want:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), (k, ref_sd[k].shape, sd[k].shape)
    missing, unexpected = net.load_state_dict({k: sd[k] for k in have}, strict=False)
    assert not unexpected
    return net.to(device)
