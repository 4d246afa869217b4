"""Shared helpers for the parity tests: golden-case table and loaders."""
import os

import numpy as np
import torch

from mb_istft_vits_b200 import configs as cfgs
from mb_istft_vits_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> (config name, overrides)  -- must mirror tools/make_golden.py:CASES
GOLDEN_CASES = {
    "mini_mb": ("ljs_mini_mb_istft_vits", {}),
    "mb": ("ljs_mb_istft_vits", {}),
    "mb_gscale": ("ljs_mb_istft_vits", {}),
    "ms": ("ljs_ms_istft_vits", {}),
    "istft": ("ljs_istft_vits", {}),
    "mini_istft": ("ljs_mini_istft_vits", {}),
    "uudb_spk8": ("uudb_spk8_istft_vits", {}),
    "ms_spk": ("uudb_ms_istft_vits_ms", {}),
    "mb_resblock2": ("ljs_mb_istft_vits", {"resblock": "2", "resblock_dilation_sizes": [[1, 3], [1, 3], [1, 3]]}),
    "mb_long": ("ljs_mini_mb_istft_vits", {}),
    # minted from the reference's full SynthesizerTrn.infer() (tools/make_golden.py infer): BASELINE configs 1 and 4
    "infer_mini_mb": ("ljs_mini_mb_istft_vits", {}),
    "infer_istft": ("ljs_istft_vits", {}),
}


def load_case(name):
    cname, over = GOLDEN_CASES[name]
    cfg = cfgs.get_config(cname)
    cfg.update(over)
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    t = {k: torch.from_numpy(d[k]) for k in d.files if k not in ("meta", "g_scale", "lengths", "sid")}
    B, T, wseed, zseed = [int(v) for v in d["meta"]]
    sd = synth.make_state_dict(cfg, seed=wseed, g_scale=float(d["g_scale"]))
    return cfg, sd, t, dict(B=B, T=T, zseed=zseed, lengths=[int(v) for v in d["lengths"]],
                            sid=(torch.from_numpy(d["sid"]) if "sid" in d.files else None))


# voice-conversion / posterior-encoder cases (tools/make_golden.py vc): name -> config name
VC_CASES = {"vc_ms_spk": "uudb_ms_istft_vits_ms", "posterior_mini": "ljs_mini_mb_istft_vits"}


def load_vc_case(name):
    cfg = cfgs.get_config(VC_CASES[name])
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    t = {k: torch.from_numpy(d[k]) for k in d.files}
    sd = synth.make_state_dict(cfg, seed=1234, enc_q=True)
    return cfg, sd, t


# text-encoder cases (tools/make_golden.py text): name -> config name
TEXT_CASES = {"text_mb": "ljs_mb_istft_vits", "text_mini": "ljs_mini_mb_istft_vits", "text_short": "ljs_mb_istft_vits"}


def load_text_case(name):
    cfg = cfgs.get_config(TEXT_CASES[name])
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    t = {k: torch.from_numpy(d[k]) for k in d.files}
    sd = synth.make_state_dict(cfg, seed=1234, enc_p=True)
    return cfg, sd, t
